// TEST INFRASTRUCTURE.  Stand-in for <distributions/vector_math.hpp> (distributions 2.0.28): the one function
// src/kind_proposer.cc calls, restated from the library's published definition (special.hpp: subtract the maximum,
// exponentiate in place, return the total), in float32 like the library.
#pragma once
#include <cmath>
#include <vector>
#include <distributions/random.hpp>
#define DIST_ASSUME_ALIGNED(data) (data)
namespace distributions {
inline std::vector<std::vector<float>> &feature_scores_log() { static std::vector<std::vector<float>> l; return l; }   // L[f][.] as infer_assignments computed it
inline float scores_to_likelihoods(VectorFloat &scores) {
    feature_scores_log().emplace_back(scores.begin(), scores.end());
    float max_score = scores[0];
    for (float s : scores) if (s > max_score) max_score = s;
    float total = 0;
    for (float &s : scores) total += s = std::exp(s - max_score);
    return total;
}
}  // namespace distributions
