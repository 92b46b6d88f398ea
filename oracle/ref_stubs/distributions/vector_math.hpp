// TEST INFRASTRUCTURE.  Stand-in for <distributions/vector_math.hpp> (distributions 2.0.28): the one function
// src/kind_proposer.cc calls, restated from the library's published definition (special.hpp: subtract the maximum,
// exponentiate in place, return the total), in float32 like the library.
#pragma once
#include <cmath>
#include <vector>
#include <distributions/random.hpp>
#define DIST_ASSUME_ALIGNED(data) (data)
namespace distributions {
inline float scores_to_likelihoods(VectorFloat &scores) {
    float max_score = scores[0];
    for (float s : scores) if (s > max_score) max_score = s;
    float total = 0;
    for (float &s : scores) total += s = std::exp(s - max_score);
    return total;
}
}  // namespace distributions
