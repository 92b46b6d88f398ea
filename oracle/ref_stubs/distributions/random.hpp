// TEST INFRASTRUCTURE.  Stand-in for <distributions/random.hpp> of forcedotcom/distributions 2.0.28 (un-vendored; see
// DESIGN.md section 2) so that the reference's OWN src/kind_proposer.cc compiles here.  `rng_t` is a SCRIPTED source:
// the uniforms come from a list the harness supplies, so that the reference's decisions can be compared draw for draw
// with the oracle's (which consumes the same uniforms from its Philox stream).  sample_from_likelihoods restates the
// library's published inverse-CDF walk (random.hpp: t = total * unif01; subtract until negative; last index otherwise).
#pragma once
#include <cmath>      // the library's headers bring <cmath> in; src/schedules.hpp relies on it
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <numeric>
#include <utility>
#include <vector>
namespace distributions {
// distributions::Packed_ (vector.hpp of 2.0.28), restated from its published behaviour: a vector whose removal moves the
// LAST element into the hole -- the renumbering KindKernel::remove_featureless_kind (src/kind_kernel.cc:196-206) and the
// group bookkeeping of ProductMixture rely on.  VectorFloat is the library's Packed_<float> (aligned there).
template <class Value> struct Packed_ : std::vector<Value> {
    Packed_() {}
    explicit Packed_(size_t size) : std::vector<Value>(size) {}
    Packed_(size_t size, const Value &value) : std::vector<Value>(size, value) {}
    Value &packed_add() { this->emplace_back(); return this->back(); }
    Value &packed_add(const Value &value) { this->push_back(value); return this->back(); }
    void packed_remove(size_t pos) {
        if (pos + 1 != this->size()) (*this)[pos] = std::move(this->back());
        this->pop_back();
    }
};
typedef Packed_<float> VectorFloat;
struct rng_t;
inline std::vector<std::pair<uint64_t, const std::vector<double> *>> &seeded_scripts() {    // scripts for generators built from a seed
    static std::vector<std::pair<uint64_t, const std::vector<double> *>> scripts;
    return scripts;
}
struct rng_t {
    typedef uint64_t result_type;
    const std::vector<double> *script;
    size_t pos;
    rng_t() : script(nullptr), pos(0) {}
    // generators the reference builds itself: KindKernel's `rng_(seed)` finds the script the harness registered for
    // that seed; the per-feature `rng_t(seed + f)` of infer_assignments find none and never draw
    explicit rng_t(uint64_t seed) : script(nullptr), pos(0) {
        for (const auto &s : seeded_scripts()) if (s.first == seed) script = s.second;
    }
    explicit rng_t(const std::vector<double> &s) : script(&s), pos(0) {}
    uint64_t operator()() { return 0; }                            // `seed = rng()`: does not consume the script
    double next() {
        if (!script || pos >= script->size()) std::abort();        // a draw the harness did not foresee is an error
        return (*script)[pos++];
    }
};
inline std::vector<uint32_t> &pick_log() { static std::vector<uint32_t> log; return log; }   // every draw, in order
inline std::vector<float> &total_log() { static std::vector<float> log; return log; }      // and the total it was drawn against
inline float sample_unif01(rng_t &rng) { return (float)rng.next(); }
inline size_t sample_from_likelihoods(rng_t &rng, const VectorFloat &likelihoods, float total_likelihood) {
    const size_t size = likelihoods.size();
    float t = total_likelihood * sample_unif01(rng);
    for (size_t i = 0; i < size; ++i) {
        t -= likelihoods[i];
        if (t < 0) { pick_log().push_back((uint32_t)i); total_log().push_back(total_likelihood); return i; }
    }
    pick_log().push_back((uint32_t)(size - 1)); total_log().push_back(total_likelihood);
    return size - 1;
}
// names src/infer_grid.hpp and src/kind_kernel.hpp use; never reached by the harnesses (grids of size 0, no rows scored)
inline int sample_int(rng_t &rng, int low, int high) { return low + (int)(rng.next() * (high - low + 1)); }
// sample_from_scores_overwrite: the cat kernel's draw.  The harness SCRIPTS it -- the next entry of pick_script() is the
// group to take (the oracle's own pick for that row and kind) -- and keeps the score vector it was called with
inline std::vector<uint32_t> &pick_script() { static std::vector<uint32_t> s; return s; }
inline size_t &pick_script_pos() { static size_t p = 0; return p; }
inline std::vector<std::vector<float>> &scores_log() { static std::vector<std::vector<float>> l; return l; }
inline bool &draw_from_uniforms() { static bool b = false; return b; }
inline size_t sample_from_scores_overwrite(rng_t &rng, VectorFloat &scores) {
    scores_log().emplace_back(scores.begin(), scores.end());
    if (draw_from_uniforms()) {
        // the library's published algorithm (random.hpp / special.hpp of 2.0.28): subtract the maximum, exponentiate in
        // place, walk the inverse CDF with one uniform -- here the generator's next scripted uniform
#ifdef LB_STUB_DOUBLE_SAMPLER                               // diagnosis only: is a disagreement float32 rounding at a CDF boundary?
        typedef double real_t;
#else
        typedef float real_t;
#endif
        real_t max_score = scores[0], total = 0;
        for (float s : scores) if (s > max_score) max_score = s;
        std::vector<real_t> likelihoods(scores.size());
        for (size_t i = 0; i < scores.size(); ++i) total += likelihoods[i] = std::exp((real_t)scores[i] - max_score);
        real_t t = total * (real_t)rng.next();
        for (size_t i = 0; i < scores.size(); ++i) { t -= likelihoods[i]; if (t < 0) return i; }
        return scores.size() - 1;
    }
    if (pick_script_pos() >= pick_script().size()) std::abort();
    const size_t pick = pick_script()[pick_script_pos()++];
    if (pick >= scores.size()) std::abort();
    return pick;
}
// PitmanYor::sample_assignments: the prior partition of a fresh ephemeral kind, scripted the same way (one vector per
// kind, in creation order); without a script every row alternates between two groups
inline std::vector<std::vector<int>> &seating_script() { static std::vector<std::vector<int>> s; return s; }
inline size_t &seating_script_pos() { static size_t p = 0; return p; }
inline std::vector<int> scripted_seating(size_t size) {
    if (seating_script_pos() < seating_script().size()) {
        std::vector<int> a = seating_script()[seating_script_pos()++];
        if (a.size() != size) std::abort();
        return a;
    }
    std::vector<int> a(size);
    for (size_t i = 0; i < size; ++i) a[i] = (int)(i & 1);
    return a;
}
}  // namespace distributions
