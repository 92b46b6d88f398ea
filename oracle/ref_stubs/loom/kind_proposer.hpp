// TEST INFRASTRUCTURE.  Stand-in for <loom/kind_proposer.hpp> and the headers it pulls in (cross_cat.hpp,
// product_model.hpp, product_mixture.hpp: all of them need the un-vendored distributions library and protoc), shaped
// so that the reference's OWN src/kind_proposer.cc compiles UNMODIFIED from where it lies (oracle/Makefile `ref`).
// KindProposer is declared member for member as the .cc defines it; its collaborators are reduced to what the .cc
// touches, and SmallProductMixture::score_feature -- the marginal likelihood the real class computes with
// distributions' arithmetic -- returns the entry of a score matrix the harness supplies.  What the compiled program
// therefore pins is everything kind_proposer.cc ITSELF does with those scores: scores_to_likelihoods per feature, the
// block Pitman-Yor sampler's prior / empty-kind bookkeeping / posterior, the order of draws.
#pragma once
#include <cstdint>
#include <vector>
#include <loom/common.hpp>
#include <loom/timer.hpp>
#include <distributions/random.hpp>
#include <distributions/vector_math.hpp>
namespace distributions {
template <class count_t> struct Clustering {
    struct PitmanYor { float alpha, d; };
};
}  // namespace distributions
namespace loom {
using distributions::rng_t;
using distributions::VectorFloat;
struct TimedScope { explicit TimedScope(usec_t &) {} };
struct ProductModel {
    int schema = 0;
    std::vector<int> tares;
    void clear() {}
    void extend(const ProductModel &) {}
};
struct SmallProductMixture {
    struct ClusteringStub { std::vector<int> counts_; const std::vector<int> &counts() const { return counts_; } } clustering;
    bool maintaining_cache = false;
    const std::vector<float> *scores = nullptr;       // [feature_count] scores of this kind, supplied by the harness
    void init_unobserved(const ProductModel &, const std::vector<int> &, rng_t &) {}
    void add_diff_step_2_of_2(const ProductModel &, rng_t &) {}
    void validate_subset(const SmallProductMixture &) const {}
    void validate(const ProductModel &) const {}
    float score_feature(const ProductModel &, size_t featureid, rng_t &) const { return (*scores)[featureid]; }
};
struct CrossCat {
    struct Kind { ProductModel model; SmallProductMixture mixture; };
    int schema = 0;
    std::vector<int> tares;
    std::vector<Kind> kinds;
    distributions::Clustering<int>::PitmanYor topology;
};
struct KindProposer {
    struct Kind { ProductModel model; SmallProductMixture mixture; };
    std::vector<Kind> kinds;
    void clear() { kinds.clear(); }
    void model_load(const CrossCat &cross_cat);
    void mixture_init_unobserved(const CrossCat &cross_cat, rng_t &rng);
    struct Timers { usec_t tare, score, sample; };
    Timers infer_assignments(const CrossCat &cross_cat, std::vector<uint32_t> &featureid_to_kindid, size_t iterations,
                             bool parallel, rng_t &rng);

private:
    static void model_load(const CrossCat &cross_cat, ProductModel &model);
    class BlockPitmanYorSampler;
};
}  // namespace loom
