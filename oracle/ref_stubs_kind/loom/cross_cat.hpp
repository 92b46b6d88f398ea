// TEST INFRASTRUCTURE.  Stand-in for <loom/cross_cat.hpp> (and product_model.hpp / product_mixture.hpp behind it, whose
// arithmetic lives in the un-vendored distributions library) shaped so that the reference's OWN src/kind_kernel.{hpp,cc}
// and src/kind_proposer.{hpp,cc} compile UNMODIFIED from where they lie (oracle/Makefile `ref`).  CrossCat keeps the
// members of the real struct that the kind kernel's BOOKKEEPING touches -- kinds (a Packed_ vector), each kind's
// featureids, featureid_to_kindid, topology -- with their real types; the mixtures are reduced to no-ops, except
// score_feature, which returns the entry of a score matrix the harness supplies, and a label that follows a kind's
// tables through every move.  What the compiled program pins is therefore everything kind_kernel.cc itself decides:
// which features move, which kinds die, how swap-removal renumbers the survivors, where the new ephemeral kinds go.
#pragma once
#include <numeric>
#include <unordered_set>
#include <vector>
#include <loom/common.hpp>
#include <loom/protobuf.hpp>
#include <loom/models.hpp>
#include <loom/product_value.hpp>

namespace loom {

struct ScoreScript {                       // harness-supplied: scores[label of the kind's tables][feature]
    std::vector<std::vector<float>> by_label;
    int next_label = 0;
};
inline ScoreScript &score_script() { static ScoreScript s; return s; }

struct ProductModel {
    int schema = 0;
    std::vector<ProductValue> tares;
    Clustering::Shared clustering;
    void clear() {}
    void extend(const ProductModel &) {}
    template <class V> void add_value(const V &, rng_t &) {}
    template <class V> void remove_value(const V &, rng_t &) {}
    template <class D> void add_diff(const D &, rng_t &) {}
    template <class D> void remove_diff(const D &, rng_t &) {}
};

struct IdTrackerStub {
    size_t packed_to_global(size_t i) const { return i; }
    size_t global_to_packed(size_t i) const { return i; }
};

template <bool cached> struct ProductMixture_ {
    struct Counts : std::vector<int> { int label = -1; };     // the counts of a kind's clustering, tagged with whose they are
    struct ClusteringStub {
        Counts counts_;
        const Counts &counts() const { return counts_; }
    } clustering;
    IdTrackerStub id_tracker;
    bool maintaining_cache = false;
    int label = -1;                        // which kind's tables these are (proposer mixtures: the kind they shadow)
    // KindKernel::add_featureless_kind: a new kind's own counts
    void init_unobserved(const ProductModel &, const std::vector<int> &counts, rng_t &) {
        static_cast<std::vector<int> &>(clustering.counts_) = counts;
        clustering.counts_.label = label;
    }
    // KindProposer::mixture_init_unobserved: proposer kind i is rebuilt from cross_cat.kinds[i]'s counts and from then on
    // shadows that kind -- it answers score_feature with that kind's row of the score script
    template <class OtherCounts> void init_unobserved(const ProductModel &, const OtherCounts &counts, rng_t &) {
        static_cast<std::vector<int> &>(clustering.counts_) = counts;
        label = clustering.counts_.label = counts.label;
    }
    size_t count_rows() const { return (size_t)std::accumulate(clustering.counts_.begin(), clustering.counts_.end(), 0); }
    template <class V> void score_value(const ProductModel &, const V &, VectorFloat &, rng_t &) const {}
    template <class D> void score_diff(const ProductModel &, const D &, VectorFloat &, rng_t &) const {}
    template <class V> void add_value(const ProductModel &, size_t, const V &, rng_t &) {}
    template <class V> void remove_value(const ProductModel &, size_t, const V &, rng_t &) {}
    template <class D> void add_diff(const ProductModel &, size_t, const D &, rng_t &) {}
    template <class D> void remove_diff(const ProductModel &, size_t, const D &, rng_t &) {}
    template <class D> void add_diff_step_1_of_2(const ProductModel &, size_t, const D &, rng_t &) {}
    void add_diff_step_2_of_2(const ProductModel &, rng_t &) {}
    void remove_unobserved_value(const ProductModel &, size_t) {}
    void init_feature_cache(const ProductModel &, size_t, rng_t &) {}
    void init_tare_cache(const ProductModel &, rng_t &) {}
    void validate(const ProductModel &) const {}
    template <bool other> void validate_subset(const ProductMixture_<other> &) const {}
    float score_feature(const ProductModel &, size_t featureid, rng_t &) const { return score_script().by_label.at(label).at(featureid); }
    template <class Other>
    void move_feature_to(size_t, ProductModel &, Other &, ProductModel &, Other &) {}
};
typedef ProductMixture_<true> ProductMixture;
typedef ProductMixture_<false> SmallProductMixture;

struct CrossCat : noncopyable {
    struct Kind {
        ProductModel model;
        ProductMixture mixture;
        std::unordered_set<size_t> featureids;
        // every kind born anywhere (the harness's originals, KindKernel::add_featureless_kind's packed_add) takes the
        // next label: the harness reads the labels back to see which kind ended where
        Kind() { mixture.label = score_script().next_label++; }
    };
    int schema = 0;
    std::vector<ProductValue> tares;
    struct SplitterStub {
        template <class D, class P> void split(const D &, P &) const {}
    } splitter;
    protobuf::HyperPrior hyper_prior;                 // the stub's clustering grid is empty
    Clustering::Shared topology;
    distributions::Packed_<Kind> kinds;
    std::vector<uint32_t> featureid_to_kindid;
    template <class P> void simplify(P &) const {}
    void update_splitter() {}
    template <class T> void update_tares(T &, rng_t &) {}
    void validate() const {
        for (size_t f = 0; f < featureid_to_kindid.size(); ++f) {
            const auto &ids = kinds[featureid_to_kindid[f]].featureids;
            LOOM_ASSERT(ids.find(f) != ids.end(), "feature " << f << " is not in the kind the map names");
        }
    }
};

}  // namespace loom
