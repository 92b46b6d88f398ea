// TEST INFRASTRUCTURE.  FUNCTIONAL stand-ins for the mixture classes of forcedotcom/distributions 2.0.28 (un-vendored),
// so that the reference's OWN src/product_mixture.{hpp,cc}, src/product_model.hpp, src/product_value.{hpp,cc},
// src/indexed_vector.hpp and src/models.hpp compile UNMODIFIED and RUN (oracle/Makefile `ref`, target ref_product_mixture).
// This header shadows ref_stubs_value/distributions/models/feature_stub.hpp for that target only.
//
// What the library computes per feature -- the conjugate update of a group's sufficient statistics and its posterior
// predictive / marginal -- is delegated to the oracle's OWN per-feature functions (om_add_value, om_score_value,
// om_score_data of oracle/oracle_math.c, the ones pinned to scipy's closed forms by tests/test_oracle_math.py), so that
// what the compiled program pins is everything product_mixture.cc ITSELF does around them: the clustering prior + the
// sum over observed features, when groups are born and die and which packed index they get, the tare cache and
// score_diff / add_diff / remove_diff.  The bookkeeping classes the library owns (the Pitman-Yor clustering mixture,
// MixtureIdTracker, Packed_) are restated from their published behaviour and say so where they are defined.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <unordered_map>
#include <vector>
#include "../../../ref_stubs/distributions/random.hpp"
#include "../../../../include/loom_b200.h"

extern "C" {
// oracle/oracle_math.c (float64): stride = distance between the rows of one group's statistics
void om_add_value(const lb_feature_spec *s, uint32_t *ints, double *flts, int stride, uint32_t raw, int sign);
double om_score_value(const lb_feature_spec *s, const float *aux, const uint32_t *ints, const double *flts, int stride, uint32_t raw);
double om_score_data(const lb_feature_spec *s, const float *aux, const uint32_t *ints, const double *flts, int stride);
double lo_py_score_counts(double alpha, double d, const uint32_t *counts, int n);
}

namespace distributions {

template <class Value> struct Packed_;   // vector.hpp

inline uint32_t raw_cell(bool v) { return v ? 1u : 0u; }
inline uint32_t raw_cell(uint32_t v) { return v; }
inline uint32_t raw_cell(float v) { uint32_t r; std::memcpy(&r, &v, 4); return r; }

// The hyperparameters of one feature, with the member names the library's Shared structs have (src/hyper_prior.hpp
// writes `shared.alpha = ...`), and their image as the oracle's lb_feature_spec + aux array
struct SharedBase {
    int dim = 0;
    void add_value(...) {}                             // DPD dictionaries are complete at ingest
    void remove_value(...) {}
    void realize(rng_t &) {}
    template <class M> void protobuf_load(const M &) {}
    template <class M> void protobuf_dump(M &) const {}
};
struct BBShared : SharedBase {
    float alpha = 1, beta = 1;
    void load(int, const float *hp, const std::vector<float> &) { alpha = hp[0]; beta = hp[1]; }
    lb_feature_spec spec() const { lb_feature_spec s = {LB_BB, 0, {alpha, beta, 0, 0}, 0}; return s; }
    const float *aux() const { return nullptr; }
    int int_rows() const { return 2; }
    int flt_rows() const { return 0; }
};
struct DDShared : SharedBase {
    std::vector<float> alphas;
    void load(int d, const float *, const std::vector<float> &a) { dim = d; alphas = a; }
    lb_feature_spec spec() const { lb_feature_spec s = {LB_DD, dim, {0, 0, 0, 0}, 0}; return s; }
    const float *aux() const { return alphas.data(); }
    int int_rows() const { return dim + 1; }
    int flt_rows() const { return 0; }
};
struct DPDShared : SharedBase {
    float gamma = 1, alpha = 1, beta0 = 1;
    struct Betas {                                     // the library's sparse map value -> beta; dense here (value = index)
        std::vector<float> v;
        float &get(uint32_t value) { return v[value]; }
        float get(uint32_t value) const { return v[value]; }
        size_t size() const { return v.size(); }
    } betas;
    void load(int d, const float *hp, const std::vector<float> &a) { dim = d; gamma = hp[0]; alpha = hp[1]; beta0 = hp[2]; betas.v = a; }
    lb_feature_spec spec() const { lb_feature_spec s = {LB_DPD, dim, {gamma, alpha, beta0, 0}, 0}; return s; }
    const float *aux() const { return betas.v.data(); }
    int int_rows() const { return dim + 1; }
    int flt_rows() const { return 0; }
};
struct GPShared : SharedBase {
    float alpha = 1, inv_beta = 1;
    void load(int, const float *hp, const std::vector<float> &) { alpha = hp[0]; inv_beta = hp[1]; }
    lb_feature_spec spec() const { lb_feature_spec s = {LB_GP, 0, {alpha, inv_beta, 0, 0}, 0}; return s; }
    const float *aux() const { return nullptr; }
    int int_rows() const { return 2; }
    int flt_rows() const { return 1; }
};
struct NICHShared : SharedBase {
    float mu = 0, kappa = 1, sigmasq = 1, nu = 1;
    void load(int, const float *hp, const std::vector<float> &) { mu = hp[0]; kappa = hp[1]; sigmasq = hp[2]; nu = hp[3]; }
    lb_feature_spec spec() const { lb_feature_spec s = {LB_NICH, 0, {mu, kappa, sigmasq, nu}, 0}; return s; }
    const float *aux() const { return nullptr; }
    int int_rows() const { return 1; }
    int flt_rows() const { return 2; }
};

// Shared / Group / Mixture of one feature model; V is the library's Value type, S one of the structs above
template <class V, class S> struct FeatureStub {
    typedef V Value;
    typedef S Shared;
    struct Group {
        std::vector<uint32_t> ints;
        std::vector<double> flts;
        std::unordered_map<uint32_t, uint32_t> counts;   // named by the DPD branch of src/hyper_kernel.cc (never run here)
        void init(const Shared &s, rng_t &) { ints.assign(s.int_rows(), 0u); flts.assign(s.flt_rows() + 1, 0.0); }
        void add_value(const Shared &s, const Value &v, rng_t &) { const lb_feature_spec sp = s.spec(); om_add_value(&sp, ints.data(), flts.data(), 1, raw_cell(v), +1); }
        void add_repeated_value(const Shared &s, const Value &v, int count, rng_t &rng) { for (int i = 0; i < count; ++i) add_value(s, v, rng); }
        void remove_value(const Shared &s, const Value &v, rng_t &) { const lb_feature_spec sp = s.spec(); om_add_value(&sp, ints.data(), flts.data(), 1, raw_cell(v), -1); }
        float score_value(const Shared &s, const Value &v, rng_t &) const {
            const lb_feature_spec sp = s.spec();
            return (float)om_score_value(&sp, s.aux(), ints.data(), flts.data(), 1, raw_cell(v));
        }
        float score_data(const Shared &s, rng_t &) const {
            const lb_feature_spec sp = s.spec();
            return (float)om_score_data(&sp, s.aux(), ints.data(), flts.data(), 1);
        }
        Value sample_value(const Shared &, rng_t &) const { return Value(); }
        void validate(const Shared &) const {}
        template <class M> void protobuf_load(const M &) {}
        template <class M> void protobuf_dump(M &) const {}
    };
    struct Sampler {};
    // distributions' Mixture: the groups of one feature, in PACKED order; removal moves the last group into the hole
    // (Packed_::packed_remove), the cached form only adds score tables that do not change the results
    struct Mixture {
        typedef S Shared;
        std::vector<Group> groups_;
        std::vector<Group> &groups() { return groups_; }
        const std::vector<Group> &groups() const { return groups_; }
        Group &groups(size_t i) { return groups_[i]; }
        const Group &groups(size_t i) const { return groups_[i]; }
        void init(const Shared &, rng_t &) {}
        void add_group(const Shared &s, rng_t &rng) { groups_.emplace_back(); groups_.back().init(s, rng); }
        void remove_group(const Shared &, size_t groupid) {
            if (groupid + 1 != groups_.size()) groups_[groupid] = std::move(groups_.back());
            groups_.pop_back();
        }
        void add_value(const Shared &s, size_t groupid, const Value &v, rng_t &rng) { groups_[groupid].add_value(s, v, rng); }
        void remove_value(const Shared &s, size_t groupid, const Value &v, rng_t &rng) { groups_[groupid].remove_value(s, v, rng); }
        void score_value(const Shared &s, const Value &v, VectorFloat &scores_accum, rng_t &rng) const {
            for (size_t g = 0; g < groups_.size(); ++g) scores_accum[g] += groups_[g].score_value(s, v, rng);
        }
        float score_value_group(const Shared &s, size_t groupid, const Value &v, rng_t &rng) const { return groups_[groupid].score_value(s, v, rng); }
        float score_data(const Shared &s, rng_t &rng) const {
            float score = 0;
            for (const auto &g : groups_) score += g.score_data(s, rng);
            return score;
        }
        // the hyper kernel's grid step (src/infer_grid.hpp:83): the data score of all groups under every hypothesis
        void score_data_grid(const std::vector<Shared> &shareds, VectorFloat &scores, rng_t &rng) const {
            for (size_t i = 0; i < shareds.size(); ++i) scores[i] = score_data(shareds[i], rng);
        }
        void validate(const Shared &) const {}
    };
    typedef Mixture FastMixture;
    typedef Mixture SmallMixture;
    typedef FeatureStub Model;
    static uint32_t OTHER() { return 0xFFFFFFFFu; }
    static float MIN_BETA() { return 1e-6f; }
};
struct BetaBernoulli : FeatureStub<bool, BBShared> {};
template <int max_dim> struct DirichletDiscrete : FeatureStub<uint32_t, DDShared> {};
struct DirichletProcessDiscrete : FeatureStub<uint32_t, DPDShared> {};
struct GammaPoisson : FeatureStub<uint32_t, GPShared> {};
struct NormalInverseChiSq : FeatureStub<float, NICHShared> {};

// distributions::Clustering<int>::PitmanYor and its mixture, restated from the library's published behaviour
// (clustering.hpp of 2.0.28; the formulas are SURVEY.md 8c's): counts in packed order with exactly the empty groups the
// caller keeps; add_value reports that an EMPTY group was just occupied (the caller appends a fresh empty group to every
// feature, the mixture appends its count), remove_value reports that a group just emptied (the caller removes it
// everywhere, the mixture swap-removes its count).
template <class count_t> struct Clustering {
    struct PitmanYor {
        float alpha, d;
        PitmanYor() : alpha(1.f), d(0.f) {}
        template <class Message> void protobuf_load(const Message &m) { load_point(m, 0); }       // a hyper-prior grid point carries (alpha, d)
        template <class Message> auto load_point(const Message &m, int) -> decltype(m.alpha(), void()) { alpha = m.alpha(); d = m.d(); }
        template <class Message> void load_point(const Message &, long) {}
        template <class Message> void protobuf_dump(Message &) const {}
        float score_counts(const std::vector<int> &counts) const {
            std::vector<uint32_t> c(counts.begin(), counts.end());
            return (float)lo_py_score_counts(alpha, d, c.data(), (int)c.size());
        }
        std::vector<int> sample_assignments(size_t size, rng_t &) const { return scripted_seating(size); }
        struct Mixture {
            std::vector<int> counts_;
            std::vector<int> &counts() { return counts_; }
            const std::vector<int> &counts() const { return counts_; }
            int counts(size_t i) const { return counts_[i]; }
            void init(const PitmanYor &) {}
            bool add_value(const PitmanYor &, size_t groupid, int count = 1) {
                const bool add_group = counts_[groupid] == 0;
                counts_[groupid] += count;
                if (add_group) counts_.push_back(0);
                return add_group;
            }
            bool remove_value(const PitmanYor &, size_t groupid, int count = 1) {
                counts_[groupid] -= count;
                const bool remove_group = counts_[groupid] == 0;
                if (remove_group) {
                    if (groupid + 1 != counts_.size()) counts_[groupid] = counts_.back();
                    counts_.pop_back();
                }
                return remove_group;
            }
            // log predictive of joining each group: (n_g - d) / (n + alpha) for an occupied one, the new-group mass
            // (alpha + d m) / (n + alpha) shared by the e empty ones
            void score_value(const PitmanYor &model, VectorFloat &scores) const {
                double n = 0;
                int m = 0, e = 0;
                for (int c : counts_) { n += c; m += c > 0; e += c == 0; }
                for (size_t g = 0; g < counts_.size(); ++g)
                    scores[g] = counts_[g] ? (float)std::log((counts_[g] - (double)model.d) / (n + model.alpha))
                                           : (float)std::log((model.alpha + (double)model.d * m) / ((n + model.alpha) * e));
            }
            float score_data(const PitmanYor &model) const {
                std::vector<int> occupied;
                for (int c : counts_) if (c) occupied.push_back(c);
                return model.score_counts(occupied);
            }
        };
    };
};
template <class Model, class count_t> struct MixtureDriver : Model::Mixture {};
struct ProtobufStub { struct Shared {}; struct Group {}; };
template <class Model> struct Protobuf { typedef ProtobufStub t; };
namespace protobuf { struct Clustering { struct PitmanYor {}; }; }

// distributions::MixtureIdTracker (mixture.hpp of 2.0.28), restated: stable GLOBAL ids for groups whose PACKED index
// changes when another group is removed.  A new group takes the next unused global id; removing packed index i moves the
// last packed group to i.
struct MixtureIdTracker {
    std::vector<uint32_t> packed_to_global_;
    std::unordered_map<uint32_t, uint32_t> global_to_packed_;
    size_t global_size_ = 0;
    void init(size_t group_count = 0) {
        packed_to_global_.clear(); global_to_packed_.clear(); global_size_ = 0;
        for (size_t i = 0; i < group_count; ++i) add_group();
    }
    void add_group() {
        const uint32_t packed = (uint32_t)packed_to_global_.size(), global = (uint32_t)global_size_++;
        packed_to_global_.push_back(global);
        global_to_packed_[global] = packed;
    }
    void remove_group(size_t packed) {
        const uint32_t global = packed_to_global_[packed];
        global_to_packed_.erase(global);
        if (packed + 1 != packed_to_global_.size()) {
            const uint32_t moved = packed_to_global_.back();
            packed_to_global_[packed] = moved;
            global_to_packed_[moved] = (uint32_t)packed;
        }
        packed_to_global_.pop_back();
    }
    size_t packed_size() const { return packed_to_global_.size(); }
    size_t global_size() const { return global_size_; }
    uint32_t packed_to_global(size_t packed) const { return packed_to_global_[packed]; }
    uint32_t global_to_packed(size_t global) const { return global_to_packed_.at((uint32_t)global); }
};

// names the DPD branch of src/hyper_kernel.cc uses (Stirling row, safe Dirichlet draw, fast_log / fast_lgamma): the
// harnesses hold no DPD column when they run the hyper kernel (the oracle's DPD sampler is a documented statistical
// equivalent, not a restatement), so these only have to exist
inline void get_log_stirling1_row(size_t, VectorFloat &) { std::abort(); }
inline void sample_dirichlet_safe(rng_t &, size_t, const float *, float *, float) { std::abort(); }
inline float fast_log(float x) { return std::log(x); }
inline float fast_lgamma(float x) { return std::lgamma(x); }
inline void vector_zero(size_t size, float *x) { for (size_t i = 0; i < size; ++i) x[i] = 0.f; }
inline void vector_negate(size_t size, float *x) { for (size_t i = 0; i < size; ++i) x[i] = -x[i]; }
inline void vector_add(size_t size, float *x, const float *y) { for (size_t i = 0; i < size; ++i) x[i] += y[i]; }
inline size_t sample_from_probs(rng_t &rng, const VectorFloat &probs) {
    float t = (float)rng.next();
    for (size_t i = 0; i < probs.size(); ++i) { t -= probs[i]; if (t < 0) return i; }
    return probs.size() - 1;
}
}  // namespace distributions
namespace protobuf { namespace distributions {} }
