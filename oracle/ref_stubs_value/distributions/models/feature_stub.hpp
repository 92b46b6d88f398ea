// TEST INFRASTRUCTURE: see ../random_fwd.hpp.  The feature-model types src/models.hpp wraps, reduced to the typedefs
// it names.  Value types follow the library's published ones (bb: bool; dd, dpd, gp: uint32_t; nich: float).
#pragma once
#include <cstdint>
#include <vector>
#include "../../../ref_stubs/distributions/random.hpp"
namespace distributions {
template <class V> struct FeatureStub {
    typedef V Value;
    struct Shared {}; struct Group {}; struct Sampler {}; struct FastMixture {}; struct SmallMixture {};
};
struct BetaBernoulli : FeatureStub<bool> {};
template <int max_dim> struct DirichletDiscrete : FeatureStub<uint32_t> {};
struct DirichletProcessDiscrete : FeatureStub<uint32_t> {};
struct GammaPoisson : FeatureStub<uint32_t> {};
struct NormalInverseChiSq : FeatureStub<float> {};
template <class count_t> struct Clustering {
    struct PitmanYor {
        float alpha, d;
        struct Mixture {};
        template <class Message> void protobuf_load(const Message &) {}
        float score_counts(const std::vector<int> &) const { return 0.f; }
        // the prior partition of an ephemeral kind: not what this program pins (the oracle seats with its own Philox
        // stream); two groups, so that the group bookkeeping around it has something to count
        std::vector<int> sample_assignments(size_t size, rng_t &) const { return scripted_seating(size); }
    };
};
template <class Model, class count_t> struct MixtureDriver {};
struct ProtobufStub { struct Shared {}; struct Group {}; };
template <class Model> struct Protobuf { typedef ProtobufStub t; };
namespace protobuf { struct Clustering { struct PitmanYor {}; }; }
}  // namespace distributions
namespace protobuf { namespace distributions {} }
