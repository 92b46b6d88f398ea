// TEST INFRASTRUCTURE: see ../random_fwd.hpp.  The feature-model types src/models.hpp wraps, reduced to the typedefs
// it names.  Value types follow the library's published ones (bb: bool; dd, dpd, gp: uint32_t; nich: float).
#pragma once
#include <cstdint>
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
    struct PitmanYor { float alpha, d; struct Mixture {}; };
};
template <class Model, class count_t> struct MixtureDriver {};
template <class Model> struct Protobuf { typedef int t; };
namespace protobuf { struct Clustering { struct PitmanYor {}; }; }
}  // namespace distributions
namespace protobuf { namespace distributions {} }
