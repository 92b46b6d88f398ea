// TEST INFRASTRUCTURE: see random_fwd.hpp.  VectorFloat comes with the scripted rng stand-in.
#pragma once
#include <vector>
#include "../../ref_stubs/distributions/random.hpp"
namespace distributions {
// distributions::Packed_ (vector.hpp of 2.0.28), restated from its published behaviour: a vector whose removal moves the
// LAST element into the hole -- the renumbering KindKernel::remove_featureless_kind relies on (src/kind_kernel.cc:196-206)
template <class Value> struct Packed_ : std::vector<Value> {
    Value &packed_add() { this->emplace_back(); return this->back(); }
    void packed_remove(size_t pos) {
        if (pos + 1 != this->size()) (*this)[pos] = std::move(this->back());
        this->pop_back();
    }
};
}  // namespace distributions
