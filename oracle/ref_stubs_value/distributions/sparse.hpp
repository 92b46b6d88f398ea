// TEST INFRASTRUCTURE: see random_fwd.hpp.  src/common.hpp only declares stream operators over these two templates.
#pragma once
#include <unordered_map>
namespace distributions {
template <class Key, class Value> struct Sparse_ : std::unordered_map<Key, Value> {};
template <class Key, class Value> struct SparseCounter : std::unordered_map<Key, Value> {};
}  // namespace distributions
