// TEST INFRASTRUCTURE.  Stand-ins for the headers of forcedotcom/distributions 2.0.28 (un-vendored) that the reference's
// value path includes (src/common.hpp, src/models.hpp), so that src/product_value.{hpp,cc} and src/differ.{hpp,cc}
// compile UNMODIFIED from where they lie (oracle/Makefile `ref`).  The value path uses none of the library's
// arithmetic: only the names of the feature-model types and their Value typedefs, which is all these files declare.
#pragma once
#include "../../ref_stubs/distributions/random.hpp"
