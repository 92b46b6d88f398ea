// TEST INFRASTRUCTURE: see random_fwd.hpp.  VectorFloat comes with the scripted rng stand-in.
#pragma once
#include <vector>
#include "../../ref_stubs/distributions/random.hpp"
