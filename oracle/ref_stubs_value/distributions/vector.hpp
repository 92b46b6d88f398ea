// TEST INFRASTRUCTURE: see random_fwd.hpp.  VectorFloat and Packed_ come with the scripted rng stand-in.
#pragma once
#include <vector>
#include "../../ref_stubs/distributions/random.hpp"
