// TEST INFRASTRUCTURE: see ../random_fwd.hpp.
#pragma once
#include <distributions/models/feature_stub.hpp>
