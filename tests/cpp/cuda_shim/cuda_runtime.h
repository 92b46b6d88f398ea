// Test shim: lets loom_b200/csrc/lb_common.cuh -- the device-side score arithmetic -- compile for the HOST, so that the
// float32 formulas the kernels run can be checked against the float64 oracle on the CPU suite (tests/test_device_math.py).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#ifndef __CUDACC__
#define __CUDACC__ 1          /* the header keeps its device functions behind this */
#endif
static inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline float __frcp_rn(float x) { return 1.0f / x; }       // round-to-nearest reciprocal
