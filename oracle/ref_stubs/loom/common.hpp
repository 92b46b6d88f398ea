// TEST INFRASTRUCTURE.  Minimal stand-in for the reference's <loom/common.hpp> so that ITS OWN src/schedules.hpp
// compiles here without the un-vendored `distributions` headers (oracle/Makefile target `_ref/ref_schedules`).
// Only what schedules.hpp uses: the assertion macros and the branch hints.  Nothing of the reference is copied.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#define LOOM_LIKELY(x) __builtin_expect(bool(x), true)
#define LOOM_UNLIKELY(x) __builtin_expect(bool(x), false)
#define LOOM_ASSERT(cond, message) do { if (!(cond)) { std::cerr << "ERROR " << message << std::endl; std::abort(); } } while (0)
#define LOOM_ASSERT_LE(x, y) LOOM_ASSERT((x) <= (y), "expected " #x " <= " #y)
#define LOOM_ASSERT_LT(x, y) LOOM_ASSERT((x) < (y), "expected " #x " < " #y)
#define LOOM_ASSERT_EQ(x, y) LOOM_ASSERT((x) == (y), "expected " #x " == " #y)
#define LOOM_ASSERT1(cond, message) LOOM_ASSERT(cond, message)
#define LOOM_ASSERT2(cond, message) LOOM_ASSERT(cond, message)
#define LOOM_DEBUG_LEVEL 0
