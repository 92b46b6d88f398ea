// TEST INFRASTRUCTURE: the clock CheckpointingSchedule reads (src/timer.hpp), restated with <chrono>.
#pragma once
#include <chrono>
#include <cstdint>
namespace loom {
typedef uint64_t usec_t;
inline usec_t current_time_usec() {
    return (usec_t)std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
}  // namespace loom
