// TEST INFRASTRUCTURE: plain structs with the accessor names of the generated protobuf classes schedules.hpp reads
// (src/schema.proto: Config.Schedule, Checkpoint.Schedule) -- the generated code needs protoc, which is not here.
#pragma once
#include <cstdint>
namespace loom { namespace protobuf {
struct Config {
    struct Schedule {
        double extra_passes_ = 0, small_data_size_ = 0, big_data_size_ = 0, checkpoint_period_sec_ = 1e9;
        uint32_t max_reject_iters_ = 0;
        double extra_passes() const { return extra_passes_; }
        double small_data_size() const { return small_data_size_; }
        double big_data_size() const { return big_data_size_; }
        uint32_t max_reject_iters() const { return max_reject_iters_; }
        double checkpoint_period_sec() const { return checkpoint_period_sec_; }
    };
};
struct Checkpoint {
    struct Schedule {
        double annealing_state_ = 0;
        uint64_t row_count_ = 0;
        uint32_t reject_iters_ = 0;
        double annealing_state() const { return annealing_state_; }
        void set_annealing_state(double x) { annealing_state_ = x; }
        uint64_t row_count() const { return row_count_; }
        void set_row_count(uint64_t x) { row_count_ = x; }
        uint32_t reject_iters() const { return reject_iters_; }
        void set_reject_iters(uint32_t x) { reject_iters_ = x; }
    };
};
}}  // namespace loom::protobuf
