// TEST INFRASTRUCTURE.  Stand-in for <loom/logger.hpp>: the message KindKernel::log_metrics fills (src/kind_kernel.hpp:133-145).
#pragma once
#include <cstdint>
namespace loom {
struct Logger {
    struct Message {
        struct Kind {
            uint64_t total_count = 0, change_count = 0, birth_count = 0, death_count = 0;
            void set_total_count(uint64_t x) { total_count = x; }
            void set_change_count(uint64_t x) { change_count = x; }
            void set_birth_count(uint64_t x) { birth_count = x; }
            void set_death_count(uint64_t x) { death_count = x; }
            void set_tare_time(uint64_t) {}
            void set_score_time(uint64_t) {}
            void set_sample_time(uint64_t) {}
            void set_total_time(uint64_t) {}
        } kind_;
        struct Hyper { void set_total_time(uint64_t) {} } hyper_;
        struct KernelStatus {
            Kind *kind; Hyper *hyper;
            Kind *mutable_kind() { return kind; }
            Hyper *mutable_hyper() { return hyper; }
            Hyper *mutable_cat() { return hyper; }      // Cat status: total_time only, like Hyper
        } status_{&kind_, &hyper_};
        KernelStatus *mutable_kernel_status() { return &status_; }
    };
};
}  // namespace loom
