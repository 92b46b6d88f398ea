// TEST INFRASTRUCTURE.  Drives the REFERENCE'S OWN schedule classes (compiled from /root/reference/src/schedules.hpp
// where it lies, with the reference's own common.hpp / protobuf.hpp / timer.hpp / assignments.hpp behind it;
// oracle/ref_stubs_value/ and oracle/ref_stubs/ stand in for protoc's generated classes, libprotobuf's streams and the
// un-vendored distributions headers) through the loop of
// Loom::infer_kind_structure_sequential (src/loom.cc:217-262) and prints, per configuration, every ADD / REMOVE decision
// and every batch boundary.  tests/golden/make_reference_schedules.py turns the output into the committed fixture
// tests/golden/reference_schedules.json that the schedule driver of the product (loom_b200/infer.py, loom.hpp) is
// checked against.
//   usage: ref_schedules EXTRA_PASSES SMALL BIG ROWS [MAX_ACTIONS]
#include <cstdio>
#include <cstdlib>
#include <loom/schedules.hpp>

int main(int argc, char **argv) {
    if (argc < 5) return 2;
    loom::protobuf::Config::Schedule config;
    config.extra_passes_ = std::atof(argv[1]);
    config.small_data_size_ = std::atof(argv[2]);
    config.big_data_size_ = std::atof(argv[3]);
    config.max_reject_iters_ = 100;
    const size_t rows = (size_t)std::atoll(argv[4]);
    const size_t max_actions = argc > 5 ? (size_t)std::atoll(argv[5]) : ~size_t(0);
    loom::AcceleratingSchedule accelerating(config);
    loom::AnnealingSchedule annealing(config);
    loom::BatchingSchedule batching(config);
    // Loom::infer_multi_pass (src/loom.cc:167-215) sets the rate from the rows already assigned before the loop
    annealing.set_extra_passes(accelerating.extra_passes(0));
    size_t assigned = 0, actions = 0, batches = 0;
    std::printf("{\"extra_passes\": %s, \"small\": %s, \"big\": %s, \"rows\": %zu, \"decisions\": \"", argv[1], argv[2], argv[3], rows);
    while (assigned != rows && actions < max_actions) {
        if (annealing.next_action_is_add()) { ++assigned; batching.add(); std::putchar('A'); }
        else { --assigned; batching.remove(); std::putchar('R'); }
        ++actions;
        if (batching.test()) {
            annealing.set_extra_passes(accelerating.extra_passes(assigned));
            ++batches;
            std::putchar('|');
        }
    }
    std::printf("\", \"actions\": %zu, \"batches\": %zu, \"assigned\": %zu}\n", actions, batches, assigned);
    return 0;
}
