// Test harness: the C++ host's schedule classes (loom_b200/host/loom_b200.hpp) driven through the loop of
// loom::b200::infer_multi_pass, printing every ADD / REMOVE decision and batch boundary in the format of
// tests/golden/reference_schedules.json (produced by the reference's own src/schedules.hpp).  Only the classes are
// used: nothing of libloomb200 is called, so the program links without it.
//   usage: host_schedules EXTRA_PASSES SMALL BIG ROWS [MAX_ACTIONS]
#include <cstdio>
#include <cstdlib>
#include "../../loom_b200/host/loom_b200.hpp"

int main(int argc, char **argv) {
    if (argc < 5) return 2;
    loom::b200::ScheduleConfig config;
    config.extra_passes = std::atof(argv[1]);
    config.small_data_size = std::atof(argv[2]);
    config.big_data_size = std::atof(argv[3]);
    const uint64_t rows = (uint64_t)std::atoll(argv[4]);
    const uint64_t max_actions = argc > 5 ? (uint64_t)std::atoll(argv[5]) : ~uint64_t(0);
    loom::b200::AcceleratingSchedule accelerating(config);
    loom::b200::AnnealingSchedule annealing(accelerating.extra_passes(0));
    loom::b200::BatchingSchedule batching;
    uint64_t assigned = 0, actions = 0;
    while (assigned != rows && actions < max_actions) {
        if (annealing.next_action_is_add()) { ++assigned; batching.add(); std::putchar('A'); }
        else { --assigned; batching.remove(); std::putchar('R'); }
        ++actions;
        if (batching.test()) {
            annealing.set_extra_passes(accelerating.extra_passes(assigned));
            std::putchar('|');
        }
    }
    std::printf("\n%llu\n", (unsigned long long)assigned);
    return 0;
}
