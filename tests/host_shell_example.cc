// Compile-and-link check of the C++ host shell (run only where a GPU exists).
#include <cstring>
#include "../loom_b200/host/loom_b200.hpp"

// compiled, never called: the query / generate entry points of the shell type-check against the ABI
[[maybe_unused]] static void api_surface(loom::b200::CrossCat &cc) {
    (void)cc.query_score(0, cc.row_count());
    (void)cc.query_sample(0, std::vector<uint32_t>{0}, 4, 1);
    cc.generate_rows(10, 0.5f, 2);
}

int main(int argc, char **argv) {
    if (argc < 2 || std::strcmp(argv[1], "run") != 0) return 0;  // link check only
    using namespace loom::b200;
    CrossCat cross_cat(0);
    std::vector<lb_feature_spec> feats(2);
    for (auto &f : feats) { f.type = LB_BB; f.dim = 0; f.hp[0] = 0.5f; f.hp[1] = 0.5f; f.aux_off = 0; }
    const float topo[2] = {2.f, 0.1f};
    cross_cat.model_load(feats, {0.f}, {0, 0}, {2.f, 0.1f}, topo);
    RowBlock rows;
    rows.row_count = 64;
    rows.cat8.assign(2 * 64, 1);
    rows.mask.assign(2 * 2, 0xFFFFFFFFu);
    cross_cat.rows_load(rows);
    CatKernel cat(cross_cat, 1);
    for (uint64_t r = 0; r < 64; ++r) cat.add_row(r);
    cat.wait();
    std::printf("score %f kinds %zu\n", cross_cat.score_data(), cross_cat.kind_count());
    // two more samples over the same rows, swept together (loom/tasks.py:173-194)
    CrossCat second(0);
    second.model_load(feats, {0.f}, {0, 0}, {2.f, 0.1f}, topo);
    second.rows_share(cross_cat);
    for (uint64_t r = 0; r < 64; ++r) cat.remove_row(r);
    cat.wait();
    MultiCatKernel multi({&cross_cat, &second}, {7, 8});
    for (uint64_t r = 0; r < 64; ++r) multi.add_row(r);
    multi.wait();
    std::printf("scores %f %f\n", cross_cat.score_data(), second.score_data());
    // a whole annealed run on the first chain (schedules of src/schedules.hpp)
    for (uint64_t r = 0; r < 64; ++r) cat.remove_row(r);
    cat.wait();
    ScheduleConfig sched;
    sched.extra_passes = 3.0; sched.small_data_size = 16; sched.big_data_size = 1e4;
    const uint64_t batches = infer_multi_pass(cross_cat, cat, nullptr, nullptr, sched, 5);
    std::printf("batches %llu score %f\n", (unsigned long long)batches, cross_cat.score_data());
    return 0;
}
