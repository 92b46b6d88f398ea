// TEST INFRASTRUCTURE.  Drives the REFERENCE'S OWN KindKernel (src/kind_kernel.{hpp,cc} + src/kind_proposer.{hpp,cc},
// compiled unmodified from /root/reference by `make -C oracle ref`; oracle/ref_stubs_kind/ stands in for CrossCat and
// the mixtures, oracle/ref_stubs_value/ and oracle/ref_stubs/ for protoc's classes and the un-vendored distributions
// library) through its whole life inside Loom::infer_kind_structure (src/loom.cc:399-450): construction (E ephemeral
// kinds appended), a number of try_run() calls on score matrices read from stdin, destruction (ephemeral kinds
// dropped).  After every step it prints the feature -> kind map, the number of kinds and which kind's tables sit at
// which position, as one JSON document.  tests/golden/make_reference_kind_kernel.py turns that into the committed
// fixture the oracle's kind_run + init_featureless_kinds must reproduce.
//   stdin: alpha d F K E iterations n_rows | f2k[F] | n_rounds | per round: n_kinds, scores[n_kinds][F] (row p = the
//          kind at position p), n_uniforms, u[n_uniforms]
#include <cstdio>
#include <vector>
#include <loom/kind_kernel.hpp>

static void print_state(const char *name, const loom::CrossCat &cross_cat, const loom::Logger::Message *log, bool last) {
    std::printf("{\"step\": \"%s\", \"f2k\": [", name);
    for (size_t f = 0; f < cross_cat.featureid_to_kindid.size(); ++f) std::printf("%s%u", f ? "," : "", cross_cat.featureid_to_kindid[f]);
    std::printf("], \"n_kinds\": %zu, \"labels\": [", cross_cat.kinds.size());
    for (size_t k = 0; k < cross_cat.kinds.size(); ++k) std::printf("%s%d", k ? "," : "", cross_cat.kinds[k].mixture.label);
    std::printf("], \"n_features\": [");
    for (size_t k = 0; k < cross_cat.kinds.size(); ++k) std::printf("%s%zu", k ? "," : "", cross_cat.kinds[k].featureids.size());
    std::printf("]");
    if (log) std::printf(", \"change_count\": %llu, \"birth_count\": %llu, \"death_count\": %llu", (unsigned long long)log->kind_.change_count,
                         (unsigned long long)log->kind_.birth_count, (unsigned long long)log->kind_.death_count);
    std::printf("}%s\n", last ? "" : ",");
}

int main() {
    double alpha, d;
    int F, K, E, iterations, n_rows, n_rounds;
    if (std::scanf("%lf %lf %d %d %d %d %d", &alpha, &d, &F, &K, &E, &iterations, &n_rows) != 7) return 2;
    loom::CrossCat cross_cat;
    cross_cat.topology.alpha = (float)alpha;
    cross_cat.topology.d = (float)d;
    cross_cat.featureid_to_kindid.resize(F);
    for (int k = 0; k < K; ++k) cross_cat.kinds.packed_add();
    for (int f = 0; f < F; ++f) {
        if (std::scanf("%u", &cross_cat.featureid_to_kindid[f]) != 1) return 2;
        cross_cat.kinds[cross_cat.featureid_to_kindid[f]].featureids.insert(f);
    }
    loom::rng_t unused;
    for (int k = 0; k < K; ++k) cross_cat.kinds[k].mixture.init_unobserved(cross_cat.kinds[k].model, std::vector<int>(1, n_rows), unused);
    loom::Assignments assignments;
    for (int r = 0; r < n_rows; ++r) assignments.rowids().try_push(1000 + r);
    for (int k = 0; k < K; ++k) {
        auto &queue = assignments.packed_add();
        for (int r = 0; r < n_rows; ++r) queue.push(0);
    }
    if (std::scanf("%d", &n_rounds) != 1) return 2;
    std::vector<std::vector<std::vector<float>>> scores(n_rounds);
    std::vector<double> uniforms;
    for (int round = 0; round < n_rounds; ++round) {
        int n_kinds, n;
        if (std::scanf("%d", &n_kinds) != 1) return 2;
        scores[round].assign(n_kinds, std::vector<float>(F));
        for (auto &row : scores[round]) for (auto &x : row) if (std::scanf("%f", &x) != 1) return 2;
        if (std::scanf("%d", &n) != 1) return 2;
        for (int i = 0; i < n; ++i) { double u; if (std::scanf("%lf", &u) != 1) return 2; uniforms.push_back(u); }
    }
    const uint64_t seed = 20141;
    distributions::seeded_scripts().push_back(std::make_pair(seed, &uniforms));

    loom::protobuf::Config::Kernels config;
    config.kind_.empty_kind_count_ = (uint32_t)E;
    config.kind_.iterations_ = (uint32_t)iterations;
    std::printf("{\"steps\": [\n");
    {
        loom::KindKernel kind_kernel(config, cross_cat, assignments, seed);
        print_state("constructed", cross_cat, nullptr, false);
        for (int round = 0; round < n_rounds; ++round) {
            if (scores[round].size() != cross_cat.kinds.size()) { std::fprintf(stderr, "round %d: %zu score rows for %zu kinds\n", round, scores[round].size(), cross_cat.kinds.size()); return 3; }
            auto &by_label = loom::score_script().by_label;
            by_label.assign(loom::score_script().next_label, std::vector<float>());
            for (size_t p = 0; p < cross_cat.kinds.size(); ++p) by_label[cross_cat.kinds[p].mixture.label] = scores[round][p];
            kind_kernel.try_run();
            loom::Logger::Message log;
            kind_kernel.log_metrics(log);
            print_state("try_run", cross_cat, &log, false);
        }
    }
    print_state("destroyed", cross_cat, nullptr, true);
    std::printf("]}\n");
    return 0;
}
