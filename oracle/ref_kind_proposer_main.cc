// TEST INFRASTRUCTURE.  Drives the REFERENCE'S OWN KindProposer::infer_assignments (src/kind_proposer.cc, compiled
// unmodified from /root/reference by `make -C oracle ref`, with the reference's own kind_proposer.hpp; oracle/ref_stubs_kind/
// stands in for CrossCat and the mixtures, oracle/ref_stubs_value/ and oracle/ref_stubs/ for protoc's classes and the
// un-vendored distributions library) on a score matrix, a feature -> kind assignment and a list of uniforms read
// from stdin, and prints every draw of the block Pitman-Yor sampler and the final assignment as one JSON line.
// tests/golden/make_reference_kind_proposer.py turns that into the committed fixture the oracle's sampler
// (oracle_structure.c: oracle_block_py) must reproduce draw for draw.
//   stdin: alpha d F K iterations | assign[F] | scores[F][K] (row f = feature f) | n | u[n]
#include <cstdio>
#include <vector>
#include <loom/kind_proposer.hpp>

int main() {
    double alpha, d;
    int F, K, iterations, n;
    if (std::scanf("%lf %lf %d %d %d", &alpha, &d, &F, &K, &iterations) != 5) return 2;
    std::vector<uint32_t> assign(F);
    for (auto &a : assign) if (std::scanf("%u", &a) != 1) return 2;
    std::vector<std::vector<float>> &by_kind = loom::score_script().by_label;     // kind k's tables carry label k
    by_kind.assign(K, std::vector<float>(F));
    for (int f = 0; f < F; ++f)
        for (int k = 0; k < K; ++k) if (std::scanf("%f", &by_kind[k][f]) != 1) return 2;
    if (std::scanf("%d", &n) != 1) return 2;
    std::vector<double> uniforms(n);
    for (auto &u : uniforms) if (std::scanf("%lf", &u) != 1) return 2;

    loom::CrossCat cross_cat;
    cross_cat.topology.alpha = (float)alpha;
    cross_cat.topology.d = (float)d;
    for (int k = 0; k < K; ++k) cross_cat.kinds.packed_add();
    loom::KindProposer proposer;
    proposer.kinds.resize(K);
    for (int k = 0; k < K; ++k) proposer.kinds[k].mixture.label = k;             // proposer kind k shadows kind k
    loom::rng_t rng(uniforms);
    proposer.infer_assignments(cross_cat, assign, (size_t)iterations, false, rng);

    std::printf("{\"picks\": [");
    const auto &picks = distributions::pick_log();
    for (size_t i = 0; i < picks.size(); ++i) std::printf("%s%u", i ? "," : "", picks[i]);
    std::printf("], \"totals\": [");
    const auto &totals = distributions::total_log();
    for (size_t i = 0; i < totals.size(); ++i) std::printf("%s%.9g", i ? "," : "", totals[i]);
    std::printf("], \"assignments\": [");
    for (int f = 0; f < F; ++f) std::printf("%s%u", f ? "," : "", assign[f]);
    std::printf("]}\n");
    return 0;
}
