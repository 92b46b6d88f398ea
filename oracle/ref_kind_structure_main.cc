// TEST INFRASTRUCTURE.  One kind-structure pass of the reference, for real: the reference's OWN CrossCat, KindKernel,
// KindProposer, ProductModel, ProductMixture, ValueSplitter and Assignments (src/cross_cat.{hpp,cc}, kind_kernel.{hpp,cc},
// kind_proposer.{hpp,cc}, product_model.{hpp,cc}, product_mixture.{hpp,cc}, product_value.{hpp,cc}, assignments.hpp,
// compiled unmodified from /root/reference by `make -C oracle ref`) over the functional mixtures of oracle/ref_stubs_mix/
// (per-feature arithmetic = the oracle's own om_* functions).  The program follows Loom::infer_kind_structure
// (src/loom.cc:399-450): a KindKernel is constructed over an unassigned CrossCat and init_cache()d, then per round the
// scripted actions go through KindKernel::add_row / remove_row (cat kernel + kind proposer accumulation, a9) and
// try_run() closes the round (a10 score_feature for every (feature, kind), a11 block sampler, a12 move_feature_to and
// the ephemeral-kind bookkeeping); the kernel is destroyed at the end.  Every random decision is scripted from the
// oracle's own run -- the group of every ADD, the uniforms of the block sampler, the prior partitions of fresh
// ephemeral kinds -- so what is printed and compared is what the reference COMPUTES along that trajectory: the score
// vector of every draw, L[f][k], the new feature -> kind map, and at the end every kind's sufficient statistics and
// CrossCat::score_data.  tests/golden/make_reference_kind_structure.py makes the committed fixture from it.
//   stdin: see tests/golden/make_reference_kind_structure.py (the only writer of this format)
#include <cstdio>
#include <cstring>
#include <vector>
#include <loom/kind_kernel.hpp>
#include <distributions/vector_math.hpp>

static bool read_row(const std::vector<int> &types, loom::ProductValue &value) {
    char token[64];
    value.Clear();
    value.mutable_observed()->set_sparsity(loom::ProductValue::Observed::DENSE);
    for (size_t f = 0; f < types.size(); ++f) {
        if (std::scanf("%63s", token) != 1) return false;
        const bool observed = std::strcmp(token, "-") != 0;
        value.mutable_observed()->add_dense(observed);
        if (!observed) continue;
        if (types[f] == LB_BB) value.add_booleans(std::atoi(token) != 0);
        else if (types[f] == LB_NICH) value.add_reals(std::strtof(token, nullptr));
        else value.add_counts((uint32_t)std::strtoul(token, nullptr, 10));
    }
    return true;
}

template <class S> static void read_spec(S &shared, int type, int dim, const float *hp, int aux_n) {
    (void)type;
    std::vector<float> aux(aux_n);
    for (auto &a : aux) if (std::scanf("%f", &a) != 1) std::abort();
    shared.load(dim, hp, aux);
}

struct print_stats_fun {                      // every feature of a kind: its groups' statistics in packed order
    bool first;
    template <class T, class Mixture> void operator()(T *, size_t, const Mixture &) {}
};
struct print_kind_fun {
    const loom::CrossCat::Kind &kind;
    bool first;
    template <class T> void operator()(T *t) {
        const auto &mixtures = kind.mixture.features[t];
        for (size_t i = 0; i < mixtures.size(); ++i) {
            std::printf("%s{\"feature\": %zu, \"ints\": [", first ? "" : ", ", (size_t)mixtures.index(i));
            first = false;
            const auto &groups = mixtures[i].groups();
            const auto &shared = kind.model.features[t][i];
            for (size_t g = 0; g < groups.size(); ++g) {
                std::printf("%s[", g ? "," : "");
                for (int j = 0; j < shared.int_rows(); ++j) std::printf("%s%u", j ? "," : "", groups[g].ints[j]);
                std::printf("]");
            }
            std::printf("], \"flts\": [");
            for (size_t g = 0; g < groups.size(); ++g) {
                std::printf("%s[", g ? "," : "");
                for (int j = 0; j < shared.flt_rows(); ++j) std::printf("%s%.17g", j ? "," : "", groups[g].flts[j]);
                std::printf("]");
            }
            std::printf("]}");
        }
    }
};

static void print_map(const loom::CrossCat &cross_cat) {
    std::printf("\"f2k\": [");
    for (size_t f = 0; f < cross_cat.featureid_to_kindid.size(); ++f) std::printf("%s%u", f ? "," : "", cross_cat.featureid_to_kindid[f]);
    std::printf("], \"n_kinds\": %zu", cross_cat.kinds.size());
}

int main() {
    loom::CrossCat cross_cat;
    float kind_alpha, kind_d;
    int E, iterations, F, R, n_rounds;
    if (std::scanf("%f %f %f %f %d %d %d", &cross_cat.topology.alpha, &cross_cat.topology.d, &kind_alpha, &kind_d, &E, &iterations, &F) != 7) return 2;
    std::vector<int> types(F), kind_of(F);
    struct Spec { int type, dim; float hp[4]; std::vector<float> aux; };
    std::vector<Spec> specs(F);
    int K = 0;
    for (int f = 0; f < F; ++f) {
        Spec &s = specs[f];
        if (std::scanf("%d %d %d %f %f %f %f", &kind_of[f], &s.type, &s.dim, &s.hp[0], &s.hp[1], &s.hp[2], &s.hp[3]) != 7) return 2;
        const int aux_n = s.type == LB_DD || s.type == LB_DPD ? s.dim : 0;
        s.aux.resize(aux_n);
        for (auto &a : s.aux) if (std::scanf("%f", &a) != 1) return 2;
        types[f] = s.type;
        if (kind_of[f] + 1 > K) K = kind_of[f] + 1;
    }
    cross_cat.featureid_to_kindid.assign(kind_of.begin(), kind_of.end());
    for (int k = 0; k < K; ++k) cross_cat.kinds.packed_add();
    for (int f = 0; f < F; ++f) {               // the containers ProductModel::load fills (src/product_model.cc:50-84)
        loom::CrossCat::Kind &kind = cross_cat.kinds[kind_of[f]];
        const Spec &s = specs[f];
        kind.featureids.insert(f);
#define LB_PUT(container) kind.model.features.container.insert(f).load(s.dim, s.hp, s.aux)
        switch (s.type) {
            case LB_BB: LB_PUT(bb); break;
            case LB_DD: if (s.dim <= 16) LB_PUT(dd16); else LB_PUT(dd256); break;
            case LB_DPD: LB_PUT(dpd); break;
            case LB_GP: LB_PUT(gp); break;
            case LB_NICH: LB_PUT(nich); break;
            default: return 2;
        }
#undef LB_PUT
    }
    for (int k = 0; k < K; ++k) {
        loom::CrossCat::Kind &kind = cross_cat.kinds[k];
        kind.model.clustering.alpha = kind_alpha;
        kind.model.clustering.d = kind_d;
        kind.model.schema.load(kind.model.features);
        cross_cat.schema += kind.model.schema;
    }
    cross_cat.update_splitter();
    loom::rng_t unused;
    cross_cat.mixture_init_unobserved(1, unused);            // Loom::Loom with no groups_in (src/loom.cc:60-62)

    if (std::scanf("%d", &R) != 1) return 2;
    std::vector<loom::protobuf::Row> rows(R);
    for (int r = 0; r < R; ++r) {
        rows[r].set_id(r);
        if (!read_row(types, *rows[r].mutable_diff()->mutable_pos())) return 2;
        rows[r].mutable_diff()->mutable_neg()->mutable_observed()->set_sparsity(loom::ProductValue::Observed::NONE);
    }
    loom::Assignments assignments;
    for (int k = 0; k < K; ++k) assignments.packed_add();

    std::vector<double> uniforms;
    const uint64_t seed = 20142;
    distributions::seeded_scripts().push_back(std::make_pair(seed, &uniforms));
    loom::protobuf::Config::Kernels config;
    config.kind_.empty_kind_count_ = (uint32_t)E;
    config.kind_.iterations_ = (uint32_t)iterations;
    if (std::scanf("%d", &n_rounds) != 1) return 2;
    std::printf("{\"rounds\": [\n");
    {
        loom::KindKernel kind_kernel(config, cross_cat, assignments, seed);
        kind_kernel.init_cache();
        for (int round = 0; round < n_rounds; ++round) {
            int n_actions, n_picks, n_uniforms, n_seatings;
            if (std::scanf("%d", &n_actions) != 1) return 2;
            std::vector<std::pair<char, int>> actions(n_actions);
            for (auto &a : actions) { char what[4]; if (std::scanf("%3s %d", what, &a.second) != 2) return 2; a.first = what[0]; }
            if (std::scanf("%d", &n_picks) != 1) return 2;
            distributions::pick_script().resize(n_picks);
            distributions::pick_script_pos() = 0;
            for (auto &p : distributions::pick_script()) if (std::scanf("%u", &p) != 1) return 2;
            if (std::scanf("%d", &n_uniforms) != 1) return 2;
            for (int i = 0; i < n_uniforms; ++i) { double u; if (std::scanf("%lf", &u) != 1) return 2; uniforms.push_back(u); }
            if (std::scanf("%d", &n_seatings) != 1) return 2;
            distributions::seating_script().assign(n_seatings, std::vector<int>());
            distributions::seating_script_pos() = 0;
            for (auto &s : distributions::seating_script()) {
                int n;
                if (std::scanf("%d", &n) != 1) return 2;
                s.resize(n);
                for (auto &g : s) if (std::scanf("%d", &g) != 1) return 2;
            }
            distributions::scores_log().clear();
            distributions::feature_scores_log().clear();
            for (const auto &a : actions) {
                if (a.first == 'A') kind_kernel.add_row(rows[a.second]);
                else kind_kernel.remove_row(rows[a.second]);
            }
            if (distributions::pick_script_pos() != distributions::pick_script().size()) return 4;
            std::printf("%s{\"scores\": [", round ? ",\n" : "");
            const auto &log = distributions::scores_log();
            for (size_t i = 0; i < log.size(); ++i) {
                std::printf("%s[", i ? "," : "");
                for (size_t g = 0; g < log[i].size(); ++g) std::printf("%s%.9g", g ? "," : "", log[i][g]);
                std::printf("]");
            }
            std::printf("],\n \"kinds_scored\": %zu, ", cross_cat.kinds.size());
            kind_kernel.try_run();
            std::printf("\"L\": [");
            const auto &L = distributions::feature_scores_log();
            for (size_t f = 0; f < L.size(); ++f) {
                std::printf("%s[", f ? "," : "");
                for (size_t k = 0; k < L[f].size(); ++k) std::printf("%s%.9g", k ? "," : "", L[f][k]);
                std::printf("]");
            }
            std::printf("],\n ");
            print_map(cross_cat);
            std::printf("}");
        }
    }
    std::printf("\n], \"final\": {");
    print_map(cross_cat);
    loom::rng_t rng;
    std::printf(", \"score_data\": %.9g, \"kinds\": [\n", cross_cat.score_data(rng));
    for (size_t k = 0; k < cross_cat.kinds.size(); ++k) {
        const auto &kind = cross_cat.kinds[k];
        std::printf("%s{\"counts\": [", k ? ",\n" : "");
        const auto &counts = kind.mixture.clustering.counts();
        for (size_t g = 0; g < counts.size(); ++g) std::printf("%s%d", g ? "," : "", counts[g]);
        std::printf("], \"features\": [");
        print_kind_fun fun = {kind, true};
        loom::for_each_feature_type(fun);
        std::printf("]}");
    }
    std::printf("\n]}}\n");
    return 0;
}
