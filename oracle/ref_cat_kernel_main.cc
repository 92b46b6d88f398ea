// TEST INFRASTRUCTURE.  The reference's OWN CatKernel (src/cat_kernel.hpp) over its own CrossCat, ValueSplitter, Assignments,
// ProductModel and ProductMixture (compiled unmodified from /root/reference by `make -C oracle ref`; functional mixtures
// of oracle/ref_stubs_mix/) on a multi-kind CrossCat: every scripted action goes through CatKernel::add_row /
// remove_row(rng, row, assignments) -- split the row over the kinds, per kind score, draw (scripted: the oracle's pick),
// add, push the GLOBAL group id (src/cat_kernel.hpp:199-299).  With a tare row the rows are first compressed by the
// reference's own Differ and the kernel takes the score_diff / add_diff / remove_diff branch.  Printed: the score
// vector of every draw, and at the end per kind the clustering counts and every row's group as the assignments queue
// holds it (global id) and as the id tracker maps it back (packed id), and CrossCat::score_data.
//   stdin: see tests/golden/make_reference_cat_kernel.py (the only writer of this format)
#include <cstdio>
#include <cstring>
#include <vector>
#include <loom/cat_kernel.hpp>
#include <loom/differ.hpp>

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


int main() {
    loom::CrossCat cross_cat;
    float kind_alpha, kind_d;
    int F, R;
    if (std::scanf("%f %f %f %f %d", &cross_cat.topology.alpha, &cross_cat.topology.d, &kind_alpha, &kind_d, &F) != 5) return 2;
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
    cross_cat.mixture_init_unobserved(1, unused);

    int has_tare;
    if (std::scanf("%d", &has_tare) != 1) return 2;
    loom::ProductValue tare;
    if (has_tare && !read_row(types, tare)) return 2;
    if (std::scanf("%d", &R) != 1) return 2;
    std::vector<loom::protobuf::Row> rows(R);
    for (int r = 0; r < R; ++r) {
        rows[r].set_id(1000 + r);
        if (!read_row(types, *rows[r].mutable_diff()->mutable_pos())) return 2;
        rows[r].mutable_diff()->mutable_neg()->mutable_observed()->set_sparsity(loom::ProductValue::Observed::NONE);
    }
    if (has_tare) {                                  // `loom sparsify`, then CrossCat::tares_load's second half
        loom::Differ differ(cross_cat.schema, tare);
        loom::protobuf::memory_files()["rows"] = rows;
        differ.compress_rows("rows", "diffs");
        rows = loom::protobuf::memory_files()["diffs"];
        cross_cat.tares.push_back(differ.get_tare());
        std::vector<loom::ProductValue *> temp_values;
        cross_cat.update_tares(temp_values, unused);
    }
    loom::Assignments assignments;
    for (int k = 0; k < K; ++k) assignments.packed_add();

    int n_actions, n_picks;
    if (std::scanf("%d", &n_actions) != 1) return 2;
    std::vector<std::pair<char, int>> actions(n_actions);
    for (auto &a : actions) { char what[4]; if (std::scanf("%3s %d", what, &a.second) != 2) return 2; a.first = what[0]; }
    if (std::scanf("%d", &n_picks) != 1) return 2;
    distributions::pick_script().resize(n_picks);
    for (auto &p : distributions::pick_script()) if (std::scanf("%u", &p) != 1) return 2;

    loom::protobuf::Config::Kernels config;
    loom::CatKernel cat_kernel(config.cat(), cross_cat);
    loom::rng_t rng;
    for (const auto &a : actions) {
        if (a.first == 'A') cat_kernel.add_row(rng, rows[a.second], assignments);
        else cat_kernel.remove_row(rng, rows[a.second], assignments);
    }
    if (distributions::pick_script_pos() != distributions::pick_script().size()) return 4;

    std::printf("{\"scores\": [");
    const auto &log = distributions::scores_log();
    for (size_t i = 0; i < log.size(); ++i) {
        std::printf("%s[", i ? "," : "");
        for (size_t g = 0; g < log[i].size(); ++g) std::printf("%s%.9g", g ? "," : "", log[i][g]);
        std::printf("]");
    }
    std::printf("],\n \"score_data\": %.9g, \"rowids\": [", cross_cat.score_data(rng));
    for (size_t i = 0; i < assignments.row_count(); ++i) std::printf("%s%llu", i ? "," : "", (unsigned long long)assignments.rowids()[i]);
    std::printf("], \"kinds\": [\n");
    for (int k = 0; k < K; ++k) {
        const auto &kind = cross_cat.kinds[k];
        std::printf("%s{\"counts\": [", k ? ",\n" : "");
        const auto &counts = kind.mixture.clustering.counts();
        for (size_t g = 0; g < counts.size(); ++g) std::printf("%s%d", g ? "," : "", counts[g]);
        std::printf("], \"global\": [");
        for (size_t i = 0; i < assignments.row_count(); ++i) std::printf("%s%u", i ? "," : "", assignments.groupids(k)[i]);
        std::printf("], \"packed\": [");
        for (size_t i = 0; i < assignments.row_count(); ++i)
            std::printf("%s%u", i ? "," : "", kind.mixture.id_tracker.global_to_packed(assignments.groupids(k)[i]));
        std::printf("]}");
    }
    std::printf("\n]}\n");
    return 0;
}
