// TEST INFRASTRUCTURE.  Drives the REFERENCE'S OWN ProductMixture_<true> (src/product_mixture.{hpp,cc} over
// src/product_model.hpp, src/product_value.{hpp,cc}, src/indexed_vector.hpp, src/models.hpp, and -- for tared data --
// src/differ.{hpp,cc}, all compiled unmodified from /root/reference by `make -C oracle ref`) through a cat-kernel
// trajectory: for every ADD it prints the score vector of ProductMixture::score_value (score_diff when the model has a
// tare row) and then calls add_value (add_diff) with the packed group the script names; for every REMOVE it calls
// remove_value (remove_diff); at the end it prints score_data and the clustering counts.  The feature mixtures behind it
// are oracle/ref_stubs_mix/ (distributions' arithmetic delegated to the oracle's own per-feature functions), so what is
// compared is the reference's own code around them.  tests/golden/make_reference_product_mixture.py turns the output
// into the committed fixture the oracle's (and, through the -m gpu replay tests, the product's) sweep must reproduce.
//   stdin: alpha d | F | per feature: type dim hp0 hp1 hp2 hp3 [dim aux values] | has_tare [F tare tokens] |
//          R | R lines of F tokens ('-' = unobserved) | N | N lines "A row group" or "R row group"
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <loom/product_mixture.hpp>
#include <loom/differ.hpp>

static bool read_value_row(const std::vector<int> &types, loom::ProductValue &value) {
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

int main() {
    loom::ProductModel model;
    int F;
    if (std::scanf("%f %f %d", &model.clustering.alpha, &model.clustering.d, &F) != 3) return 2;
    std::vector<int> types(F);
    for (int f = 0; f < F; ++f) {
        int dim;
        float hp[4];
        if (std::scanf("%d %d %f %f %f %f", &types[f], &dim, &hp[0], &hp[1], &hp[2], &hp[3]) != 6) return 2;
        switch (types[f]) {                         // the containers ProductModel::load fills (src/product_model.cc:50-84)
            case LB_BB: read_spec(model.features.bb.insert(f), LB_BB, 0, hp, 0); break;
            case LB_DD: if (dim <= 16) read_spec(model.features.dd16.insert(f), LB_DD, dim, hp, dim);
                        else read_spec(model.features.dd256.insert(f), LB_DD, dim, hp, dim);
                        break;
            case LB_DPD: read_spec(model.features.dpd.insert(f), LB_DPD, dim, hp, dim); break;
            case LB_GP: read_spec(model.features.gp.insert(f), LB_GP, 0, hp, 0); break;
            case LB_NICH: read_spec(model.features.nich.insert(f), LB_NICH, 0, hp, 0); break;
            default: return 2;
        }
    }
    model.schema.load(model.features);
    int has_tare, R, N;
    if (std::scanf("%d", &has_tare) != 1) return 2;
    loom::ProductValue tare;
    if (has_tare && !read_value_row(types, tare)) return 2;
    if (std::scanf("%d", &R) != 1) return 2;
    std::vector<loom::protobuf::Row> rows(R);
    for (int r = 0; r < R; ++r) {
        rows[r].set_id(r);
        if (!read_value_row(types, *rows[r].mutable_diff()->mutable_pos())) return 2;
        rows[r].mutable_diff()->mutable_neg()->mutable_observed()->set_sparsity(loom::ProductValue::Observed::NONE);
    }
    if (has_tare) {                                 // `loom sparsify`: the reference's own Differ against this tare row
        loom::Differ differ(model.schema, tare);
        loom::protobuf::memory_files()["rows"] = rows;
        differ.compress_rows("rows", "diffs");
        rows = loom::protobuf::memory_files()["diffs"];
        model.tares.push_back(differ.get_tare());
    }

    loom::rng_t rng;
    loom::ProductMixture_<true> mixture;
    mixture.maintaining_cache = true;
    mixture.init_unobserved(model, std::vector<int>(1, 0), rng);     // one empty group, as a fresh kind starts
    if (std::scanf("%d", &N) != 1) return 2;
    loom::VectorFloat scores;
    std::printf("{\"adds\": [\n");
    bool first = true;
    for (int i = 0; i < N; ++i) {
        char what[8];
        int r, g;
        if (std::scanf("%7s %d %d", what, &r, &g) != 3) return 2;
        const loom::ProductValue::Diff &diff = rows[r].diff();
        if (what[0] == 'A') {
            if (has_tare) mixture.score_diff(model, diff, scores, rng);
            else mixture.score_value(model, diff.pos(), scores, rng);
            std::printf("%s[", first ? "" : ",\n");
            first = false;
            for (size_t k = 0; k < scores.size(); ++k) std::printf("%s%.9g", k ? "," : "", scores[k]);
            std::printf("]");
            if ((size_t)g >= scores.size()) return 3;
            if (has_tare) mixture.add_diff(model, g, diff, rng);
            else mixture.add_value(model, g, diff.pos(), rng);
        } else {
            if (has_tare) mixture.remove_diff(model, g, diff, rng);
            else mixture.remove_value(model, g, diff.pos(), rng);
        }
    }
    std::printf("\n], \"score_data\": %.9g, \"counts\": [", mixture.score_data(model, rng));
    const auto &counts = mixture.clustering.counts();
    for (size_t k = 0; k < counts.size(); ++k) std::printf("%s%d", k ? "," : "", counts[k]);
    std::printf("], \"packed_to_global\": [");
    for (size_t k = 0; k < counts.size(); ++k) std::printf("%s%u", k ? "," : "", mixture.id_tracker.packed_to_global(k));
    std::printf("]}\n");
    return 0;
}
