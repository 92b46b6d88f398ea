// TEST INFRASTRUCTURE.  The reference's OWN HyperKernel::run (src/hyper_kernel.{hpp,cc} over infer_grid.hpp, hyper_prior.hpp
// and the real CrossCat / ProductModel / ProductMixture, compiled unmodified from /root/reference by `make -C oracle ref`)
// on a CrossCat whose rows were added to scripted groups, over the functional mixtures of oracle/ref_stubs_mix/ (data
// scores = the oracle's own om_score_data).  Every task draws from the uniforms the oracle's task of the same number
// reads, through the library's published sample_from_scores_overwrite, so the two must CHOOSE the same grid point for
// every coordinate of every feature, every kind's clustering and the topology -- which they only do if the reference
// scores the same hypotheses in the same order from the same statistics.  No DPD column: the oracle's DPD sampler is a
// documented statistical equivalent of src/hyper_kernel.cc:90-181, not a restatement.
//   stdin: see tests/golden/make_reference_hyper.py (the only writer of this format)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <loom/hyper_kernel.hpp>

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


struct print_hypers_fun {                     // every feature of a kind: its Shared after the run
    const loom::CrossCat::Kind &kind;
    std::vector<std::vector<float>> &out;     // [feature] = {hp0..hp3, aux...}
    template <class T> void operator()(T *t) {
        const auto &shareds = kind.model.features[t];
        for (size_t i = 0; i < shareds.size(); ++i) {
            const lb_feature_spec s = shareds[i].spec();
            std::vector<float> &v = out[shareds.index(i)];
            v.assign(s.hp, s.hp + 4);
            const int aux_n = s.type == LB_DD || s.type == LB_DPD ? s.dim : 0;
            for (int j = 0; j < aux_n; ++j) v.push_back(shareds[i].aux()[j]);
        }
    }
};

static void read_grid(google::protobuf::RepeatedField<float> &grid) {
    int n;
    if (std::scanf("%d", &n) != 1) std::abort();
    for (int i = 0; i < n; ++i) { float x; if (std::scanf("%f", &x) != 1) std::abort(); grid.Add(x); }
}
static void read_py_grid(::protobuf::loom::HyperPrior::PitmanYorGrid &grid) {
    int n;
    if (std::scanf("%d", &n) != 1) std::abort();
    grid.points.resize(n);
    for (auto &p : grid.points) if (std::scanf("%f %f", &p.alpha_, &p.d_) != 2) std::abort();
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

    // the state the kernel runs on: every row added to the packed group the script names, kind by kind
    if (std::scanf("%d", &R) != 1) return 2;
    std::vector<loom::ProductValue::Diff> partial_diffs;
    for (int r = 0; r < R; ++r) {
        loom::protobuf::Row row;
        row.set_id(r);
        if (!read_row(types, *row.mutable_diff()->mutable_pos())) return 2;
        row.mutable_diff()->mutable_neg()->mutable_observed()->set_sparsity(loom::ProductValue::Observed::NONE);
        cross_cat.splitter.split(row.diff(), partial_diffs);
        for (int k = 0; k < K; ++k) {
            unsigned g;
            if (std::scanf("%u", &g) != 1) return 2;
            cross_cat.kinds[k].mixture.add_value(cross_cat.kinds[k].model, g, partial_diffs[k].pos(), unused);
        }
    }

    auto &prior = cross_cat.hyper_prior;
    read_py_grid(prior.topology_);
    read_py_grid(prior.clustering_);
    read_grid(prior.bb_.alpha_); read_grid(prior.bb_.beta_);
    read_grid(prior.dd_.alpha_);
    read_grid(prior.gp_.alpha_); read_grid(prior.gp_.inv_beta_);
    read_grid(prior.nich_.mu_); read_grid(prior.nich_.kappa_); read_grid(prior.nich_.sigmasq_); read_grid(prior.nich_.nu_);

    // HyperKernel::run gives task t the generator rng_t(seed + t) (src/hyper_kernel.cc:203-207); seed = rng() is 0 here,
    // so task t finds the uniforms the harness registers under t: the ones the oracle's task t reads from its stream
    const int n_tasks = 1 + K + F;
    std::vector<std::vector<double>> uniforms(n_tasks);
    for (int t = 0; t < n_tasks; ++t) {
        int n;
        if (std::scanf("%d", &n) != 1) return 2;
        uniforms[t].resize(n);
        for (auto &u : uniforms[t]) if (std::scanf("%lf", &u) != 1) return 2;
        distributions::seeded_scripts().push_back(std::make_pair((uint64_t)t, &uniforms[t]));
    }
    distributions::draw_from_uniforms() = true;
    loom::protobuf::Config::Kernels config;
    loom::HyperKernel hyper_kernel(config.hyper(), cross_cat);
    loom::rng_t rng;
    hyper_kernel.run(rng);

    if (std::getenv("LB_REF_DUMP_SCORES"))                     // diagnosis: every score vector the kernel drew from, in order
        for (const auto &v : distributions::scores_log()) {
            for (float x : v) std::fprintf(stderr, "%.9g ", x);
            std::fprintf(stderr, "\n");
        }
    std::printf("{\"topology\": [%.9g, %.9g], \"kinds\": [", cross_cat.topology.alpha, cross_cat.topology.d);
    for (int k = 0; k < K; ++k) std::printf("%s[%.9g, %.9g]", k ? ", " : "", cross_cat.kinds[k].model.clustering.alpha, cross_cat.kinds[k].model.clustering.d);
    std::printf("], \"features\": [");
    std::vector<std::vector<float>> hypers(F);
    for (int k = 0; k < K; ++k) { print_hypers_fun fun = {cross_cat.kinds[k], hypers}; loom::for_each_feature_type(fun); }
    for (int f = 0; f < F; ++f) {
        std::printf("%s[", f ? ", " : "");
        for (size_t j = 0; j < hypers[f].size(); ++j) std::printf("%s%.9g", j ? "," : "", hypers[f][j]);
        std::printf("]");
    }
    std::printf("], \"draws\": %zu}\n", distributions::scores_log().size());
    return 0;
}
