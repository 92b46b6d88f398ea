/*
 * oracle_structure.c -- CPU restatement of loom's kind kernel and hyper kernel
 * (float64).  TEST INFRASTRUCTURE ONLY; PARITY UNPINNED vs the reference
 * binary (see loom_oracle.h), except oracle_block_py, which reproduces the
 * reference's own compiled kind_proposer.cc draw for draw
 * (tests/test_reference_pins.py).  Control flow follows:
 *   src/kind_kernel.hpp:193-217       add_to_kind_proposer (bulk form)
 *   src/kind_proposer.cc:131-362      BlockPitmanYorSampler, infer_assignments
 *   src/kind_kernel.cc:83-255         move_features, init_featureless_kinds
 *   src/product_mixture.cc:886-915    move_feature_to
 *   src/hyper_kernel.cc:40-235        HyperKernel::run
 *   src/infer_grid.hpp:45-125         InferShared, sample_clustering_posterior
 *   src/hyper_prior.hpp:38-146        grid visitors
 */
#ifdef _OPENMP
#include <omp.h>
#endif
#include "oracle_internal.h"

void oracle_free_proposer(engine_t *e) {
    if (e->prop_ints) for (int k = 0; k < e->prop_K; ++k) free(e->prop_ints[k]);
    if (e->prop_flts) for (int k = 0; k < e->prop_K; ++k) free(e->prop_flts[k]);
    free(e->prop_ints); free(e->prop_flts); free(e->prop_G);
    free(e->prop_int_off); free(e->prop_flt_off); free(e->prop_scores);
    e->prop_ints = NULL; e->prop_flts = NULL; e->prop_G = NULL;
    e->prop_int_off = e->prop_flt_off = NULL; e->prop_scores = NULL; e->prop_K = 0;
}

/* Bulk KindKernel::add_to_kind_proposer over rows [b, en): every assigned row goes, with ALL
 * its features, into group g_k(row) of every kind.  Float statistics are kept as mergeable
 * sums (sum x, sum x^2, sum log x!) until proposer_finish, so row shards can be all-reduced. */
int lo_proposer_accumulate(lb_handle h, uint64_t b, uint64_t en) {
    engine_t *e = h;
    if (!e->R) return lo_fail("proposer_accumulate before set_rows");
    if (b > en || en > e->R) return lo_fail("bad row range");
    oracle_free_proposer(e);
    const int F = e->F, K = e->K;
    e->prop_K = K;
    e->prop_int_off = calloc(F, sizeof(int));
    e->prop_flt_off = calloc(F, sizeof(int));
    e->prop_int_rows = e->prop_flt_rows = 0;
    for (int f = 0; f < F; ++f) {
        e->prop_int_off[f] = e->prop_int_rows; e->prop_int_rows += spec_int_rows(&e->specs[f]);
        e->prop_flt_off[f] = e->prop_flt_rows; e->prop_flt_rows += spec_flt_rows(&e->specs[f]);
    }
    e->prop_ints = calloc(K, sizeof(void *));
    e->prop_flts = calloc(K, sizeof(void *));
    e->prop_G = calloc(K, sizeof(int));
    e->prop_scores = calloc((size_t)F * K, sizeof(float));
#pragma omp parallel for schedule(dynamic, 1)
    for (int kid = 0; kid < K; ++kid) {
        const kind_t *k = &e->kinds[kid];
        const int G = k->G;
        e->prop_G[kid] = G;
        uint32_t *ints = calloc((size_t)e->prop_int_rows * G + 1, 4);
        double *flts = calloc((size_t)e->prop_flt_rows * G + 1, 8);
        e->prop_ints[kid] = ints; e->prop_flts[kid] = flts;
        for (uint64_t r = b; r < en; ++r) {
            if (k->assign[r] == LB_UNASSIGNED) continue;
            const uint32_t g = k->g2p[k->assign[r]];
            for (int f = 0; f < F; ++f) {
                if (!cell_observed(e, f, r)) continue;
                const lb_feature_spec *s = &e->specs[f];
                uint32_t *in = ints + (size_t)e->prop_int_off[f] * G + g;
                double *fl = flts + (size_t)e->prop_flt_off[f] * G + g;
                const uint32_t raw = cell_raw(e, f, r);
                if (s->type == LB_NICH) {
                    float xf; memcpy(&xf, &raw, 4);
                    in[0] += 1; fl[0] += (double)xf; fl[G] += (double)xf * (double)xf;
                } else {
                    om_add_value(s, in, fl, G, raw, +1);
                }
            }
        }
    }
    return 0;
}

int lo_proposer_tables(lb_handle h, int32_t kind, uint32_t **ints, uint64_t *n_ints, double **flts,
                       uint64_t *n_flts) {
    engine_t *e = h;
    if (!e->prop_ints || kind < 0 || kind >= e->prop_K) return lo_fail("proposer_tables: call proposer_accumulate first");
    *ints = e->prop_ints[kind]; *n_ints = (uint64_t)e->prop_int_rows * e->prop_G[kind];
    *flts = e->prop_flts[kind]; *n_flts = (uint64_t)e->prop_flt_rows * e->prop_G[kind];
    return 0;
}

/* sums -> (mean, count_times_variance), then the KindProposer score phase
 * (src/kind_proposer.cc:339-348, OMP over features). */
int lo_proposer_finish(lb_handle h, float *out) {
    engine_t *e = h;
    if (!e->prop_ints) return lo_fail("proposer_finish: call proposer_accumulate first");
    const int F = e->F, K = e->prop_K;
    for (int kid = 0; kid < K; ++kid) {
        const int G = e->prop_G[kid];
        for (int f = 0; f < F; ++f) {
            if (e->specs[f].type != LB_NICH) continue;
            for (int g = 0; g < G; ++g) {
                const double n = e->prop_ints[kid][(size_t)e->prop_int_off[f] * G + g];
                double *fl = e->prop_flts[kid] + (size_t)e->prop_flt_off[f] * G + g;
                if (n == 0) { fl[0] = 0; fl[G] = 0; }
                else {
                    const double mean = fl[0] / n;
                    fl[0] = mean;
                    fl[G] = fmax(fl[G] - n * mean * mean, 0.0);
                }
            }
        }
    }
#pragma omp parallel for schedule(dynamic, 1)
    for (int f = 0; f < F; ++f) {
        for (int kid = 0; kid < K; ++kid) {
            const int G = e->prop_G[kid];
            double score = 0;
            for (int g = 0; g < G; ++g)
                score += om_score_data(&e->specs[f], e->aux,
                                       e->prop_ints[kid] + (size_t)e->prop_int_off[f] * G + g,
                                       e->prop_flts[kid] + (size_t)e->prop_flt_off[f] * G + g, G);
            e->prop_scores[(size_t)f * K + kid] = (float)score;
        }
    }
    if (out) memcpy(out, e->prop_scores, sizeof(float) * (size_t)F * K);
    return 0;
}

int lo_kind_proposer_score(lb_handle h, float *out) {
    engine_t *e = h;
    int rc = lo_proposer_accumulate(h, 0, e->R);
    return rc ? rc : lo_proposer_finish(h, out);
}

int lo_set_proposer_scores(lb_handle h, const float *scores) {
    engine_t *e = h;
    if (!e->prop_scores || e->prop_K != e->K) return lo_fail("set_proposer_scores: call kind_proposer_score first");
    memcpy(e->prop_scores, scores, sizeof(float) * (size_t)e->F * e->K);
    return 0;
}

/* ---- Block Pitman-Yor sampler (src/kind_proposer.cc:131-300) ------------ */
static double likelihood_empty(double alpha, double d, int K, int n_empty) {
    return n_empty ? (alpha + d * (K - n_empty)) / n_empty : 0.0;
}

void oracle_block_py(double alpha, double d, int F, int K, const double *lik /*[F][K]*/,
                     uint32_t *assign, int iterations, uint64_t seed) {
    uint32_t *counts = calloc(K, 4);
    double *prior = calloc(K, 8), *post = calloc(K, 8);
    int n_empty = 0;
    for (int f = 0; f < F; ++f) counts[assign[f]]++;
    for (int k = 0; k < K; ++k) if (!counts[k]) ++n_empty;
    for (int k = 0; k < K; ++k)
        prior[k] = counts[k] ? counts[k] - d : likelihood_empty(alpha, d, K, n_empty);
    for (int it = 0; it < iterations; ++it) {
        for (int f = 0; f < F; ++f) {
            int k = (int)assign[f];
            if (--counts[k] == 0) { /* add_empty_kind */
                ++n_empty;
                const double le = likelihood_empty(alpha, d, K, n_empty);
                for (int j = 0; j < K; ++j) if (!counts[j]) prior[j] = le;
            } else {
                prior[k] = counts[k] - d;
            }
            double total = 0;
            for (int j = 0; j < K; ++j) total += post[j] = prior[j] * lik[(size_t)f * K + j];
            const double u = lo_uniform(seed, 1, (uint64_t)it * F + f, 0);
            k = om_sample_from_likelihoods(u, post, K, total);
            assign[f] = (uint32_t)k;
            if (counts[k]++ == 0) { /* remove_empty_kind */
                --n_empty;
                const double le = likelihood_empty(alpha, d, K, n_empty);
                for (int j = 0; j < K; ++j) if (!counts[j]) prior[j] = le;
            }
            prior[k] = counts[k] - d;
        }
    }
    free(counts); free(prior); free(post);
}

/* rebuild a kind's per-feature tables for a new feature list; unmoved features
 * keep their statistics, moved ones take the proposer's (move_feature_to). */
static void kind_relayout(engine_t *e, int kid, const int *fids, int nf) {
    kind_t *k = &e->kinds[kid];
    kind_t old = *k;
    const int G = old.G;
    k->nf = nf;
    k->fids = calloc(nf ? nf : 1, sizeof(int));
    k->int_off = calloc(nf ? nf : 1, sizeof(int));
    k->flt_off = calloc(nf ? nf : 1, sizeof(int));
    k->cache_off = calloc(nf ? nf : 1, sizeof(int));
    k->n_int_rows = k->n_flt_rows = k->n_cache_rows = 0;
    for (int j = 0; j < nf; ++j) {
        const lb_feature_spec *s = &e->specs[fids[j]];
        k->fids[j] = fids[j];
        k->int_off[j] = k->n_int_rows; k->n_int_rows += spec_int_rows(s);
        k->flt_off[j] = k->n_flt_rows; k->n_flt_rows += spec_flt_rows(s);
        k->cache_off[j] = k->n_cache_rows; k->n_cache_rows += spec_cache_rows(s);
    }
    const int st = old.Gcap;
    k->ints = calloc((size_t)k->n_int_rows * st + 1, 4);
    k->flts = calloc((size_t)k->n_flt_rows * st + 1, 8);
    k->cache = calloc((size_t)k->n_cache_rows * st + 1, 8);
    for (int j = 0; j < nf; ++j) {
        const int f = fids[j];
        const lb_feature_spec *s = &e->specs[f];
        int oj = -1;
        for (int q = 0; q < old.nf; ++q) if (old.fids[q] == f) oj = q;
        for (int row = 0; row < spec_int_rows(s); ++row)
            for (int g = 0; g < G; ++g)
                k->ints[(size_t)(k->int_off[j] + row) * st + g] =
                    oj >= 0 ? old.ints[(size_t)(old.int_off[oj] + row) * st + g]
                            : e->prop_ints[kid][(size_t)(e->prop_int_off[f] + row) * G + g];
        for (int row = 0; row < spec_flt_rows(s); ++row)
            for (int g = 0; g < G; ++g)
                k->flts[(size_t)(k->flt_off[j] + row) * st + g] =
                    oj >= 0 ? old.flts[(size_t)(old.flt_off[oj] + row) * st + g]
                            : e->prop_flts[kid][(size_t)(e->prop_flt_off[f] + row) * G + g];
    }
    free(old.fids); free(old.int_off); free(old.flt_off); free(old.cache_off);
    free(old.ints); free(old.flts); free(old.cache);
    kind_refresh_all(e, k);
}

/* KindKernel::try_run minus init_featureless_kinds (src/kind_kernel.cc:118-157). */
int lo_kind_run(lb_handle h, int32_t iterations, uint64_t seed, uint32_t *new_f2k,
                int32_t *n_changed) {
    engine_t *e = h;
    if (iterations <= 0) return lo_fail("kind_run: iterations must be > 0");
    if (!e->prop_scores || e->prop_K != e->K)
        return lo_fail("kind_run: call kind_proposer_score first");
    const int F = e->F, K = e->K;
    for (int k = 0; k < K; ++k)
        if (e->prop_G[k] != e->kinds[k].G) return lo_fail("kind_run: stale proposer tables");
    double *lik = calloc((size_t)F * K, 8);
    for (int f = 0; f < F; ++f) { /* scores_to_likelihoods (kind_proposer.cc:347) */
        double m = -INFINITY;
        for (int k = 0; k < K; ++k) if (e->prop_scores[(size_t)f * K + k] > m) m = e->prop_scores[(size_t)f * K + k];
        for (int k = 0; k < K; ++k) lik[(size_t)f * K + k] = exp((double)e->prop_scores[(size_t)f * K + k] - m);
    }
    uint32_t *assign = calloc(F, 4);
    memcpy(assign, e->f2k, 4u * F);
    oracle_block_py(e->topo_alpha, e->topo_d, F, K, lik, assign, iterations, seed);
    free(lik);
    int changed = 0;
    for (int f = 0; f < F; ++f) if (assign[f] != e->f2k[f]) ++changed;
    int *fids = calloc(F, sizeof(int));
    for (int k = 0; k < K; ++k) {
        int nf = 0, same = 1;
        for (int f = 0; f < F; ++f) if (assign[f] == (uint32_t)k) fids[nf++] = f;
        if (nf != e->kinds[k].nf) same = 0;
        for (int j = 0; same && j < nf; ++j) if (fids[j] != e->kinds[k].fids[j]) same = 0;
        if (!same) kind_relayout(e, k, fids, nf);
    }
    free(fids);
    memcpy(e->f2k, assign, 4u * F);
    if (new_f2k) memcpy(new_f2k, assign, 4u * F);
    if (n_changed) *n_changed = changed;
    free(assign);
    oracle_free_proposer(e);
    return 0;
}

/* the CUDA library's overlap hook; the checker seats when asked (same result) */
int lo_prepare_featureless_kinds(lb_handle h, int32_t count, uint64_t seed) { (void)h; (void)count; (void)seed; return 0; }

/* KindKernel::init_featureless_kinds (src/kind_kernel.cc:159-227). */
int lo_init_featureless_kinds(lb_handle h, int32_t count, uint64_t seed) {
    engine_t *e = h;
    oracle_free_proposer(e);
    for (int i = e->K - 1; i >= 0; --i) {
        if (e->kinds[i].nf) continue;
        if (e->K == 1) return lo_fail("cannot remove the only kind");
        kind_free(&e->kinds[i]); /* remove_featureless_kind: packed_remove */
        const int last = e->K - 1;
        if (i != last) {
            e->kinds[i] = e->kinds[last];
            memset(&e->kinds[last], 0, sizeof(kind_t));
            for (int j = 0; j < e->kinds[i].nf; ++j) e->f2k[e->kinds[i].fids[j]] = (uint32_t)i;
        }
        e->K = last;
    }
    for (int c = 0; c < count; ++c) {
        const int kid = e->K;
        e->kinds = realloc(e->kinds, sizeof(kind_t) * (size_t)(kid + 1));
        kind_t *k = &e->kinds[kid];
        memset(k, 0, sizeof *k);
        e->K = kid + 1;
        double alpha = e->kinds[0].alpha, d = e->kinds[0].d;
        if (e->hp.n_clustering) { /* sample_clustering_prior (infer_grid.hpp:127-139) */
            const double u = lo_uniform(seed, 3, 0xFFFFFFFFFFFFFFFFull, (uint32_t)kid);
            int i = (int)(u * e->hp.n_clustering);
            if (i >= e->hp.n_clustering) i = e->hp.n_clustering - 1;
            alpha = e->hp.clustering[2 * i]; d = e->hp.clustering[2 * i + 1];
        }
        /* PitmanYor::sample_assignments over all assigned rows (kind_kernel.cc:172-187) */
        uint32_t *pick = calloc(e->R ? e->R : 1, 4);
        uint32_t *counts = NULL;
        double *p = NULL;
        int m = 0, cap = 0;
        uint64_t seated = 0;
        const kind_t *k0 = &e->kinds[0];
        for (uint64_t r = 0; r < e->R; ++r) {
            pick[r] = LB_UNASSIGNED;
            if (k0->assign[r] == LB_UNASSIGNED) continue;
            if (m + 1 > cap) {
                cap = 2 * cap + 16;
                counts = realloc(counts, 4u * (size_t)cap);
                p = realloc(p, 8u * (size_t)cap);
            }
            /* weights counts[g] - d for the occupied tables and alpha + d m for a new one; their sum is the number of
             * seated rows + alpha in closed form, so the scan stops at the pick instead of summing all m weights first */
            const double total = (double)seated + alpha;
            const double u = lo_uniform(seed, 3, r, (uint32_t)kid);
            const double t = u * total;
            double c = 0;
            int g = m;                              /* nothing reached: the new table (its weight is always > 0) */
            for (int i = 0; i < m; ++i) {
                c += (double)counts[i] - d;
                if (c > t) { g = i; break; }
            }
            if (g == m) counts[m++] = 0;
            counts[g]++;
            ++seated;
            pick[r] = (uint32_t)g;
        }
        const int G = m + e->empty_group_count;
        uint32_t *c2 = calloc(G, 4);
        if (m) memcpy(c2, counts, 4u * m);
        kind_init(e, k, NULL, 0, alpha, d, c2, G);
        k->assign = pick; /* global == packed right after init */
        free(c2); free(counts); free(p);
    }
    return 0;
}

/* ---- hyper kernel -------------------------------------------------------- */
int lo_set_hyper_prior(lb_handle h, const lb_hyper_prior *pr) {
    engine_t *e = h;
    free(e->hp_store);
    memset(&e->hp, 0, sizeof e->hp);
    e->hp_store = NULL;
    if (!pr) return 0;
    size_t n = 2 * (size_t)(pr->n_topology + pr->n_clustering) + pr->n_bb_alpha + pr->n_bb_beta +
               pr->n_dd_alpha + pr->n_dpd_gamma + pr->n_dpd_alpha + pr->n_gp_alpha +
               pr->n_gp_inv_beta + pr->n_nich_mu + pr->n_nich_kappa + pr->n_nich_sigmasq +
               pr->n_nich_nu;
    float *s = e->hp_store = calloc(n + 1, 4);
#define COPY(field, cnt, mult) \
    do { e->hp.field = s; e->hp.cnt = pr->cnt; \
         if (pr->cnt) { memcpy(s, pr->field, 4u * (mult) * pr->cnt); } \
         s += (mult) * pr->cnt; } while (0)
    COPY(topology, n_topology, 2); COPY(clustering, n_clustering, 2);
    COPY(bb_alpha, n_bb_alpha, 1); COPY(bb_beta, n_bb_beta, 1);
    COPY(dd_alpha, n_dd_alpha, 1);
    COPY(dpd_gamma, n_dpd_gamma, 1); COPY(dpd_alpha, n_dpd_alpha, 1);
    COPY(gp_alpha, n_gp_alpha, 1); COPY(gp_inv_beta, n_gp_inv_beta, 1);
    COPY(nich_mu, n_nich_mu, 1); COPY(nich_kappa, n_nich_kappa, 1);
    COPY(nich_sigmasq, n_nich_sigmasq, 1); COPY(nich_nu, n_nich_nu, 1);
#undef COPY
    return 0;
}

/* sample_clustering_posterior (src/infer_grid.hpp:101-125) */
static void sample_py_posterior(const float *grid, int n, const uint32_t *counts, int nc,
                                uint64_t seed, uint64_t task, double *alpha, double *d) {
    if (n <= 0) return;
    int i = 0;
    if (n > 1) {
        double *scores = calloc(n, 8);
        for (int j = 0; j < n; ++j)
            scores[j] = om_py_score_counts(grid[2 * j], grid[2 * j + 1], counts, nc);
        i = om_sample_from_scores(lo_uniform(seed, 2, task, 0), scores, n);
        free(scores);
    }
    *alpha = grid[2 * i]; *d = grid[2 * i + 1];
}

/* InferShared::done (src/infer_grid.hpp:73-89) for one scalar coordinate. */
static void grid_coordinate(engine_t *e, kind_t *k, int j, float *coord, const float *grid, int n,
                            uint64_t seed, uint64_t task, uint32_t *draw) {
    if (n <= 0) return;
    if (n == 1) { *coord = grid[0]; return; }
    const lb_feature_spec *s = &e->specs[k->fids[j]];
    double *scores = calloc(n, 8);
    for (int i = 0; i < n; ++i) {
        *coord = grid[i];
        double sc = 0;
        for (int g = 0; g < k->G; ++g) /* mixture.score_data_grid: sum over groups */
            sc += om_score_data(s, e->aux, k->ints + (size_t)k->int_off[j] * k->Gcap + g,
                                k->flts + (size_t)k->flt_off[j] * k->Gcap + g, k->Gcap);
        scores[i] = sc;
    }
    const int pick = om_sample_from_scores(lo_uniform(seed, 2, task, (*draw)++), scores, n);
    *coord = grid[pick];
    free(scores);
}

/* Gamma(shape, 1) by Marsaglia-Tsang from the task's uniform stream. */
static double gamma_from_uniforms(double shape, uint64_t seed, uint64_t task, uint32_t *draw) {
    if (shape < 1.0) {
        const double u = lo_uniform(seed, 2, task, (*draw)++);
        return gamma_from_uniforms(shape + 1.0, seed, task, draw) * pow(u, 1.0 / shape);
    }
    const double d = shape - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
    for (;;) {
        const double u1 = lo_uniform(seed, 2, task, (*draw)++), u2 = lo_uniform(seed, 2, task, (*draw)++);
        const double z = sqrt(-2.0 * log(1.0 - u1)) * cos(6.283185307179586 * u2);
        double v = 1.0 + c * z;
        if (v <= 0) continue;
        v = v * v * v;
        const double u = lo_uniform(seed, 2, task, (*draw)++);
        if (log(1.0 - u) < 0.5 * z * z + d - d * v + d * log(v)) return d * v;
    }
}

#define DPD_MIN_BETA 1e-6 /* distributions' DirichletProcessDiscrete::MIN_BETA() is not vendored: stand-in */

/* The DPD branch of infer_feature_hypers_fun (src/hyper_kernel.cc:90-181):
 *  - auxiliary table counts per value (:101-122).  The reference samples each cell's count m from a
 *    row of log Stirling numbers, P(m) ~ |s(n,m)| (alpha beta_v)^m; the number of tables of a CRP with
 *    concentration alpha beta_v after n customers has exactly that law and needs no table:
 *    m = 1 + sum_{i=1}^{n-1} Bernoulli(a / (a + i));
 *  - only when every dictionary value is observed (:124): gamma by grid Gibbs on
 *    V log gamma + lgamma(gamma) - lgamma(gamma + sum aux) (:127-141); betas, beta0 ~
 *    Dirichlet(aux, gamma) floored at MIN_BETA (:143-167); alpha by grid Gibbs on the data score of
 *    every group (:169-176).  The last two only when the alpha grid is non-empty, as in the reference. */
static void dpd_hyper(engine_t *e, kind_t *k, int j, lb_feature_spec *s, uint64_t seed, uint64_t task) {
    const int dim = s->dim, G = k->G;
    const size_t st = k->Gcap;
    const uint32_t *tab = k->ints + (size_t)k->int_off[j] * st;
    float *betas = e->aux + s->aux_off;
    uint32_t draw = 0;
    if (dim <= 0) return;
    uint64_t *aux = calloc(dim, sizeof(uint64_t));
    for (int g = 0; g < G; ++g)
        for (int v = 0; v < dim; ++v) {
            const uint32_t n = tab[(size_t)v * st + g];
            if (!n) continue;
            const double a = (double)s->hp[1] * (double)betas[v];
            uint64_t m = 1;
            for (uint32_t i = 1; i < n; ++i)
                if (lo_uniform(seed, 2, task, draw++) * (a + (double)i) < a) ++m;
            aux[v] += m;
        }
    int seen = 0;
    uint64_t aux_total = 0;
    for (int v = 0; v < dim; ++v) { seen += aux[v] > 0; aux_total += aux[v]; }
    if (seen == dim) {
        if (e->hp.n_dpd_gamma > 0) {
            const int n = e->hp.n_dpd_gamma;
            double *scores = calloc(n, sizeof(double));
            for (int i = 0; i < n; ++i) {
                const double gm = e->hp.dpd_gamma[i];
                scores[i] = dim * log(gm) + lgamma(gm) - lgamma(gm + (double)aux_total);
            }
            s->hp[0] = e->hp.dpd_gamma[om_sample_from_scores(lo_uniform(seed, 2, task, draw++), scores, n)];
            free(scores);
        }
        if (e->hp.n_dpd_alpha > 0) {
            double *x = calloc(dim + 1, sizeof(double)), total = 0, floored = 0;
            for (int i = 0; i <= dim; ++i) {
                x[i] = gamma_from_uniforms(i < dim ? (double)aux[i] : (double)s->hp[0], seed, task, &draw);
                total += x[i];
            }
            for (int i = 0; i <= dim; ++i) { x[i] = fmax(x[i] / total, DPD_MIN_BETA); floored += x[i]; }
            for (int v = 0; v < dim; ++v) betas[v] = (float)(x[v] / floored);
            s->hp[2] = (float)(x[dim] / floored);
            free(x);
            const int n = e->hp.n_dpd_alpha;
            double *scores = calloc(n, sizeof(double));
            for (int i = 0; i < n; ++i) {
                const double al = e->hp.dpd_alpha[i];
                double sc = 0;
                for (int g = 0; g < G; ++g) {
                    const uint32_t total_g = tab[(size_t)dim * st + g];
                    if (!total_g) continue;
                    for (int v = 0; v < dim; ++v) {
                        const uint32_t c = tab[(size_t)v * st + g];
                        if (c) sc += lgamma(al * (double)betas[v] + (double)c) - lgamma(al * (double)betas[v]);
                    }
                    sc += lgamma(al) - lgamma(al + (double)total_g);
                }
                scores[i] = sc;
            }
            s->hp[1] = e->hp.dpd_alpha[om_sample_from_scores(lo_uniform(seed, 2, task, draw++), scores, n)];
            free(scores);
        }
    }
    free(aux);
}

/* HyperKernel::run (src/hyper_kernel.cc:195-235). */
int lo_hyper_run(lb_handle h, uint64_t seed) { return lo_hyper_run_tasks(h, seed, NULL); }

int lo_hyper_run_tasks(lb_handle h, uint64_t seed, const uint8_t *task_mask) {
    engine_t *e = h;
    const int K = e->K, F = e->F;
    /* task 0: topology */
    if (e->hp.n_topology && (!task_mask || task_mask[0])) {
        uint32_t *fc = calloc(K, 4);
        for (int k = 0; k < K; ++k) fc[k] = (uint32_t)e->kinds[k].nf;
        sample_py_posterior(e->hp.topology, e->hp.n_topology, fc, K, seed, 0, &e->topo_alpha,
                            &e->topo_d);
        free(fc);
    }
#pragma omp parallel for schedule(dynamic, 1)
    for (int task = 1; task < 1 + K + F; ++task) {
        if (task_mask && !task_mask[task]) continue;
        if (task < 1 + K) {
            kind_t *k = &e->kinds[task - 1];
            if (e->hp.n_clustering) {
                sample_py_posterior(e->hp.clustering, e->hp.n_clustering, k->counts, k->G, seed,
                                    (uint64_t)task, &k->alpha, &k->d);
                kind_refresh_clustering(k);
            }
        } else {
            const int f = task - 1 - K;
            kind_t *k = &e->kinds[e->f2k[f]];
            int j = -1;
            for (int q = 0; q < k->nf; ++q) if (k->fids[q] == f) j = q;
            lb_feature_spec *s = &e->specs[f];
            uint32_t draw = 0;
            switch (s->type) {
            case LB_BB:
                grid_coordinate(e, k, j, &s->hp[0], e->hp.bb_alpha, e->hp.n_bb_alpha, seed, task, &draw);
                grid_coordinate(e, k, j, &s->hp[1], e->hp.bb_beta, e->hp.n_bb_beta, seed, task, &draw);
                break;
            case LB_DD:
                for (int v = 0; v < s->dim; ++v)
                    grid_coordinate(e, k, j, &e->aux[s->aux_off + v], e->hp.dd_alpha,
                                    e->hp.n_dd_alpha, seed, task, &draw);
                break;
            case LB_GP:
                grid_coordinate(e, k, j, &s->hp[0], e->hp.gp_alpha, e->hp.n_gp_alpha, seed, task, &draw);
                grid_coordinate(e, k, j, &s->hp[1], e->hp.gp_inv_beta, e->hp.n_gp_inv_beta, seed, task, &draw);
                break;
            case LB_NICH:
                grid_coordinate(e, k, j, &s->hp[0], e->hp.nich_mu, e->hp.n_nich_mu, seed, task, &draw);
                grid_coordinate(e, k, j, &s->hp[1], e->hp.nich_kappa, e->hp.n_nich_kappa, seed, task, &draw);
                grid_coordinate(e, k, j, &s->hp[2], e->hp.nich_sigmasq, e->hp.n_nich_sigmasq, seed, task, &draw);
                grid_coordinate(e, k, j, &s->hp[3], e->hp.nich_nu, e->hp.n_nich_nu, seed, task, &draw);
                break;
            case LB_DPD:
                dpd_hyper(e, k, j, s, seed, (uint64_t)task);
                break;
            default:
                break;
            }
        }
    }
    for (int k = 0; k < K; ++k) kind_refresh_all(e, &e->kinds[k]); /* mixture.init(shared) */
    return 0;
}

/* the gather step of the sharded hyper kernel: install hypers chosen elsewhere */
int lo_set_hypers(lb_handle h, const lb_feature_spec *specs, const float *aux, const float *kind_py,
                  const float *topology_py) {
    engine_t *e = h;
    if (!e->F) return lo_fail("set_hypers before set_model");
    if (specs)
        for (int f = 0; f < e->F; ++f) {
            if (specs[f].type != e->specs[f].type || specs[f].dim != e->specs[f].dim || specs[f].aux_off != e->specs[f].aux_off)
                return lo_fail("set_hypers: feature %d does not match the model", f);
            memcpy(e->specs[f].hp, specs[f].hp, sizeof(specs[f].hp));
        }
    if (aux && e->n_aux) memcpy(e->aux, aux, 4u * (size_t)e->n_aux);
    if (kind_py)
        for (int k = 0; k < e->K; ++k) {
            e->kinds[k].alpha = kind_py[2 * k];
            e->kinds[k].d = kind_py[2 * k + 1];
            kind_refresh_clustering(&e->kinds[k]);
        }
    if (topology_py) { e->topo_alpha = topology_py[0]; e->topo_d = topology_py[1]; }
    for (int k = 0; k < e->K; ++k) kind_refresh_all(e, &e->kinds[k]);
    return 0;
}
