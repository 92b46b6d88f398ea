/*
 * loom_oracle.c -- CPU restatement of loom's cat-kernel hot path (float64).
 * TEST INFRASTRUCTURE ONLY; PARITY UNPINNED vs the reference binary (see
 * loom_oracle.h).  Control flow follows, file by file:
 *   src/cat_kernel.hpp:199-299        process_add_task / process_remove_task
 *   src/product_mixture.cc:165-239    add_value / remove_value (+ group add/remove)
 *   src/product_mixture.cc:421-434    score_value
 *   src/product_mixture.cc:581-592    score_data
 *   src/cross_cat.cc:238-250          CrossCat::score_data
 *   src/assignments.hpp:56-74         FIFO of global group ids
 * Packed<->global ids restate distributions::MixtureIdTracker.
 */
#include <stdarg.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "oracle_internal.h"

static char g_err[512]; /* shared: errors are raised inside OpenMP workers */
static int g_threads = 0;

int lo_fail(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return -1;
}
const char *lo_last_error(void) { return g_err; }
void lo_set_threads(int n) { g_threads = n; }
static int n_threads(void) {
#ifdef _OPENMP
    return g_threads > 0 ? g_threads : omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------ */
void kind_free(kind_t *k) {
    free(k->fids); free(k->counts); free(k->p2g); free(k->g2p);
    free(k->int_off); free(k->flt_off); free(k->cache_off);
    free(k->ints); free(k->flts); free(k->cache); free(k->shifted); free(k->assign);
    memset(k, 0, sizeof *k);
}

static void *xcalloc(size_t n, size_t sz) {
    void *p = calloc(n ? n : 1, sz);
    if (!p) { fprintf(stderr, "oracle: out of memory\n"); abort(); }
    return p;
}

/* grow the column capacity of every per-group table */
void kind_ensure_cap(kind_t *k, int G) {
    if (G <= k->Gcap) return;
    int ncap = k->Gcap ? k->Gcap : 64;
    while (ncap < G) ncap *= 2;
    uint32_t *counts = xcalloc(ncap, 4), *p2g = xcalloc(ncap, 4);
    double *shifted = xcalloc(ncap, 8);
    uint32_t *ints = xcalloc((size_t)k->n_int_rows * ncap, 4);
    double *flts = xcalloc((size_t)k->n_flt_rows * ncap, 8);
    double *cache = xcalloc((size_t)k->n_cache_rows * ncap, 8);
    if (k->Gcap) {
        memcpy(counts, k->counts, 4u * k->G);
        memcpy(p2g, k->p2g, 4u * k->G);
        memcpy(shifted, k->shifted, 8u * k->G);
        for (int r = 0; r < k->n_int_rows; ++r)
            memcpy(ints + (size_t)r * ncap, k->ints + (size_t)r * k->Gcap, 4u * k->G);
        for (int r = 0; r < k->n_flt_rows; ++r)
            memcpy(flts + (size_t)r * ncap, k->flts + (size_t)r * k->Gcap, 8u * k->G);
        for (int r = 0; r < k->n_cache_rows; ++r)
            memcpy(cache + (size_t)r * ncap, k->cache + (size_t)r * k->Gcap, 8u * k->G);
    }
    free(k->counts); free(k->p2g); free(k->shifted); free(k->ints); free(k->flts); free(k->cache);
    k->counts = counts; k->p2g = p2g; k->shifted = shifted;
    k->ints = ints; k->flts = flts; k->cache = cache; k->Gcap = ncap;
}

static void g2p_ensure(kind_t *k, uint32_t id) {
    if (id < k->g2p_cap) return;
    uint32_t ncap = k->g2p_cap ? k->g2p_cap : 256;
    while (ncap <= id) ncap *= 2;
    k->g2p = realloc(k->g2p, 4u * (size_t)ncap);
    for (uint32_t i = k->g2p_cap; i < ncap; ++i) k->g2p[i] = LB_UNASSIGNED;
    k->g2p_cap = ncap;
}

/* cached scorer entries of feature j (kind-local) for group g; `raw` < 0 means
 * refresh every value row (DD/DPD), else only the row of that value + shift. */
static void feature_refresh(engine_t *e, kind_t *k, int j, int g, int64_t raw) {
    const lb_feature_spec *s = &e->specs[k->fids[j]];
    const int st = k->Gcap;
    const uint32_t *in = k->ints + (size_t)k->int_off[j] * st + g;
    const double *fl = k->flts + (size_t)k->flt_off[j] * st + g;
    double *c = k->cache + (size_t)k->cache_off[j] * st + g;
    switch (s->type) {
    case LB_BB: {
        double a = s->hp[0], b = s->hp[1], h = in[0], t = in[st];
        c[0] = log((a + h) / (a + b + h + t));
        c[st] = log((b + t) / (a + b + h + t));
        break;
    }
    case LB_DD: case LB_DPD: {
        const float *ax = e->aux + s->aux_off;
        double scale = s->type == LB_DPD ? s->hp[1] : 1.0, tot;
        if (s->type == LB_DPD) tot = s->hp[1];
        else { tot = 0; for (int v = 0; v < s->dim; ++v) tot += ax[v]; }
        if (raw < 0) {
            for (int v = 0; v < s->dim; ++v)
                c[(size_t)v * st] = log(scale * ax[v] + in[(size_t)v * st]);
        } else {
            c[(size_t)raw * st] = log(scale * ax[raw] + in[(size_t)raw * st]);
        }
        c[(size_t)s->dim * st] = log(tot + in[(size_t)s->dim * st]);
        break;
    }
    case LB_GP: {
        /* (double) first: float + uint32 is a FLOAT sum in C, and an a rounded to float32 here against the exact a of the
         * scoring loop's lgamma(a + x) left an error of psi(a) * ulp(a) / 2 in every score of a large group whenever alpha
         * is not an integer (found by benchmarks/fuzz_reference_pins.py against the compiled reference) */
        double a = (double)s->hp[0] + in[st], b = (double)s->hp[1] + in[0];
        c[0] = -lgamma(a) + a * log(b / (b + 1.0));
        c[st] = -log(1.0 + b);
        break;
    }
    case LB_NICH: {
        double mu = s->hp[0], kappa = s->hp[1], sigmasq = s->hp[2], nu = s->hp[3];
        double n = in[0], mean = fl[0], ctv = fl[st];
        double kn = kappa + n, mun = (kappa * mu + n * mean) / kn, nun = nu + n;
        double dm = mu - mean;
        double sn = (nu * sigmasq + ctv + n * kappa * dm * dm / kn) / nun;
        double lambda = kn / ((kn + 1.0) * sn);
        c[0] = lgamma(0.5 * (nun + 1.0)) - lgamma(0.5 * nun) + 0.5 * log(lambda / (M_PI * nun));
        c[st] = 0.5 * (nun + 1.0);
        c[2 * (size_t)st] = lambda / nun;
        c[3 * (size_t)st] = mun;
        break;
    }
    }
}

void kind_refresh_group(engine_t *e, kind_t *k, int g) {
    k->shifted[g] = k->counts[g] ? log(k->counts[g] - k->d) : 0.0;
    for (int j = 0; j < k->nf; ++j) feature_refresh(e, k, j, g, -1);
}
void kind_refresh_clustering(kind_t *k) {
    for (int g = 0; g < k->G; ++g)
        k->shifted[g] = k->counts[g] ? log(k->counts[g] - k->d) : 0.0;
}
void kind_refresh_all(engine_t *e, kind_t *k) {
    for (int g = 0; g < k->G; ++g) kind_refresh_group(e, k, g);
}

/* ProductMixture_::init_unobserved (src/product_mixture.cc, called from
 * src/cross_cat.cc:50-60 and src/kind_kernel.cc:188): groups with the given
 * counts and empty feature statistics; id tracker reset to global == packed. */
void kind_init(engine_t *e, kind_t *k, const int *fids, int nf, double alpha, double d,
               const uint32_t *counts, int G) {
    uint32_t *assign = k->assign; /* keep the assignment column */
    k->assign = NULL;
    kind_free(k);
    k->assign = assign;
    k->nf = nf;
    k->fids = xcalloc(nf, sizeof(int));
    k->int_off = xcalloc(nf, sizeof(int));
    k->flt_off = xcalloc(nf, sizeof(int));
    k->cache_off = xcalloc(nf, sizeof(int));
    for (int j = 0; j < nf; ++j) {
        k->fids[j] = fids[j];
        const lb_feature_spec *s = &e->specs[fids[j]];
        k->int_off[j] = k->n_int_rows; k->n_int_rows += spec_int_rows(s);
        k->flt_off[j] = k->n_flt_rows; k->n_flt_rows += spec_flt_rows(s);
        k->cache_off[j] = k->n_cache_rows; k->n_cache_rows += spec_cache_rows(s);
    }
    k->alpha = alpha; k->d = d;
    kind_ensure_cap(k, G > 32 ? 2 * G : 64);
    k->G = G;
    k->sample_size = 0;
    for (int g = 0; g < G; ++g) {
        k->counts[g] = counts ? counts[g] : 0;
        k->sample_size += k->counts[g];
        k->p2g[g] = (uint32_t)g;
    }
    g2p_ensure(k, (uint32_t)G);
    for (uint32_t i = 0; i < k->g2p_cap; ++i) k->g2p[i] = i < (uint32_t)G ? i : LB_UNASSIGNED;
    k->next_global = (uint32_t)G;
    kind_refresh_all(e, k);
}

/* ------------------------------------------------------------------------ */
int lo_create(int device, lb_handle *out) {
    (void)device;
    engine_t *e = xcalloc(1, sizeof *e);
    e->empty_group_count = 1;
    *out = e;
    return 0;
}

static void engine_clear_model(engine_t *e) {
    for (int k = 0; k < e->K; ++k) kind_free(&e->kinds[k]);
    free(e->kinds); e->kinds = NULL; e->K = 0;
    free(e->specs); e->specs = NULL;
    free(e->aux); e->aux = NULL;
    free(e->f2k); e->f2k = NULL;
}
void oracle_free_proposer(engine_t *e);

void lo_destroy(lb_handle h) {
    engine_t *e = h;
    if (!e) return;
    oracle_free_proposer(e);
    engine_clear_model(e);
    if (!e->rows_borrowed) { free(e->cat8); free(e->wide); free(e->mask); }
    free(e->hp_store);
    free(e->differ_hist);
    free(e->diff_pos_ptr); free(e->diff_neg_ptr); free(e->diff_pos_feature); free(e->diff_pos_value); free(e->diff_neg_feature);
    free(e);
}

static void alloc_assign(engine_t *e, kind_t *k) {
    free(k->assign);
    k->assign = xcalloc(e->R, 4);
    for (uint64_t r = 0; r < e->R; ++r) k->assign[r] = LB_UNASSIGNED;
}

int lo_set_model(lb_handle h, int32_t F, const lb_feature_spec *specs, const float *aux,
                 int32_t n_aux, int32_t K, const uint32_t *f2k, const float *kind_py,
                 const float *topology_py, int32_t empty_group_count) {
    engine_t *e = h;
    if (F <= 0 || K <= 0) return lo_fail("set_model: need features and kinds");
    if (empty_group_count < 1) return lo_fail("empty_group_count must be >= 1"); /* cat_kernel.hpp:55 */
    int F8 = 0;
    for (int f = 0; f < F; ++f) {
        if (f && specs[f].type < specs[f - 1].type)
            return lo_fail("features must be sorted by type (bb<dd<dpd<gp<nich)");
        if (specs[f].type == LB_DD && (specs[f].dim < 1 || specs[f].dim > 256))
            return lo_fail("DD dim must be in [1,256] (product_model.cc:58-68)");
        if (specs[f].type == LB_BB || specs[f].type == LB_DD) ++F8;
        if (f2k[f] >= (uint32_t)K) return lo_fail("featureid_to_kindid out of range");
    }
    oracle_free_proposer(e);
    engine_clear_model(e);
    e->F = F; e->F8 = F8; e->FW = F - F8;
    e->specs = xcalloc(F, sizeof *specs);
    memcpy(e->specs, specs, sizeof(*specs) * F);
    e->aux = xcalloc(n_aux, 4); e->n_aux = n_aux;
    if (n_aux) memcpy(e->aux, aux, 4u * n_aux);
    e->f2k = xcalloc(F, 4);
    memcpy(e->f2k, f2k, 4u * F);
    e->K = K;
    e->kinds = xcalloc(K, sizeof(kind_t));
    e->topo_alpha = topology_py[0]; e->topo_d = topology_py[1];
    e->empty_group_count = empty_group_count;
    int *fids = xcalloc(F, sizeof(int));
    for (int k = 0; k < K; ++k) {
        int nf = 0;
        for (int f = 0; f < F; ++f) if (f2k[f] == (uint32_t)k) fids[nf++] = f;
        kind_init(e, &e->kinds[k], fids, nf, kind_py[2 * k], kind_py[2 * k + 1], NULL,
                  empty_group_count);
        if (e->R) alloc_assign(e, &e->kinds[k]);
    }
    free(fids);
    return 0;
}

int lo_set_rows(lb_handle h, uint64_t R, const uint8_t *cat8, const uint32_t *wide,
                const uint32_t *mask) {
    engine_t *e = h;
    if (!e->F) return lo_fail("set_rows before set_model");
    if (!e->rows_borrowed) { free(e->cat8); free(e->wide); free(e->mask); }
    e->rows_borrowed = 0;
    e->R = R; e->W = (R + 31) / 32;
    e->cat8 = xcalloc((size_t)e->F8 * R, 1);
    e->wide = xcalloc((size_t)e->FW * R, 4);
    e->mask = xcalloc((size_t)e->F * e->W, 4);
    if (e->F8) memcpy(e->cat8, cat8, (size_t)e->F8 * R);
    if (e->FW) memcpy(e->wide, wide, (size_t)e->FW * R * 4);
    memcpy(e->mask, mask, (size_t)e->F * e->W * 4);
    for (uint64_t r = 0; r < R; ++r)
        for (int f = 0; f < e->F8; ++f)
            if (cell_observed(e, f, r)) {
                const lb_feature_spec *s = &e->specs[f];
                uint32_t v = cell_raw(e, f, r);
                if ((s->type == LB_BB && v > 1) || (s->type == LB_DD && v >= (uint32_t)s->dim))
                    return lo_fail("row %llu feature %d value %u out of range",
                                   (unsigned long long)r, f, v);
            }
    for (int k = 0; k < e->K; ++k) alloc_assign(e, &e->kinds[k]);
    return 0;
}

/* the same block streamed in again (src/stream_interval.hpp:38-75): cells replaced, assignments kept */
int lo_reload_rows(lb_handle h, uint64_t R, const uint8_t *cat8, const uint32_t *wide, const uint32_t *mask) {
    engine_t *e = h;
    if (!e->F || !e->mask || e->rows_borrowed) return lo_fail("reload_rows: the engine holds no rows of its own");
    if (R != e->R) return lo_fail("reload_rows: %llu rows given, the block holds %llu", (unsigned long long)R, (unsigned long long)e->R);
    if (e->F8) memcpy(e->cat8, cat8, (size_t)e->F8 * R);
    if (e->FW) memcpy(e->wide, wide, (size_t)e->FW * R * 4);
    memcpy(e->mask, mask, (size_t)e->F * e->W * 4);
    return 0;
}

/* loom/tasks.py:173-194: every sample reads the same rows file */
int lo_share_rows(lb_handle h, lb_handle owner) {
    engine_t *e = h, *o = owner;
    if (!e || !o || e == o) return lo_fail("share_rows: need two distinct handles");
    if (!e->F) return lo_fail("set_rows before set_model");
    if (!o->R || !o->mask) return lo_fail("share_rows: the owner holds no rows");
    if (o->rows_borrowed) return lo_fail("share_rows: the owner itself borrows its rows");
    if (e->F != o->F || e->F8 != o->F8) return lo_fail("share_rows: schemas differ");
    for (int f = 0; f < e->F; ++f)
        if (e->specs[f].type != o->specs[f].type || e->specs[f].dim != o->specs[f].dim)
            return lo_fail("share_rows: feature %d differs between the two models", f);
    if (!e->rows_borrowed) { free(e->cat8); free(e->wide); free(e->mask); }
    e->R = o->R; e->W = o->W;
    e->cat8 = o->cat8; e->wide = o->wide; e->mask = o->mask;
    e->rows_borrowed = 1;
    for (int k = 0; k < e->K; ++k) alloc_assign(e, &e->kinds[k]);
    return 0;
}

/* Tare-compressed rows (src/differ.cc:248-298): reconstruct true row = tare + pos - neg. */
int lo_set_rows_diff(lb_handle h, uint64_t R, const uint8_t *tare_obs, const uint32_t *tare_val,
                     const uint8_t *row_has_tare, const uint64_t *pos_ptr, const uint32_t *pos_feature,
                     const uint32_t *pos_value, const uint64_t *neg_ptr, const uint32_t *neg_feature) {
    engine_t *e = h;
    if (!e->F) return lo_fail("set_rows_diff before set_model");
    const int F = e->F, F8 = e->F8;
    const uint64_t W = (R + 31) / 32;
    uint8_t *cat8 = xcalloc((size_t)F8 * R, 1);
    uint32_t *wide = xcalloc((size_t)(F - F8) * R, 4);
    uint32_t *mask = xcalloc((size_t)F * W, 4);
    int rc = 0;
    for (uint64_t r = 0; r < R && !rc; ++r) {
        if (!row_has_tare || row_has_tare[r])
            for (int f = 0; f < F; ++f)
                if (tare_obs[f]) {
                    if (f < F8) cat8[(size_t)f * R + r] = (uint8_t)tare_val[f];
                    else wide[(size_t)(f - F8) * R + r] = tare_val[f];
                    mask[(size_t)f * W + (r >> 5)] |= 1u << (r & 31);
                }
        for (uint64_t i = neg_ptr[r]; i < neg_ptr[r + 1]; ++i) {
            const uint32_t f = neg_feature[i];
            if (f >= (uint32_t)F || !((mask[(size_t)f * W + (r >> 5)] >> (r & 31)) & 1u)) { rc = 1; break; }
            mask[(size_t)f * W + (r >> 5)] &= ~(1u << (r & 31));
        }
        for (uint64_t i = pos_ptr[r]; i < pos_ptr[r + 1] && !rc; ++i) {
            const uint32_t f = pos_feature[i];
            if (f >= (uint32_t)F || ((mask[(size_t)f * W + (r >> 5)] >> (r & 31)) & 1u)) { rc = 1; break; }
            if (f < (uint32_t)F8) cat8[(size_t)f * R + r] = (uint8_t)pos_value[i];
            else wide[(size_t)(f - F8) * R + r] = pos_value[i];
            mask[(size_t)f * W + (r >> 5)] |= 1u << (r & 31);
        }
    }
    if (!rc) rc = lo_set_rows(h, R, cat8, wide, mask);
    else rc = lo_fail("set_rows_diff: neg removes a cell the tare does not hold, or pos overwrites a live cell");
    free(cat8); free(wide); free(mask);
    return rc;
}

/* ------------------------------------------------------------------------ */
/* ProductMixture_<true>::score_value (src/product_mixture.cc:421-434):
 * clustering prior, then one cached-scorer pass per observed feature. */
static void kind_score_row(const engine_t *e, const kind_t *k, uint64_t r, double *scores) {
    const int G = k->G, st = k->Gcap;
    int m = 0, ne = 0;
    for (int g = 0; g < G; ++g) { if (k->counts[g]) ++m; else ++ne; }
    const double base = -log((double)k->sample_size + k->alpha);
    const double sh_empty = ne ? log((k->alpha + k->d * m) / ne) : 0.0;
    for (int g = 0; g < G; ++g) scores[g] = (k->counts[g] ? k->shifted[g] : sh_empty) + base;
    for (int j = 0; j < k->nf; ++j) {
        const int f = k->fids[j];
        if (!cell_observed(e, f, r)) continue;
        const lb_feature_spec *s = &e->specs[f];
        const uint32_t raw = cell_raw(e, f, r);
        const double *c = k->cache + (size_t)k->cache_off[j] * st;
        switch (s->type) {
        case LB_BB: {
            const double *row = c + (raw ? 0 : st);
            for (int g = 0; g < G; ++g) scores[g] += row[g];
            break;
        }
        case LB_DD: case LB_DPD: {
            const double *row = c + (size_t)raw * st, *sh = c + (size_t)s->dim * st;
            for (int g = 0; g < G; ++g) scores[g] += row[g] - sh[g];
            break;
        }
        case LB_GP: {
            const uint32_t *sum = k->ints + (size_t)(k->int_off[j] + 1) * st;
            const double x = raw, lf = lgamma(x + 1.0), al = s->hp[0];
            for (int g = 0; g < G; ++g)
                scores[g] += c[g] + lgamma(al + sum[g] + x) - lf + c[st + g] * x;
            break;
        }
        case LB_NICH: {
            float xf; memcpy(&xf, &raw, 4);
            const double x = xf;
            for (int g = 0; g < G; ++g) {
                double dx = x - c[3 * (size_t)st + g];
                scores[g] += c[g] - c[st + g] * log1p(c[2 * (size_t)st + g] * dx * dx);
            }
            break;
        }
        }
    }
}

/* ProductMixture_::add_value (src/product_mixture.cc:165-184). */
static void kind_add_row(engine_t *e, kind_t *k, uint64_t r, int g) {
    const int st = k->Gcap;
    const int was_empty = k->counts[g] == 0;
    k->counts[g] += 1; k->sample_size += 1;
    k->shifted[g] = log(k->counts[g] - k->d);
    for (int j = 0; j < k->nf; ++j) {
        const int f = k->fids[j];
        if (!cell_observed(e, f, r)) continue;
        const uint32_t raw = cell_raw(e, f, r);
        om_add_value(&e->specs[f], k->ints + (size_t)k->int_off[j] * st + g,
                     k->flts + (size_t)k->flt_off[j] * st + g, st, raw, +1);
        feature_refresh(e, k, j, g, raw);
    }
    if (was_empty) { /* add_group (:178-183): append a fresh empty group */
        kind_ensure_cap(k, k->G + 1);
        const int ng = k->G, st2 = k->Gcap;
        k->counts[ng] = 0;
        for (int row = 0; row < k->n_int_rows; ++row) k->ints[(size_t)row * st2 + ng] = 0;
        for (int row = 0; row < k->n_flt_rows; ++row) k->flts[(size_t)row * st2 + ng] = 0;
        k->G = ng + 1;
        kind_refresh_group(e, k, ng);
        g2p_ensure(k, k->next_global);
        k->p2g[ng] = k->next_global;
        k->g2p[k->next_global] = (uint32_t)ng;
        k->next_global += 1;
    }
}

/* ProductMixture_::remove_value (src/product_mixture.cc:220-239). */
static void kind_remove_row(engine_t *e, kind_t *k, uint64_t r, int g) {
    const int st = k->Gcap;
    k->counts[g] -= 1; k->sample_size -= 1;
    for (int j = 0; j < k->nf; ++j) {
        const int f = k->fids[j];
        if (!cell_observed(e, f, r)) continue;
        const uint32_t raw = cell_raw(e, f, r);
        om_add_value(&e->specs[f], k->ints + (size_t)k->int_off[j] * st + g,
                     k->flts + (size_t)k->flt_off[j] * st + g, st, raw, -1);
        feature_refresh(e, k, j, g, raw);
    }
    if (k->counts[g] == 0) { /* remove_group (:233-238): swap-with-last */
        const int last = k->G - 1;
        k->g2p[k->p2g[g]] = LB_UNASSIGNED;
        if (g != last) {
            k->counts[g] = k->counts[last];
            k->shifted[g] = k->shifted[last];
            for (int row = 0; row < k->n_int_rows; ++row)
                k->ints[(size_t)row * st + g] = k->ints[(size_t)row * st + last];
            for (int row = 0; row < k->n_flt_rows; ++row)
                k->flts[(size_t)row * st + g] = k->flts[(size_t)row * st + last];
            for (int row = 0; row < k->n_cache_rows; ++row)
                k->cache[(size_t)row * st + g] = k->cache[(size_t)row * st + last];
            k->p2g[g] = k->p2g[last];
            k->g2p[k->p2g[g]] = (uint32_t)g;
        }
        k->G = last;
    } else {
        k->shifted[g] = log(k->counts[g] - k->d);
    }
}

/* CatKernel::process_add_task / process_remove_task for one kind over the
 * whole action list (src/cat_kernel.hpp:199-224, 280-299). */
static int kind_sweep(engine_t *e, int kid, const uint32_t *actions, uint64_t n, uint64_t seed,
                      uint64_t offset, const uint32_t *replay_picks, uint32_t *debug_picks,
                      float *debug_scores, int stride, double tol, const float *dev_scores,
                      int64_t *n_bad, double *max_err) {
    kind_t *k = &e->kinds[kid];
    const int K = e->K;
    double *scores = NULL;
    int cap = 0;
    for (uint64_t i = 0; i < n; ++i) {
        const uint64_t r = actions[i] >> 1;
        if (r >= e->R) return lo_fail("action %llu: row out of range", (unsigned long long)i);
        if (actions[i] & 1u) {
            if (k->assign[r] == LB_UNASSIGNED)
                return lo_fail("remove of unassigned row %llu", (unsigned long long)r);
            const uint32_t g = k->g2p[k->assign[r]];
            if (replay_picks && replay_picks[i * K + kid] != g) ++*n_bad;
            if (debug_picks) debug_picks[i * K + kid] = g;
            kind_remove_row(e, k, r, (int)g);
            k->assign[r] = LB_UNASSIGNED;
        } else {
            if (k->assign[r] != LB_UNASSIGNED)
                return lo_fail("duplicate row: %llu", (unsigned long long)r); /* cat_kernel.hpp:184 */
            if (k->G > cap) { cap = 2 * k->G + 64; scores = realloc(scores, 8u * (size_t)cap); }
            const int G = k->G;
            kind_score_row(e, k, r, scores);
            if (debug_scores)
                for (int g = 0; g < G && g < stride; ++g)
                    debug_scores[(i * K + kid) * (uint64_t)stride + g] = (float)scores[g];
            if (dev_scores)
                for (int g = 0; g < G && g < stride; ++g) {
                    double dv = dev_scores[(i * K + kid) * (uint64_t)stride + g];
                    double err = fabs(dv - scores[g]) / fmax(1.0, fabs(scores[g]));
                    if (!(err <= *max_err)) *max_err = err; /* also catches NaN */
                }
            const double u = lo_uniform(seed, 0, offset + i, (uint32_t)kid);
            int g;
            if (replay_picks) {
                g = (int)replay_picks[i * K + kid];
                if (g < 0 || g >= G) { ++*n_bad; free(scores); return lo_fail("replay pick out of range"); }
                double m = -INFINITY, total = 0, lo = 0;
                for (int j = 0; j < G; ++j) if (scores[j] > m) m = scores[j];
                for (int j = 0; j < G; ++j) { scores[j] = exp(scores[j] - m); total += scores[j]; }
                for (int j = 0; j < g; ++j) lo += scores[j];
                const double t = u * total, hi = lo + scores[g];
                if (!(scores[g] > 0 && t >= lo - tol * total && t <= hi + tol * total)) ++*n_bad;
            } else {
                g = om_sample_from_scores(u, scores, G);
            }
            if (debug_picks) debug_picks[i * K + kid] = (uint32_t)g;
            const uint32_t gid = k->p2g[g];
            kind_add_row(e, k, r, g);
            k->assign[r] = gid;
        }
    }
    free(scores);
    return 0;
}

static int sweep_all(engine_t *e, const uint32_t *actions, uint64_t n, uint64_t seed,
                     uint64_t offset, const uint8_t *kind_mask, const uint32_t *replay_picks,
                     uint32_t *debug_picks, float *debug_scores, int stride, double tol,
                     const float *dev_scores, int64_t *n_bad, double *max_err) {
    if (!e->R) return lo_fail("cat_sweep before set_rows");
    int rc = 0;
    int64_t bad = 0;
    double merr = 0;
    /* one thread per kind: src/cat_pipeline.cc:103-125 */
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads()) reduction(+ : bad) reduction(max : merr)
    for (int k = 0; k < e->K; ++k) {
        if (kind_mask && !kind_mask[k]) continue;
        int64_t b = 0;
        double me = 0;
        int r = kind_sweep(e, k, actions, n, seed, offset, replay_picks, debug_picks,
                           debug_scores, stride, tol, dev_scores, &b, &me);
        bad += b;
        if (me > merr || me != me) merr = me != me ? INFINITY : me;
        if (r) {
#pragma omp critical
            rc = r;
        }
    }
    if (n_bad) *n_bad = bad;
    if (max_err) *max_err = merr;
    return rc;
}

int lo_cat_sweep(lb_handle h, const uint32_t *actions, uint64_t n, uint64_t seed,
                 uint64_t offset, const uint8_t *kind_mask, uint32_t *debug_picks,
                 float *debug_scores, int32_t stride) {
    return sweep_all(h, actions, n, seed, offset, kind_mask, NULL, debug_picks, debug_scores,
                     stride, 0, NULL, NULL, NULL);
}

/* loom/tasks.py:173-194: samples are independent chains; the restatement runs them one after the
 * other (bench.py runs reference chains in parallel processes, as loom's parallel_map does) */
int lo_cat_sweep_multi(const lb_handle *handles, int32_t n_handles, const uint32_t *actions, uint64_t n,
                       const uint64_t *seeds, uint64_t offset) {
    for (int32_t c = 0; c < n_handles; ++c) {
        int rc = lo_cat_sweep(handles[c], actions, n, seeds[c], offset, NULL, NULL, NULL, 0);
        if (rc) return rc;
    }
    return 0;
}

/* loom/tasks.py:228-233: every sample shuffles the rows with its own seed -> a row order per chain */
int lo_cat_sweep_streams(const lb_handle *handles, int32_t n_handles, const uint32_t *actions, uint64_t n,
                         const uint64_t *seeds, uint64_t offset) {
    for (int32_t c = 0; c < n_handles; ++c) {
        int rc = lo_cat_sweep(handles[c], actions + (uint64_t)c * n, n, seeds[c], offset, NULL, NULL, NULL, 0);
        if (rc) return rc;
    }
    return 0;
}

int64_t lo_cat_replay(lb_handle h, const uint32_t *actions, uint64_t n, uint64_t seed,
                      uint64_t offset, const uint32_t *picks, const float *dev_scores,
                      int32_t stride, double tol, double *max_score_err) {
    int64_t bad = 0;
    double merr = 0;
    int rc = sweep_all(h, actions, n, seed, offset, NULL, picks, NULL, NULL, stride, tol,
                       dev_scores, &bad, &merr);
    if (max_score_err) *max_score_err = merr;
    return rc ? -1 : bad;
}

/* ------------------------------------------------------------------------ */
int lo_get_assignments(lb_handle h, int32_t kind, uint32_t *out) {
    engine_t *e = h;
    if (kind < 0 || kind >= e->K) return lo_fail("bad kind %d", kind);
    const kind_t *k = &e->kinds[kind];
    for (uint64_t r = 0; r < e->R; ++r)
        out[r] = k->assign[r] == LB_UNASSIGNED ? LB_UNASSIGNED : k->g2p[k->assign[r]];
    return 0;
}

int lo_set_assignments(lb_handle h, int32_t kind, const uint32_t *packed) {
    engine_t *e = h;
    if (kind < 0 || kind >= e->K) return lo_fail("bad kind %d", kind);
    kind_t *k = &e->kinds[kind];
    for (uint64_t r = 0; r < e->R; ++r) {
        if (packed[r] != LB_UNASSIGNED && packed[r] >= (uint32_t)k->G)
            return lo_fail("set_assignments: group id %u >= G=%d", packed[r], k->G);
    }
    for (uint64_t r = 0; r < e->R; ++r)
        k->assign[r] = packed[r] == LB_UNASSIGNED ? LB_UNASSIGNED : k->p2g[packed[r]];
    return 0;
}

int lo_kind_info(lb_handle h, int32_t kind, lb_kind_info_t *out) {
    engine_t *e = h;
    if (kind < 0 || kind >= e->K) return lo_fail("bad kind %d", kind);
    const kind_t *k = &e->kinds[kind];
    out->n_features = k->nf; out->n_groups = k->G;
    out->n_int_rows = k->n_int_rows; out->n_flt_rows = k->n_flt_rows;
    out->sample_size = k->sample_size;
    out->py_alpha = (float)k->alpha; out->py_d = (float)k->d;
    return 0;
}

int lo_get_groups(lb_handle h, int32_t kind, uint32_t *counts, uint32_t *ints, double *flts) {
    engine_t *e = h;
    if (kind < 0 || kind >= e->K) return lo_fail("bad kind %d", kind);
    const kind_t *k = &e->kinds[kind];
    const int G = k->G;
    if (counts) memcpy(counts, k->counts, 4u * G);
    if (ints)
        for (int r = 0; r < k->n_int_rows; ++r)
            memcpy(ints + (size_t)r * G, k->ints + (size_t)r * k->Gcap, 4u * G);
    if (flts)
        for (int r = 0; r < k->n_flt_rows; ++r)
            memcpy(flts + (size_t)r * G, k->flts + (size_t)r * k->Gcap, 8u * G);
    return 0;
}

int lo_set_groups(lb_handle h, int32_t kind, int32_t n, const uint32_t *counts,
                  const uint32_t *ints, const double *flts) {
    engine_t *e = h;
    if (kind < 0 || kind >= e->K) return lo_fail("bad kind %d", kind);
    kind_t *k = &e->kinds[kind];
    const int G = n + e->empty_group_count;
    int *fids = xcalloc(k->nf, sizeof(int));
    memcpy(fids, k->fids, sizeof(int) * k->nf);
    uint32_t *c = xcalloc(G, 4);
    if (counts) memcpy(c, counts, 4u * n);
    const int nf = k->nf;
    const double alpha = k->alpha, d = k->d;
    kind_init(e, k, fids, nf, alpha, d, c, G);
    free(fids); free(c);
    if (ints)
        for (int r = 0; r < k->n_int_rows; ++r)
            memcpy(k->ints + (size_t)r * k->Gcap, ints + (size_t)r * n, 4u * n);
    if (flts)
        for (int r = 0; r < k->n_flt_rows; ++r)
            memcpy(k->flts + (size_t)r * k->Gcap, flts + (size_t)r * n, 8u * n);
    /* count_sum rows are derived */
    for (int j = 0; j < k->nf; ++j) {
        const lb_feature_spec *s = &e->specs[k->fids[j]];
        if (s->type != LB_DD && s->type != LB_DPD) continue;
        for (int g = 0; g < G; ++g) {
            uint32_t tot = 0;
            for (int v = 0; v < s->dim; ++v) tot += k->ints[(size_t)(k->int_off[j] + v) * k->Gcap + g];
            k->ints[(size_t)(k->int_off[j] + s->dim) * k->Gcap + g] = tot;
        }
    }
    kind_refresh_all(e, k);
    if (e->R) for (uint64_t r = 0; r < e->R; ++r) k->assign[r] = LB_UNASSIGNED;
    return 0;
}

/* batch score_value with frozen tables */
/* score_diff of a tare-compressed row == score_value of the reconstructed row (src/product_mixture.cc:436-463 adds
 * the tare cache and the pos cells and subtracts the neg cells of the same mixture): the checker holds the decoded
 * block (lo_set_rows_diff), so both forms are one function here */
int lo_score_rows(lb_handle h, int32_t kind, uint64_t b, uint64_t en, float *out);
int lo_score_rows_diff(lb_handle h, int32_t kind, uint64_t b, uint64_t en, float *out) {
    return lo_score_rows(h, kind, b, en, out);
}

int lo_score_rows(lb_handle h, int32_t kind, uint64_t b, uint64_t en, float *out) {
    engine_t *e = h;
    if (kind < 0 || kind >= e->K) return lo_fail("bad kind %d", kind);
    if (b > en || en > e->R) return lo_fail("bad row range");
    if (!out) return 0;
    const kind_t *k = &e->kinds[kind];
    const int G = k->G;
#pragma omp parallel num_threads(n_threads())
    {
        double *scores = xcalloc(G, 8);
#pragma omp for schedule(static)
        for (int64_t r = (int64_t)b; r < (int64_t)en; ++r) {
            kind_score_row(e, k, (uint64_t)r, scores);
            for (int g = 0; g < G; ++g) out[(size_t)(r - (int64_t)b) * G + g] = (float)scores[g];
        }
        free(scores);
    }
    return 0;
}

/* QueryServer::call(Score) for one sample (src/query_server.cc:189-231) */
int lo_query_score(lb_handle h, uint64_t b, uint64_t en, float *out) {
    engine_t *e = h;
    if (b > en || en > e->R) return lo_fail("bad row range");
    if (!out) return 0;
#pragma omp parallel num_threads(n_threads())
    {
        int gmax = 1;
        for (int k = 0; k < e->K; ++k) if (e->kinds[k].G > gmax) gmax = e->kinds[k].G;
        double *scores = xcalloc(gmax, 8);
#pragma omp for schedule(static)
        for (int64_t r = (int64_t)b; r < (int64_t)en; ++r) {
            double total = 0;
            for (int k = 0; k < e->K; ++k) {
                const kind_t *kd = &e->kinds[k];
                kind_score_row(e, kd, (uint64_t)r, scores);
                double m = -INFINITY, sum = 0;
                for (int g = 0; g < kd->G; ++g) if (scores[g] > m) m = scores[g];
                for (int g = 0; g < kd->G; ++g) sum += exp(scores[g] - m);
                total += m + log(sum);
            }
            out[r - (int64_t)b] = (float)total;
        }
        free(scores);
    }
    return 0;
}

/* The per-latent half of QueryServer::call(Sample) (src/query_server.cc:100-176): probs of the
 * conditioning row per kind, then ProductMixture_::sample_value (src/product_mixture.cc:637-649)
 * per sample = a group from probs and every requested feature from that group. */
int lo_query_sample(lb_handle h, uint64_t row, const uint32_t *to_sample, int32_t n_to, uint32_t n_samples,
                    uint64_t seed, uint64_t draw_offset, uint32_t *out_values, uint32_t *out_groups) {
    engine_t *e = h;
    if (row >= e->R) return lo_fail("query_sample: bad row");
    if (n_to < 0 || (n_to && !to_sample)) return lo_fail("query_sample: bad to_sample");
    for (int j = 0; j < n_to; ++j) {
        if (to_sample[j] >= (uint32_t)e->F) return lo_fail("query_sample: bad feature id %u", to_sample[j]);
        if (j && to_sample[j] <= to_sample[j - 1]) return lo_fail("query_sample: to_sample must be ascending");
    }
    if (out_groups)
        for (uint64_t i = 0; i < (uint64_t)n_samples * e->K; ++i) out_groups[i] = LB_UNASSIGNED;
    for (int kid = 0; kid < e->K; ++kid) {
        const kind_t *k = &e->kinds[kid];
        int wanted = 0;
        for (int j = 0; j < n_to; ++j) wanted += (int)e->f2k[to_sample[j]] == kid;
        if (!wanted) continue;                              /* query_server.cc:157-159 */
        double *scores = xcalloc(k->G, 8), *probs = xcalloc(k->G, 8);
        kind_score_row(e, k, row, scores);
        for (uint32_t s = 0; s < n_samples; ++s) {
            memcpy(probs, scores, 8u * (size_t)k->G);
            const int g = om_sample_from_scores(lo_uniform(seed, 5, draw_offset + s, (uint32_t)kid), probs, k->G);
            if (out_groups) out_groups[(uint64_t)s * e->K + kid] = (uint32_t)g;
            for (int j = 0; j < n_to; ++j) {
                const int f = (int)to_sample[j];
                if ((int)e->f2k[f] != kid) continue;
                int local = 0;
                while (k->fids[local] != f) ++local;
                out_values[(uint64_t)s * n_to + j] =
                    om_sample_value(&e->specs[f], e->aux, k->ints + (size_t)k->int_off[local] * k->Gcap + g,
                                    k->flts + (size_t)k->flt_off[local] * k->Gcap + g, k->Gcap, seed,
                                    draw_offset + s, (uint32_t)f, 6, 0);
            }
        }
        free(scores); free(probs);
    }
    return 0;
}

/* bulk add_value / remove_value with given assignments */
int lo_accumulate_rows(lb_handle h, int32_t kind, uint64_t b, uint64_t en, int32_t sign) {
    engine_t *e = h;
    if (kind >= e->K) return lo_fail("bad kind %d", kind);
    if (b > en || en > e->R) return lo_fail("bad row range");
    if (sign != 1 && sign != -1) return lo_fail("sign must be +1 or -1");
    int rc = 0;
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads())
    for (int kid = 0; kid < e->K; ++kid) {
        if (kind >= 0 && kid != kind) continue;
        kind_t *k = &e->kinds[kid];
        const int st = k->Gcap;
        for (uint64_t r = b; r < en; ++r) {
            if (k->assign[r] == LB_UNASSIGNED) continue;
            const uint32_t g = k->g2p[k->assign[r]];
            if (sign < 0 && k->counts[g] == 0) { rc = 1; continue; }
            k->counts[g] += (uint32_t)sign; k->sample_size += (uint64_t)(int64_t)sign;
            for (int j = 0; j < k->nf; ++j) {
                const int f = k->fids[j];
                if (!cell_observed(e, f, r)) continue;
                om_add_value(&e->specs[f], k->ints + (size_t)k->int_off[j] * st + g,
                             k->flts + (size_t)k->flt_off[j] * st + g, st, cell_raw(e, f, r), sign);
            }
        }
        kind_refresh_all(e, k);
    }
    return rc ? lo_fail("accumulate_rows: removal from an empty group") : 0;
}

/* CrossCat::score_data (src/cross_cat.cc:238-250) ->
 * ProductMixture_::score_data (src/product_mixture.cc:581-592). */
int lo_score_data(lb_handle h, double *out) {
    engine_t *e = h;
    double score = 0;
    uint32_t *fc = xcalloc(e->K, 4);
    int nk = 0;
    for (int kid = 0; kid < e->K; ++kid) {
        const kind_t *k = &e->kinds[kid];
        if (!k->nf) continue;
        fc[nk++] = (uint32_t)k->nf;
        score += om_py_score_counts(k->alpha, k->d, k->counts, k->G);
        for (int j = 0; j < k->nf; ++j) {
            const lb_feature_spec *s = &e->specs[k->fids[j]];
            for (int g = 0; g < k->G; ++g)
                score += om_score_data(s, e->aux, k->ints + (size_t)k->int_off[j] * k->Gcap + g,
                                       k->flts + (size_t)k->flt_off[j] * k->Gcap + g, k->Gcap);
        }
    }
    score += om_py_score_counts(e->topo_alpha, e->topo_d, fc, nk);
    free(fc);
    *out = score;
    return 0;
}

int lo_get_kinds(lb_handle h, int32_t *n_kinds, uint32_t *f2k) {
    engine_t *e = h;
    if (n_kinds) *n_kinds = e->K;
    if (f2k) memcpy(f2k, e->f2k, 4u * e->F);
    return 0;
}

int lo_get_hypers(lb_handle h, lb_feature_spec *specs, float *aux, float *kind_py,
                  float *topology_py) {
    engine_t *e = h;
    if (specs) memcpy(specs, e->specs, sizeof(*specs) * e->F);
    if (aux && e->n_aux) memcpy(aux, e->aux, 4u * e->n_aux);
    if (kind_py)
        for (int k = 0; k < e->K; ++k) {
            kind_py[2 * k] = (float)e->kinds[k].alpha;
            kind_py[2 * k + 1] = (float)e->kinds[k].d;
        }
    if (topology_py) { topology_py[0] = (float)e->topo_alpha; topology_py[1] = (float)e->topo_d; }
    return 0;
}

/* ---- timing (monotonic clock stands in for CUDA events) ----------------- */
static double now_ms(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}
int lo_sync(lb_handle h) { (void)h; return 0; }
int lo_timer_start(lb_handle h) { ((engine_t *)h)->t0 = now_ms(); return 0; }
int lo_timer_stop(lb_handle h, float *ms) { *ms = (float)(now_ms() - ((engine_t *)h)->t0); return 0; }
int lo_launch_count(lb_handle h, uint64_t *out, int32_t n, int32_t reset) {
    (void)h; (void)reset;
    for (int i = 0; i < n; ++i) out[i] = 0;
    return 0;
}
int lo_kernel_ms(lb_handle h, double *out, int32_t n, int32_t reset) {
    (void)h; (void)reset;
    for (int i = 0; i < n; ++i) out[i] = 0;
    return 0;
}

/* ---- loom::Differ (src/differ.hpp, src/differ.cc): the tare and sparsify passes --------------- */
#define DIFFER_BINS 16   /* CountSummary::max_count (src/differ.hpp:65); BooleanSummary uses bins 0 and 1 */

/* raw count value of a cell: DPD cells hold dictionary indices */
static uint32_t differ_raw(const engine_t *e, int f, uint32_t cell, const int32_t *dpd_offsets,
                           const uint32_t *dpd_values) {
    if (e->specs[f].type != LB_DPD || !dpd_offsets) return cell;
    int ord = 0;
    for (int g = 0; g < f; ++g) ord += e->specs[g].type == LB_DPD;
    return dpd_values[dpd_offsets[ord] + (int32_t)cell];
}

int lo_differ_reset(lb_handle h) {
    engine_t *e = h;
    if (!e->F) return lo_fail("differ_reset before set_model");
    free(e->differ_hist);
    e->differ_hist = xcalloc((size_t)e->F * DIFFER_BINS, 8);
    e->differ_rows = 0;
    return 0;
}

/* Differ::add_rows (src/differ.cc:93-127) */
int lo_differ_add_rows(lb_handle h, const int32_t *dpd_offsets, const uint32_t *dpd_values) {
    engine_t *e = h;
    if (!e->differ_hist) return lo_fail("differ_add_rows before differ_reset");
    for (int f = 0; f < e->F; ++f) {
        if (e->specs[f].type == LB_NICH) continue;          /* "do not sparsify reals" */
        for (uint64_t r = 0; r < e->R; ++r) {
            if (!cell_observed(e, f, r)) continue;
            const uint32_t raw = differ_raw(e, f, cell_raw(e, f, r), dpd_offsets, dpd_values);
            if (raw < DIFFER_BINS) ++e->differ_hist[(size_t)f * DIFFER_BINS + raw];
        }
    }
    e->differ_rows += e->R;
    return 0;
}

/* Differ::_make_tare / _make_tare_type (src/differ.cc:129-146, 191-206) */
int lo_differ_make_tare(lb_handle h, const int32_t *dpd_offsets, const uint32_t *dpd_values,
                        uint8_t *tare_observed, uint32_t *tare_values) {
    engine_t *e = h;
    if (!e->differ_hist) return lo_fail("differ_make_tare before differ_reset");
    const float count_threshold = (float)(0.5 * (double)e->differ_rows);
    int ord = 0;
    for (int f = 0; f < e->F; ++f) {
        const int is_dpd = e->specs[f].type == LB_DPD;
        const uint64_t *counts = e->differ_hist + (size_t)f * DIFFER_BINS;
        uint32_t mode = 0;
        for (uint32_t i = 0; i < DIFFER_BINS; ++i)
            if (counts[i] > counts[mode]) mode = i;
        tare_observed[f] = e->specs[f].type != LB_NICH && (float)counts[mode] > count_threshold;
        tare_values[f] = 0;
        if (tare_observed[f]) {
            uint32_t cell = mode;
            if (is_dpd && dpd_offsets) {                    /* back to the dictionary index */
                cell = LB_UNASSIGNED;
                for (int32_t i = dpd_offsets[ord]; i < dpd_offsets[ord + 1]; ++i)
                    if (dpd_values[i] == mode) cell = (uint32_t)(i - dpd_offsets[ord]);
                if (cell == LB_UNASSIGNED) return lo_fail("differ_make_tare: mode of feature %d is not in its dictionary", f);
            }
            tare_values[f] = cell;
        }
        ord += is_dpd;
    }
    return 0;
}

/* Differ::compress_rows -> _abs_to_rel_type (src/differ.cc:148-189, 248-298) */
int lo_differ_compress_rows(lb_handle h, const uint8_t *tare_observed, const uint32_t *tare_values,
                            uint64_t *n_pos, uint64_t *n_neg) {
    engine_t *e = h;
    if (!e->F) return lo_fail("differ_compress_rows before set_model");
    free(e->diff_pos_ptr); free(e->diff_neg_ptr); free(e->diff_pos_feature); free(e->diff_pos_value); free(e->diff_neg_feature);
    e->diff_pos_ptr = xcalloc(e->R + 1, 8);
    e->diff_rows = e->R;
    e->diff_neg_ptr = xcalloc(e->R + 1, 8);
    for (int pass = 0; pass < 2; ++pass) {
        uint64_t np = 0, nn = 0;
        for (uint64_t r = 0; r < e->R; ++r) {
            for (int f = 0; f < e->F; ++f) {
                const int obs = cell_observed(e, f, r);
                const uint32_t v = obs ? cell_raw(e, f, r) : 0;
                int to_pos, to_neg;
                if (tare_observed[f]) {
                    const int differs = obs && v != tare_values[f];
                    to_pos = differs;
                    to_neg = !obs || differs;
                } else {
                    to_pos = obs;
                    to_neg = 0;
                }
                if (to_pos) { if (pass) { e->diff_pos_feature[np] = (uint32_t)f; e->diff_pos_value[np] = v; } ++np; }
                if (to_neg) { if (pass) e->diff_neg_feature[nn] = (uint32_t)f; ++nn; }
            }
            e->diff_pos_ptr[r + 1] = np;
            e->diff_neg_ptr[r + 1] = nn;
        }
        if (!pass) {
            e->diff_pos_feature = xcalloc(np + 1, 4);
            e->diff_pos_value = xcalloc(np + 1, 4);
            e->diff_neg_feature = xcalloc(nn + 1, 4);
        }
        *n_pos = np; *n_neg = nn;
    }
    return 0;
}

int lo_differ_fetch(lb_handle h, uint64_t *pos_ptr, uint32_t *pos_feature, uint32_t *pos_value,
                    uint64_t *neg_ptr, uint32_t *neg_feature) {
    engine_t *e = h;
    if (!e->diff_pos_ptr || e->diff_rows != e->R) return lo_fail("differ_fetch before differ_compress_rows");
    const uint64_t np = e->diff_pos_ptr[e->R], nn = e->diff_neg_ptr[e->R];
    memcpy(pos_ptr, e->diff_pos_ptr, 8 * (e->R + 1));
    memcpy(neg_ptr, e->diff_neg_ptr, 8 * (e->R + 1));
    memcpy(pos_feature, e->diff_pos_feature, 4 * np);
    memcpy(pos_value, e->diff_pos_value, 4 * np);
    memcpy(neg_feature, e->diff_neg_feature, 4 * nn);
    return 0;
}

/* ---- Loom::generate -> generate_rows (src/generate.hpp:36-93) ------------------------------------------ */
int lo_generate_rows(lb_handle h, uint64_t R, float density, uint64_t seed) {
    engine_t *e = h;
    if (!e->F) return lo_fail("generate_rows before set_model");
    if (!(density >= 0.f && density <= 1.f)) return lo_fail("generate_rows: density must be in [0, 1]");
    if (R == 0 || R >= (1ull << 31)) return lo_fail("generate_rows: bad row count");
    const uint64_t W = (R + 31) / 32;
    uint8_t *cat8 = xcalloc((size_t)e->F8 * R + 1, 1);
    uint32_t *wide = xcalloc((size_t)e->FW * R + 1, 4), *mask = xcalloc((size_t)e->F * W, 4);
    int rc = lo_set_rows(h, R, cat8, wide, mask);
    free(cat8); free(wide); free(mask);
    if (rc) return rc;
    uint32_t *pick = xcalloc(R, 4), *counts = xcalloc(R + 1, 4);
    double *p = xcalloc(R + 2, 8);
    for (int kid = 0; kid < e->K && !rc; ++kid) {
        /* the kind's Pitman-Yor process: Clustering::PitmanYor seating, row by row */
        const double alpha = e->kinds[kid].alpha, d = e->kinds[kid].d;
        int m = 0;
        for (uint64_t r = 0; r < R; ++r) {
            double total = 0;
            for (int g = 0; g < m; ++g) total += p[g] = counts[g] - d;
            total += p[m] = alpha + d * m;
            const int g = om_sample_from_likelihoods(lo_uniform(seed, 7, r, (uint32_t)kid), p, m + 1, total);
            if (g == m) counts[m++] = 0;
            counts[g]++;
            pick[r] = (uint32_t)g;
        }
        rc = lo_set_groups(h, kid, m, counts, NULL, NULL);
        if (!rc) rc = lo_set_assignments(h, kid, pick);
        if (rc) break;
        kind_t *k = &e->kinds[kid];
        /* values: inside a group the rows are visited in order, every (group, feature) chain is independent */
        for (int j = 0; j < k->nf; ++j) {
            const int f = k->fids[j];
            const lb_feature_spec *s = &e->specs[f];
            for (uint64_t r = 0; r < R; ++r) {
                const uint32_t g = pick[r];
                if (!(lo_uniform(seed, 8, r, ((uint32_t)f << 10)) < (double)density)) continue;
                uint32_t *in = k->ints + (size_t)k->int_off[j] * k->Gcap + g;
                double *fl = k->flts + (size_t)k->flt_off[j] * k->Gcap + g;
                const uint32_t raw = om_sample_value(s, e->aux, in, fl, k->Gcap, seed, r, (uint32_t)f, 8, 1);
                if (s->type == LB_DPD && raw == LB_DPD_OTHER) continue;
                om_add_value(s, in, fl, k->Gcap, raw, +1);
                if (f < e->F8) e->cat8[(size_t)f * R + r] = (uint8_t)raw;
                else e->wide[(size_t)(f - e->F8) * R + r] = raw;
                e->mask[(size_t)f * W + (r >> 5)] |= 1u << (r & 31);
            }
        }
        kind_refresh_all(e, k);
    }
    free(pick); free(counts); free(p);
    return rc;
}

int lo_get_rows(lb_handle h, uint8_t *cat8, uint32_t *wide, uint32_t *mask) {
    engine_t *e = h;
    if (!e->R) return lo_fail("get_rows: no rows loaded");
    if (e->F8) memcpy(cat8, e->cat8, (size_t)e->F8 * e->R);
    if (e->FW) memcpy(wide, e->wide, 4 * (size_t)e->FW * e->R);
    memcpy(mask, e->mask, 4 * (size_t)e->F * e->W);
    return 0;
}
