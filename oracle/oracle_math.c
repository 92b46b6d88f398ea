/*
 * oracle_math.c -- float64 closed forms of the component models.  TEST
 * INFRASTRUCTURE ONLY (see loom_oracle.h; parity unpinned vs the reference
 * binary, pinned vs scipy in tests/test_oracle_math.py).
 *
 * The reference takes these from forcedotcom/distributions 2.0.28 (not
 * vendored; call sites src/models.hpp:78-121, src/product_mixture.cc:143-633).
 * The formulas are the standard conjugate results listed in SURVEY.md 8(c)
 * (Murphy 2007, doc/references.bib:55-62).
 */
#include "oracle_internal.h"

/* ---- Philox4x32-10 (Salmon et al. 2011, Random123) ---------------------- */
void lo_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int i = 0; i < 10; ++i) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

double lo_uniform(uint64_t seed, uint32_t stream, uint64_t a, uint32_t b) {
    uint32_t ctr[4] = {(uint32_t)a, (uint32_t)(a >> 32), b, stream};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t out[4];
    lo_philox4x32(ctr, key, out);
    return (double)(out[0] >> 8) * (1.0 / 16777216.0);
}

/* ---- sampling (distributions::sample_from_scores_overwrite, called at
 * src/cat_kernel.hpp:214; SURVEY 8c "Sampling") --------------------------- */
int om_sample_from_scores(double u, double *scores, int n) {
    double m = -INFINITY;
    for (int i = 0; i < n; ++i) if (scores[i] > m) m = scores[i];
    double total = 0;
    for (int i = 0; i < n; ++i) { scores[i] = exp(scores[i] - m); total += scores[i]; }
    return om_sample_from_likelihoods(u, scores, n, total);
}

/* distributions::sample_from_likelihoods (src/kind_proposer.cc:286-287). */
int om_sample_from_likelihoods(double u, const double *p, int n, double total) {
    double t = u * total, c = 0;
    int last = 0;
    for (int i = 0; i < n; ++i) {
        if (p[i] > 0) last = i;
        c += p[i];
        if (c > t) return i;
    }
    return last;
}

/* ---- Pitman-Yor clustering: score_counts (src/cross_cat.cc:248,
 * src/infer_grid.hpp:119) ------------------------------------------------- */
double om_py_score_counts(double alpha, double d, const uint32_t *counts, int n) {
    double score = 0, total = 0;
    int m = 0;
    for (int i = 0; i < n; ++i) {
        if (counts[i]) {
            score += log(alpha + d * m);
            score += lgamma(counts[i] - d) - lgamma(1.0 - d);
            total += counts[i];
            ++m;
        }
    }
    score += lgamma(alpha) - lgamma(alpha + total);
    return score;
}
double lo_py_score_counts(double alpha, double d, const uint32_t *counts, int n) {
    return om_py_score_counts(alpha, d, counts, n);
}

/* ---- per-feature predictive / marginal / update -------------------------
 * Stats of one group live at ints[row * stride], flts[row * stride]. */
#define I(row) ((double)ints[(size_t)(row) * stride])
#define D(row) (flts[(size_t)(row) * stride])

double om_score_value(const lb_feature_spec *s, const float *aux, const uint32_t *ints,
                      const double *flts, int stride, uint32_t raw) {
    switch (s->type) {
    case LB_BB: {
        double a = s->hp[0], b = s->hp[1], h = I(0), t = I(1);
        return log((raw ? a + h : b + t) / (a + b + h + t));
    }
    case LB_DD: {
        const float *al = aux + s->aux_off;
        double sum = 0;
        for (int v = 0; v < s->dim; ++v) sum += al[v];
        return log((al[raw] + I(raw)) / (sum + I(s->dim)));
    }
    case LB_DPD: {
        const float *be = aux + s->aux_off;
        double alpha = s->hp[1];
        return log((alpha * be[raw] + I(raw)) / (alpha + I(s->dim)));
    }
    case LB_GP: {
        double a = s->hp[0] + I(1), b = s->hp[1] + I(0), x = raw;
        return lgamma(a + x) - lgamma(a) - lgamma(x + 1.0) + a * log(b / (b + 1.0)) -
               x * log(1.0 + b);
    }
    case LB_NICH: {
        float xf; memcpy(&xf, &raw, 4);
        double x = xf, mu = s->hp[0], kappa = s->hp[1], sigmasq = s->hp[2], nu = s->hp[3];
        double n = I(0), mean = D(0), ctv = D(1);
        double kn = kappa + n, mun = (kappa * mu + n * mean) / kn, nun = nu + n;
        double dm = mu - mean;
        double sn = (nu * sigmasq + ctv + n * kappa * dm * dm / kn) / nun;
        double lambda = kn / ((kn + 1.0) * sn);
        double dx = x - mun;
        return lgamma(0.5 * (nun + 1.0)) - lgamma(0.5 * nun) +
               0.5 * log(lambda / (M_PI * nun)) -
               0.5 * (nun + 1.0) * log1p(lambda * dx * dx / nun);
    }
    }
    return NAN;
}

double om_score_data(const lb_feature_spec *s, const float *aux, const uint32_t *ints,
                     const double *flts, int stride) {
    switch (s->type) {
    case LB_BB: {
        double a = s->hp[0], b = s->hp[1], h = I(0), t = I(1);
        return lgamma(a + h) - lgamma(a) + lgamma(b + t) - lgamma(b) + lgamma(a + b) -
               lgamma(a + b + h + t);
    }
    case LB_DD: {
        const float *al = aux + s->aux_off;
        double sum = 0, score = 0;
        for (int v = 0; v < s->dim; ++v) {
            sum += al[v];
            if (ints[(size_t)v * stride]) score += lgamma(al[v] + I(v)) - lgamma((double)al[v]);
        }
        return score + lgamma(sum) - lgamma(sum + I(s->dim));
    }
    case LB_DPD: {
        const float *be = aux + s->aux_off;
        double alpha = s->hp[1], score = 0;
        for (int v = 0; v < s->dim; ++v)
            if (ints[(size_t)v * stride])
                score += lgamma(alpha * be[v] + I(v)) - lgamma(alpha * be[v]);
        return score + lgamma(alpha) - lgamma(alpha + I(s->dim));
    }
    case LB_GP: {
        double al = s->hp[0], ib = s->hp[1];
        double a = al + I(1), b = ib + I(0);
        return lgamma(a) - lgamma(al) + al * log(ib) - a * log(b) - D(0);
    }
    case LB_NICH: {
        double mu = s->hp[0], kappa = s->hp[1], sigmasq = s->hp[2], nu = s->hp[3];
        double n = I(0), mean = D(0), ctv = D(1);
        double kn = kappa + n, nun = nu + n, dm = mu - mean;
        double sn = (nu * sigmasq + ctv + n * kappa * dm * dm / kn) / nun;
        return lgamma(0.5 * nun) - lgamma(0.5 * nu) + 0.5 * log(kappa / kn) +
               0.5 * nu * log(nu * sigmasq) - 0.5 * nun * log(nun * sn) -
               0.5 * n * log(M_PI);
    }
    }
    return NAN;
}
#undef I
#undef D

/* Group::add_value / remove_value of the component models (called through
 * src/product_mixture.cc:148-163, 203-218). */
void om_add_value(const lb_feature_spec *s, uint32_t *ints, double *flts, int stride,
                  uint32_t raw, int sign) {
#define IR(row) ints[(size_t)(row) * stride]
#define DR(row) flts[(size_t)(row) * stride]
    switch (s->type) {
    case LB_BB:
        IR(raw ? 0 : 1) += (uint32_t)sign;
        break;
    case LB_DD: case LB_DPD:
        IR(raw) += (uint32_t)sign;
        IR(s->dim) += (uint32_t)sign;
        break;
    case LB_GP:
        IR(0) += (uint32_t)sign;
        IR(1) += (uint32_t)(sign * (int64_t)raw);
        DR(0) += sign * lgamma((double)raw + 1.0);
        if (IR(0) == 0) DR(0) = 0;
        break;
    case LB_NICH: {
        float xf; memcpy(&xf, &raw, 4);
        double x = xf, n = IR(0), mean = DR(0), ctv = DR(1);
        if (sign > 0) {
            n += 1;
            double delta = x - mean;
            mean += delta / n;
            ctv += delta * (x - mean);
        } else {
            double total = mean * n, delta = x - mean;
            n -= 1;
            if (n == 0) { mean = 0; ctv = 0; }
            else { mean = (total - x) / n; ctv -= delta * (x - mean); }
        }
        IR(0) = (uint32_t)n; DR(0) = mean; DR(1) = ctv;
        break;
    }
    }
#undef IR
#undef DR
}

double lo_feature_score_value(const lb_feature_spec *spec, const float *aux,
                              const uint32_t *ints, const double *flts, uint32_t raw) {
    return om_score_value(spec, aux, ints, flts, 1, raw);
}
double lo_feature_score_data(const lb_feature_spec *spec, const float *aux,
                             const uint32_t *ints, const double *flts) {
    return om_score_data(spec, aux, ints, flts, 1);
}

/* ---- Group::sample_value (distributions 2.0.28; called through
 * ProductMixture_::sample_value, src/product_mixture.cc:622-649) ------------
 * distributions draws the group's parameter from its posterior (Sampler::init) and then
 * the value from that parameter (Sampler::eval), which is a draw from the posterior
 * predictive.  Restated here on counter-based uniforms so that a draw is a pure function of
 * (seed, sample, feature): u(j) = Philox(seed, stream 6, a, featureid << 10 | j).
 *   BB    Bernoulli((alpha + heads) / (alpha + beta + heads + tails))            [1 uniform]
 *   DD    categorical over alpha_v + n_v                                          [1]
 *   DPD   categorical over alpha beta_v + n_v, plus alpha beta0 for an unseen value [1]
 *   GP    lambda ~ Gamma(alpha + sum, 1 / (inv_beta + count)) (Marsaglia-Tsang 2000),
 *         x ~ Poisson(lambda) (inversion below 30, Hoermann's PTRS 1993 above)
 *   NICH  Student-t, nu' degrees of freedom, location mu', scale^2 sigma'^2 (kappa'+1)/kappa'
 *         by Bailey's polar method (1994): t = sqrt(nu (U^(-2/nu) - 1)) cos(2 pi V)   [2] */
#define I(row) ((double)ints[(size_t)(row) * stride])
#define D(row) (flts[(size_t)(row) * stride])
typedef struct { uint64_t seed, a; uint32_t base, j, stream; } draw_t;
static double draw_u(draw_t *d) { return lo_uniform(d->seed, d->stream, d->a, d->base | d->j++); }          /* [0,1) */
static double draw_uo(draw_t *d) { return draw_u(d) + 0.5 / 16777216.0; }                          /* (0,1) */

static double sample_gamma(draw_t *d, double shape) {
    double boost = 1.0;
    if (shape < 1.0) { boost = pow(draw_uo(d), 1.0 / shape); shape += 1.0; }
    const double dd = shape - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * dd);
    for (int it = 0; it < 100; ++it) {
        const double u1 = draw_uo(d), u2 = draw_u(d), u3 = draw_uo(d);
        const double z = sqrt(-2.0 * log(u1)) * cos(2.0 * M_PI * u2);
        double v = 1.0 + c * z;
        if (v <= 0) continue;
        v = v * v * v;
        if (log(u3) < 0.5 * z * z + dd - dd * v + dd * log(v)) return boost * dd * v;
    }
    return boost * dd;
}

static double sample_poisson(draw_t *d, double lam) {
    if (lam < 30.0) {
        const double t = draw_u(d);
        double p = exp(-lam), c = p, x = 0;
        while (t >= c && x < 1000) { x += 1; p *= lam / x; c += p; }
        return x;
    }
    const double slam = sqrt(lam), loglam = log(lam), b = 0.931 + 2.53 * slam, a = -0.059 + 0.02483 * b;
    const double inv_alpha = 1.1239 + 1.1328 / (b - 3.4), vr = 0.9277 - 3.6224 / (b - 2.0);
    for (int it = 0; it < 100; ++it) {
        const double u = draw_u(d) - 0.5, v = draw_uo(d), us = 0.5 - fabs(u);
        const double k = floor((2.0 * a / us + b) * u + lam + 0.43);
        if (us >= 0.07 && v <= vr) return k;
        if (k < 0 || (us < 0.013 && v > us)) continue;
        if (log(v) + log(inv_alpha) - log(a / (us * us) + b) <= -lam + k * loglam - lgamma(k + 1.0)) return k;
    }
    return floor(lam);
}

uint32_t om_sample_value(const lb_feature_spec *s, const float *aux, const uint32_t *ints, const double *flts,
                         int stride, uint64_t seed, uint64_t a, uint32_t featureid, uint32_t stream, uint32_t first_draw) {
    draw_t d = {seed, a, featureid << 10, first_draw, stream};
    switch (s->type) {
    case LB_BB: {
        double al = s->hp[0], be = s->hp[1], h = I(0), t = I(1);
        return draw_u(&d) < (al + h) / (al + be + h + t);
    }
    case LB_DD:
    case LB_DPD: {
        const float *w0 = aux + s->aux_off;
        const double scale = s->type == LB_DPD ? (double)s->hp[1] : 1.0;
        const int n = s->dim + (s->type == LB_DPD);
        double total = 0;
        for (int v = 0; v < n; ++v) total += v < s->dim ? scale * w0[v] + I(v) : scale * s->hp[2];
        const double t = draw_u(&d) * total;
        double c = 0;
        int last = 0;
        for (int v = 0; v < n; ++v) {
            const double w = v < s->dim ? scale * w0[v] + I(v) : scale * s->hp[2];
            if (w > 0) last = v;
            c += w;
            if (c > t) { last = v; break; }
        }
        return last == s->dim ? LB_DPD_OTHER : (uint32_t)last;
    }
    case LB_GP: {
        const double shape = s->hp[0] + I(1), rate = s->hp[1] + I(0);
        const double lam = sample_gamma(&d, shape) / rate;
        const double x = sample_poisson(&d, lam);
        return x >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)x;
    }
    case LB_NICH: {
        double mu = s->hp[0], kappa = s->hp[1], sigmasq = s->hp[2], nu = s->hp[3];
        double n = I(0), mean = D(0), ctv = D(1);
        double kn = kappa + n, mun = (kappa * mu + n * mean) / kn, nun = nu + n, dm = mu - mean;
        double sn = (nu * sigmasq + ctv + n * kappa * dm * dm / kn) / nun;
        const double u = draw_uo(&d), v = draw_u(&d);
        const double t = sqrt(nun * (pow(u, -2.0 / nun) - 1.0)) * cos(2.0 * M_PI * v);
        const float x = (float)(mun + sqrt(sn * (kn + 1.0) / kn) * t);
        uint32_t raw; memcpy(&raw, &x, 4);
        return raw;
    }
    }
    return 0;
}
#undef I
#undef D
