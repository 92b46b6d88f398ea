// lb_common.cuh -- device structs and math shared by every kernel of
// libloomb200.so (sm_100a).  No torch, no CPU fallback.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/loom_b200.h"

// ---------------------------------------------------------------------------
// Device-side layout.
//
// Every per-group table of a kind is a matrix [rows][Gcap] with the group index
// fastest, so the G-wide vector a row's cell touches (one table row) is one
// coalesced line -- the GPU form of distributions' cached scorer vectors
// ("(dim+1) x G table").
//
//   ints  uint32 [n_int_rows][Gcap]   BB: heads,tails | DD/DPD: counts[dim],count_sum
//                                     GP: count,sum   | NICH: count
//   flts  double [n_flt_rows][Gcap]   GP: log_prod    | NICH: mean, count_times_variance
//   cache float  [n_cache_rows][Gcap] BB: log p(1), log p(0)
//                                     DD/DPD: log(a_v + n_v) [dim], log(A + N)
//                                     GP: a'*log(b'/(b'+1)), -log(1+b')
//                                     NICH: c0, (nu'+1)/2, lambda/nu', mu'_hi, mu'_lo
//   one extra all-zero cache row (index n_cache_rows) lets scorers read
//   "unobserved" cells branch-free
struct FeatDev {
    int32_t type, dim;
    float hp[4];
    int32_t aux_off;
    int32_t col;        // column in cat8 (f < F8) or wide (f - F8)
    int32_t is_wide;
    int32_t int_row, flt_row, cache_row;  // first rows inside the owning kind's tables
    float a_total;      // DD: sum(alphas); DPD: alpha
    float a_scale;      // DD: 1; DPD: alpha  (a_v = a_scale * aux[v])
};

struct KindDev {
    int32_t nf;
    int32_t feat_begin;   // into the engine's kind-major feature list
    float alpha, d;
    int32_t G, Gcap;
    uint64_t sample_size;
    uint32_t *counts;     // [Gcap]
    float *shifted;       // [Gcap] log(n_g - d)
    uint32_t *p2g;        // [Gcap] packed -> global (MixtureIdTracker)
    uint32_t *g2p;        // [g2p_cap] global -> packed
    uint32_t g2p_cap, next_global;
    uint32_t *ints;
    double *flts;
    float *cache;
    const int32_t *cache_row_feat;  // [n_cache_rows] kind-local feature index owning each cache row
    uint32_t *assign;     // [R] global group id per row position
    int32_t n_int_rows, n_flt_rows, n_cache_rows;
    int32_t status;       // LB_ST_*
    uint64_t progress;    // actions completed by the sweep kernel
};

enum { LB_ST_OK = 0, LB_ST_NEED_GROW = 1, LB_ST_DUPLICATE_ROW = 2, LB_ST_REMOVE_UNASSIGNED = 3,
       LB_ST_BAD_ROW = 4 };

struct RowsDev {
    uint64_t R, W;        // rows, mask words per feature
    const uint8_t *cat8;  // [F8][R]
    const uint32_t *wide; // [FW][R]
    const uint32_t *mask; // [F][W]
};

// ---------------------------------------------------------------------------
// Philox4x32-10, identical to oracle/oracle_math.c (counter a_lo,a_hi,b,stream).
__host__ __device__ inline uint32_t lb_philox_u32(uint64_t seed, uint32_t stream, uint64_t a,
                                                  uint32_t b) {
    uint32_t c0 = (uint32_t)a, c1 = (uint32_t)(a >> 32), c2 = b, c3 = stream;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
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
    return c0;
}
__host__ __device__ inline float lb_uniform(uint64_t seed, uint32_t stream, uint64_t a, uint32_t b) {
    return (float)(lb_philox_u32(seed, stream, a, b) >> 8) * (1.0f / 16777216.0f);
}

// ---------------------------------------------------------------------------
// fp32 special functions with the cancellation taken out analytically, so that
// scores agree with the float64 oracle to ~1e-6 even when a group holds 1e5+
// rows (plain lgammaf(a+x) - lgammaf(a) loses everything there).
#ifdef __CUDACC__

// Stirling tail: lgamma(z) = (z-.5)log z - z + .5 log(2pi) + lb_stirling_corr(z), z >= 8
__device__ __forceinline__ float lb_stirling_corr(float z) {
    const float r = __frcp_rn(z), r2 = r * r;
    return r * (0.0833333333f - r2 * (0.00277777778f - r2 * 0.000793650794f));
}
__device__ __forceinline__ float lb_lgamma_big(float z) {  // z >= 8
    return (z - 0.5f) * logf(z) - z + 0.918938533f + lb_stirling_corr(z);
}
// libm's lgamma is several hundred instructions: keep ONE copy out of line.  It only serves
// small arguments (groups of a few rows); inlined at every call site it pushed the hot loop of
// the sweep kernel past the 32 KB instruction cache.
static __device__ __noinline__ float lb_lgammaf_small(float z) { return lgammaf(z); }
static __device__ __noinline__ double lb_lgamma_d(double z) { return lgamma(z); }
__device__ __forceinline__ float lb_lgamma(float z) {
    return z >= 8.f ? lb_lgamma_big(z) : lb_lgammaf_small(z);
}
// lgamma(a + x) - lgamma(a), a > 0, x >= 0
__device__ __forceinline__ float lb_lgamma_diff(float a, float x) {
    if (x == 0.f) return 0.f;
    const float b = a + x;
    if (a >= 8.f) {
        return (a - 0.5f) * log1pf(x / a) + x * (logf(b) - 1.f) +
               (lb_stirling_corr(b) - lb_stirling_corr(a));
    }
    if (x <= 4.f && x == floorf(x)) {  // small integer steps: log of the rising factorial
        float p = a;
        for (float j = 1.f; j < x; j += 1.f) p *= a + j;
        return logf(p);
    }
    return lb_lgamma(b) - lb_lgammaf_small(a);
}
// lgamma(a + x) - lgamma(a) for an integer count x: short rising factorials are
// one or two logs of an exact-ish product; longer ones use the Stirling form.
__device__ __forceinline__ float lb_lgamma_diff_int(float a, uint32_t x) {
    if (x == 0u) return 0.f;
    if (x <= 4u) {
        float p = a;
        if (x > 1u) p *= a + 1.f;
        if (x > 2u) p *= a + 2.f;
        if (x > 3u) p *= a + 3.f;
        return logf(p);
    }
    if (x <= 8u) {
        const float p1 = a * (a + 1.f) * (a + 2.f) * (a + 3.f);
        float p2 = a + 4.f;
        if (x > 5u) p2 *= a + 5.f;
        if (x > 6u) p2 *= a + 6.f;
        if (x > 7u) p2 *= a + 7.f;
        return logf(p1) + logf(p2);
    }
    return lb_lgamma_diff(a, (float)x);
}
// log(x!) for an integer count
__device__ __forceinline__ float lb_log_factorial(uint32_t x) {
    const float xf = (float)x;
    if (x < 2u) return 0.f;
    if (x < 8u) {
        float p = 2.f;
        for (uint32_t j = 3u; j <= x; ++j) p *= (float)j;
        return logf(p);
    }
    return lb_lgamma_big(xf + 1.f);
}

// log(x!) in double for the GP log_prod statistic (exact table for small x)
__device__ const double LB_LOG_FACTORIAL[32] = {
    0.0, 0.0, 0.69314718055994531, 1.7917594692280550, 3.1780538303479458, 4.7874917427820458,
    6.5792512120101012, 8.5251613610654147, 10.604602902745251, 12.801827480081469,
    15.104412573075516, 17.502307845873887, 19.987214495661885, 22.552163853123425,
    25.191221182738680, 27.899271383840890, 30.671860106080672, 33.505073450136891,
    36.395445208033053, 39.339884187199495, 42.335616460753485, 45.380138898476908,
    48.471181351835227, 51.606675567764377, 54.784729398112319, 58.003605222980518,
    61.261701761002001, 64.557538627006338, 67.889743137181540, 71.257038967168015,
    74.658236348830158, 78.092223553315307};
__device__ __forceinline__ double lb_log_factorial_d(uint32_t x) {
    return x < 32u ? LB_LOG_FACTORIAL[x] : lb_lgamma_d((double)x + 1.0);
}

// ---- per-feature cached scorer refresh (one group) -------------------------
// `ints`, `flts`, `cache` point at the feature's first row, column g already
// applied; `st` is the column stride (Gcap).
// The *_v forms take the statistics by value: the update path already holds them in registers,
// and re-reading what it just stored costs a second L2 round trip on the critical path.
__device__ __forceinline__ void lb_refresh_bb_v(const FeatDev &f, uint32_t heads, uint32_t tails, float *c, int st) {
    const float h = (float)heads, t = (float)tails;
    const float inv = 1.f / (f.hp[0] + f.hp[1] + h + t);
    c[0] = logf((f.hp[0] + h) * inv);
    c[st] = logf((f.hp[1] + t) * inv);
}
__device__ __forceinline__ void lb_refresh_bb(const FeatDev &f, const uint32_t *in, float *c, int st) {
    lb_refresh_bb_v(f, in[0], in[st], c, st);
}
__device__ __forceinline__ void lb_refresh_gp_v(const FeatDev &f, uint32_t count, uint32_t sum, float *c, int st) {
    const float a = f.hp[0] + (float)sum, b = f.hp[1] + (float)count;
    c[0] = a * log1pf(-1.f / (b + 1.f));
    c[st] = -log1pf(b);
}
__device__ __forceinline__ void lb_refresh_gp(const FeatDev &f, const uint32_t *in, float *c, int st) {
    lb_refresh_gp_v(f, in[0], in[st], c, st);
}
__device__ __forceinline__ void lb_refresh_nich_v(const FeatDev &f, uint32_t count, double mean, double ctv_d,
                                                  float *c, int st) {
    // posterior NIX parameters.  mu' is formed in double and
    // cached as a hi/lo float pair so x - mu' keeps full precision when
    // |mean| >> sigma; everything else is fp32-safe.
    const float mu = f.hp[0], kappa = f.hp[1], sigmasq = f.hp[2], nu = f.hp[3];
    const float n = (float)count, ctv = (float)ctv_d;
    const float kn = kappa + n, inv_kn = 1.f / kn, nun = nu + n;
    // (a float32 reciprocal here would cost mu' a relative 1e-7 -- 5e-6 absolute at |mean| = 50, which a group with
    // sigma = 0.1 turns into 2e-4 of score: tests/test_device_math.py)
    const double mun = ((double)kappa * (double)mu + (double)count * mean) / ((double)kappa + (double)count);
    const float dm = (float)((double)mu - mean);
    const float s_nun = nu * sigmasq + ctv + n * kappa * dm * dm * inv_kn;  // nu' * sigma'^2
    const float prec = kn / ((kn + 1.f) * s_nun);                           // lambda / nu'
    const float half_nu = 0.5f * nun;
    const float mu_hi = (float)mun;
    c[0] = lb_lgamma_diff(half_nu, 0.5f) + 0.5f * logf(prec * 0.318309886f);
    c[st] = half_nu + 0.5f;
    c[2 * (size_t)st] = prec;
    c[3 * (size_t)st] = mu_hi;
    c[4 * (size_t)st] = (float)(mun - (double)mu_hi);
}
__device__ __forceinline__ void lb_refresh_nich(const FeatDev &f, const uint32_t *in,
                                                const double *fl, float *c, int st) {
    lb_refresh_nich_v(f, in[0], fl[0], fl[st], c, st);
}

// Recompute one cached-scorer row (`local` = row index inside the feature) of
// one group from its sufficient statistics.  BB/GP/NICH rows are produced
// together by the thread that owns local row 0.
__device__ __forceinline__ void lb_refresh_row(const FeatDev &f, const float *aux, const uint32_t *in,
                                               const double *fl, float *c, int st, int local) {
    switch (f.type) {
    case LB_BB:
        if (local == 0) lb_refresh_bb(f, in, c, st);
        break;
    case LB_DD:
    case LB_DPD:
        if (local < f.dim)
            c[(size_t)local * st] = logf(f.a_scale * aux[f.aux_off + local] + (float)in[(size_t)local * st]);
        else
            c[(size_t)f.dim * st] = logf(f.a_total + (float)in[(size_t)f.dim * st]);
        break;
    case LB_GP:
        if (local == 0) lb_refresh_gp(f, in, c, st);
        break;
    case LB_NICH:
        if (local == 0) lb_refresh_nich(f, in, fl, c, st);
        break;
    }
}

// Group::add_value / remove_value for one cell (sign = +1 / -1), then refresh
// the cached scorer entries it invalidated.  Exactly one thread owns a
// (feature, group) pair at a time, so plain loads/stores suffice.
__device__ __forceinline__ void lb_update_cell(const FeatDev &f, const float *aux, uint32_t *in,
                                               double *fl, float *c, int st, uint32_t raw, int sign) {
    switch (f.type) {
    case LB_BB: {
        uint32_t h = in[0], t = in[st];        // both loads in flight together
        if (raw) { h += (uint32_t)sign; in[0] = h; } else { t += (uint32_t)sign; in[st] = t; }
        lb_refresh_bb_v(f, h, t, c, st);
        break;
    }
    case LB_DD:
    case LB_DPD: {
        const uint32_t n = in[(size_t)raw * st] + (uint32_t)sign;
        const uint32_t tot = in[(size_t)f.dim * st] + (uint32_t)sign;
        in[(size_t)raw * st] = n;
        in[(size_t)f.dim * st] = tot;
        c[(size_t)raw * st] = logf(f.a_scale * aux[f.aux_off + raw] + (float)n);
        c[(size_t)f.dim * st] = logf(f.a_total + (float)tot);
        break;
    }
    case LB_GP: {
        const uint32_t n0 = in[0], s0 = in[st];
        const double lp0 = fl[0];
        const uint32_t n = n0 + (uint32_t)sign, sum = s0 + (uint32_t)(sign * (int64_t)raw);
        in[0] = n;
        in[st] = sum;
        fl[0] = n ? lp0 + sign * (double)lb_log_factorial_d(raw) : 0.0;
        lb_refresh_gp_v(f, n, sum, c, st);
        break;
    }
    case LB_NICH: {
        const double x = (double)__uint_as_float(raw);
        double n = (double)in[0], mean = fl[0], ctv = fl[st];
        if (sign > 0) {
            n += 1.0;
            const double delta = x - mean;
            mean += delta / n;
            ctv += delta * (x - mean);
        } else {
            const double total = mean * n, delta = x - mean;
            n -= 1.0;
            if (n == 0.0) { mean = 0.0; ctv = 0.0; }
            else { mean = (total - x) / n; ctv -= delta * (x - mean); }
        }
        in[0] = (uint32_t)n; fl[0] = mean; fl[st] = ctv;
        lb_refresh_nich_v(f, (uint32_t)n, mean, ctv, c, st);
        break;
    }
    }
}

// The same update, also handing back what it stored: in_new = the feature's first two integer rows at this group
// after the update, c_new = its cached scorer entries (BB 2, GP 2, NICH 5; DD / DPD: [0] the value's entry, [1] the
// shift entry).  The speculative sweep (lb_sweep_spec.cuh) scores the pending row's cell against these registers
// instead of re-reading the column it just wrote (a second dependent L2 round trip).
__device__ __forceinline__ void lb_update_cell_x(const FeatDev &f, const float *aux, uint32_t *in,
                                                 double *fl, float *c, int st, uint32_t raw, int sign,
                                                 uint32_t (&in_new)[2], float (&c_new)[5]) {
    switch (f.type) {
    case LB_BB: {
        uint32_t h = in[0], t = in[st];
        if (raw) { h += (uint32_t)sign; in[0] = h; } else { t += (uint32_t)sign; in[st] = t; }
        lb_refresh_bb_v(f, h, t, c_new, 1);
        c[0] = c_new[0]; c[st] = c_new[1];
        in_new[0] = h; in_new[1] = t;
        break;
    }
    case LB_DD:
    case LB_DPD: {
        const uint32_t n = in[(size_t)raw * st] + (uint32_t)sign;
        const uint32_t tot = in[(size_t)f.dim * st] + (uint32_t)sign;
        in[(size_t)raw * st] = n;
        in[(size_t)f.dim * st] = tot;
        c_new[0] = logf(f.a_scale * aux[f.aux_off + raw] + (float)n);
        c_new[1] = logf(f.a_total + (float)tot);
        c[(size_t)raw * st] = c_new[0];
        c[(size_t)f.dim * st] = c_new[1];
        in_new[0] = n; in_new[1] = tot;
        break;
    }
    case LB_GP: {
        const uint32_t n0 = in[0], s0 = in[st];
        const double lp0 = fl[0];
        const uint32_t n = n0 + (uint32_t)sign, sum = s0 + (uint32_t)(sign * (int64_t)raw);
        in[0] = n;
        in[st] = sum;
        fl[0] = n ? lp0 + sign * (double)lb_log_factorial_d(raw) : 0.0;
        lb_refresh_gp_v(f, n, sum, c_new, 1);
        c[0] = c_new[0]; c[st] = c_new[1];
        in_new[0] = n; in_new[1] = sum;
        break;
    }
    case LB_NICH: {
        const double x = (double)__uint_as_float(raw);
        double n = (double)in[0], mean = fl[0], ctv = fl[st];
        if (sign > 0) {
            n += 1.0;
            const double delta = x - mean;
            mean += delta / n;
            ctv += delta * (x - mean);
        } else {
            const double total = mean * n, delta = x - mean;
            n -= 1.0;
            if (n == 0.0) { mean = 0.0; ctv = 0.0; }
            else { mean = (total - x) / n; ctv -= delta * (x - mean); }
        }
        in[0] = (uint32_t)n; fl[0] = mean; fl[st] = ctv;
        lb_refresh_nich_v(f, (uint32_t)n, mean, ctv, c_new, 1);
#pragma unroll
        for (int q = 0; q < 5; ++q) c[(size_t)q * st] = c_new[q];
        in_new[0] = (uint32_t)n; in_new[1] = 0u;
        break;
    }
    }
}

// Gamma-Poisson predictive of a count x > 8 in a group with a = alpha + sum >= 8, b1 = inv_beta + count + 1, float32.
// The textbook sum lgamma(a + x) - lgamma(a) - lgamma(x + 1) + a log(b/(b+1)) - x log(1+b) is four terms of size
// x ln(a + x) -- thousands for counts in the hundreds -- that cancel to O(1): 5e-5 of error in float32.  With Stirling's
// series for all three lgammas the large parts combine before they are evaluated (Loader's saddle-point idea):
//   (a - 1/2) log1p(x/a) + a log1p(-1/b1)  =  a log1p(x/a - (1 + x/a)/b1) - log1p(x/a)/2      ~ (x - mean) + ...
//   x log(a + x) - x log x - x log b1       =  x log1p((a + x - b1 x) / (b1 x))                ~ (mean - x) + ...
// and what is left is - log(2 pi x)/2 and the three 1/(12 z) tails.  Every term is O(|x - mean|) or smaller.
// (Used by lb_score_cell: K1, K1d and the sweep's one-column patch.  The sweeps' gather loops still inline the textbook
// sum: this form reads the group's count as well as its sum, one more load per cell there -- DESIGN.md section 5.)
__device__ __forceinline__ float lb_gp_score_stirling(float a, float b1, float x) {
    const float xa = x / a;
    const float t1 = a * log1pf(xa - (1.f + xa) / b1) - 0.5f * log1pf(xa);
    const float bx = b1 * x, r = (a + x) / bx;           // r ~ mean / x
    const float t2 = x * (fabsf(r - 1.f) < 0.5f ? log1pf(((a - bx) + x) / bx) : logf(r));   // near the mean: the difference first
    return t1 + t2 - 0.5f * logf(6.28318531f * x) + (lb_stirling_corr(a + x) - lb_stirling_corr(a)) - lb_stirling_corr(x);
}

// log posterior-predictive of one cell for one group from the cached scorer.
__device__ __forceinline__ float lb_score_cell(const FeatDev &f, const uint32_t *in, const float *c,
                                               int st, uint32_t raw, float cell_const) {
    switch (f.type) {
    case LB_BB:
        return c[raw ? 0 : st];
    case LB_DD:
    case LB_DPD:
        return c[(size_t)raw * st] - c[(size_t)f.dim * st];
    case LB_GP: {
        const float a = f.hp[0] + (float)in[st];
        if (raw > 8u && a >= 8.f) return lb_gp_score_stirling(a, f.hp[1] + (float)in[0] + 1.f, (float)raw);
        return lb_lgamma_diff_int(a, raw) - cell_const + c[0] + (float)raw * c[st];
    }
    case LB_NICH: {
        const float dx = (__uint_as_float(raw) - c[3 * (size_t)st]) - c[4 * (size_t)st];
        return c[0] - c[st] * log1pf(c[2 * (size_t)st] * dx * dx);
    }
    }
    return 0.f;
}
#endif  // __CUDACC__
