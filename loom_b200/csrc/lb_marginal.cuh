// lb_marginal.cuh -- log marginal likelihood of one group's sufficient statistics (what K5 `k_prop_score` and K7
// `k_hyper_features` sum over groups; src/product_mixture.cc:594-620 -> Group::score_data of distributions).  Lives in a
// header of its own so that tests/test_device_math.py can compile exactly this arithmetic for the host and compare it
// with the float64 oracle over hypers and group sizes the GPU parity tests do not visit.  Included by lb_struct.cu only.
#pragma once
#include "lb_common.cuh"

// lgamma(a + n) - lgamma(a) for an integer count n >= 1 given lga = lgamma(a): short rising factorials as one log of
// a product, everything else the Stirling series of lgamma(a + n) (a + n >= 8 there, since n > 8)
__device__ __forceinline__ float lb_lgamma_rise(float a, float lga, uint32_t n) {
    if (n <= 8u) {
        float p1 = a, p2 = 1.f;
        if (n > 1u) p1 *= a + 1.f;
        if (n > 2u) p1 *= a + 2.f;
        if (n > 3u) p1 *= a + 3.f;
        if (n > 4u) p2 = a + 4.f;
        if (n > 5u) p2 *= a + 5.f;
        if (n > 6u) p2 *= a + 6.f;
        if (n > 7u) p2 *= a + 7.f;
        return logf(p1) + logf(p2);
    }
    const float z = a + (float)n;
    if (a >= 8.f) return lb_lgamma_diff(a, (float)n);     // both large: the cancellation-free form
    return (z - 0.5f) * logf(z) - z + 0.918938533f + lb_stirling_corr(z) - lga;
}

// lgamma(a + x) - lgamma(a) in float64, for the counts of LARGE groups: the float32 form is exact to an ulp of its own
// size, x ln x -- 0.8 nat at x = 10^6, 140 nats at 10^8 (a Gamma-Poisson group's sum of values) -- and a log marginal is a
// handful of such differences that cancel to a few nats per row (tests/test_device_math.py).  Stirling's series on both
// sides when a >= 8 (no cancellation), lgamma(a) itself below.
__device__ __forceinline__ double lb_lgamma_diff_big(double a, double x) {
    const double b = a + x, rb = 1.0 / b, rb2 = rb * rb;
    const double cb = rb * (0.08333333333333333 - rb2 * (0.002777777777777778 - rb2 * 0.0007936507936507937));
    if (a >= 8.0) {
        const double ra = 1.0 / a, ra2 = ra * ra;
        const double ca = ra * (0.08333333333333333 - ra2 * (0.002777777777777778 - ra2 * 0.0007936507936507937));
        return (a - 0.5) * log1p(x / a) + x * (log(b) - 1.0) + (cb - ca);
    }
    return (b - 0.5) * log(b) - b + 0.9189385332046727 + cb - lb_lgamma_d(a);
}
#define LB_BIG_COUNT 4096.f      // above this a float32 lgamma difference is off by more than ~2e-3 nat
__device__ __forceinline__ double lb_lgamma_diff_any(float a, float x) {
    return x > LB_BIG_COUNT ? lb_lgamma_diff_big((double)a, (double)x) : (double)lb_lgamma_diff(a, x);
}

// same closed forms as lb_marginal in lb_cat.cu (kept local: no relocatable device code)
__device__ double lb_marginal_group(const FeatDev &f, const float *aux, const uint32_t *in, const double *fl, int st) {
    switch (f.type) {
    case LB_BB: {
        const float h = (float)in[0], t = (float)in[st];
        return lb_lgamma_diff_any(f.hp[0], h) + lb_lgamma_diff_any(f.hp[1], t) - lb_lgamma_diff_any(f.hp[0] + f.hp[1], h + t);
    }
    case LB_DD:
    case LB_DPD: {
        double s = 0;
        for (int v = 0; v < f.dim; ++v) {
            const uint32_t n = in[(size_t)v * st];
            if (n) s += lb_lgamma_diff_any(f.a_scale * aux[f.aux_off + v], (float)n);
        }
        return s - lb_lgamma_diff_any(f.a_total, (float)in[(size_t)f.dim * st]);
    }
    case LB_GP: {
        const double al = f.hp[0], ib = f.hp[1];
        const double S = (double)in[st], n = (double)in[0];
        const double rise = S > (double)LB_BIG_COUNT ? lb_lgamma_diff_big(al, S) : (double)lb_lgamma_diff(f.hp[0], (float)in[st]);
        return rise + al * log(ib) - (al + S) * log(ib + n) - fl[0];
    }
    case LB_NICH: {
        const double mu = f.hp[0], kappa = f.hp[1], sigmasq = f.hp[2], nu = f.hp[3];
        const double n = (double)in[0], mean = fl[0], ctv = fl[st];
        const double kn = kappa + n, nun = nu + n, dm = mu - mean;
        const double sn = (nu * sigmasq + ctv + n * kappa * dm * dm / kn) / nun;
        const double rise = n > 2.0 * (double)LB_BIG_COUNT ? lb_lgamma_diff_big(0.5 * nu, 0.5 * n)
                                                           : (double)lb_lgamma_diff((float)(0.5 * nu), (float)(0.5 * n));
        return rise + 0.5 * log(kappa / kn) +
               0.5 * nu * log(nu * sigmasq) - 0.5 * nun * log(nun * sn) - 0.5 * n * 1.1447298858494002;
    }
    }
    return 0.0;
}
