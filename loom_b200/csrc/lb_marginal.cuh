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

// same closed forms as lb_marginal in lb_cat.cu (kept local: no relocatable device code)
__device__ double lb_marginal_group(const FeatDev &f, const float *aux, const uint32_t *in, const double *fl, int st) {
    switch (f.type) {
    case LB_BB: {
        const float h = (float)in[0], t = (float)in[st];
        return (double)lb_lgamma_diff(f.hp[0], h) + (double)lb_lgamma_diff(f.hp[1], t) -
               (double)lb_lgamma_diff(f.hp[0] + f.hp[1], h + t);
    }
    case LB_DD:
    case LB_DPD: {
        double s = 0;
        for (int v = 0; v < f.dim; ++v) {
            const uint32_t n = in[(size_t)v * st];
            if (n) s += (double)lb_lgamma_diff(f.a_scale * aux[f.aux_off + v], (float)n);
        }
        return s - (double)lb_lgamma_diff(f.a_total, (float)in[(size_t)f.dim * st]);
    }
    case LB_GP: {
        const double al = f.hp[0], ib = f.hp[1];
        const double S = (double)in[st], n = (double)in[0];
        return (double)lb_lgamma_diff(f.hp[0], (float)in[st]) + al * log(ib) - (al + S) * log(ib + n) - fl[0];
    }
    case LB_NICH: {
        const double mu = f.hp[0], kappa = f.hp[1], sigmasq = f.hp[2], nu = f.hp[3];
        const double n = (double)in[0], mean = fl[0], ctv = fl[st];
        const double kn = kappa + n, nun = nu + n, dm = mu - mean;
        const double sn = (nu * sigmasq + ctv + n * kappa * dm * dm / kn) / nun;
        return (double)lb_lgamma_diff((float)(0.5 * nu), (float)(0.5 * n)) + 0.5 * log(kappa / kn) +
               0.5 * nu * log(nu * sigmasq) - 0.5 * nun * log(nun * sn) - 0.5 * n * 1.1447298858494002;
    }
    }
    return 0.0;
}
