// lb_dpd_hyper.h -- hyper-parameter sampler of a DirichletProcessDiscrete feature: the DPD branch of
// HyperKernel::infer_feature_hypers_fun (src/hyper_kernel.cc:90-181).  Host code, float64; the
// feature's count table is a few KB, so lb_hyper_run downloads it and runs this (the other feature
// types run in k_hyper_features on the device).  No CUDA in this header: tests compile it on a CPU
// box and compare it with the oracle's restatement draw for draw.
//
// Steps, in the reference's order:
//  1. auxiliary table counts (hyper_kernel.cc:101-122): for every (group, value) cell with count n,
//     m ~ P(m) proportional to |s(n, m)| (alpha beta_v)^m, summed over groups per value.  The
//     reference draws m from a row of log Stirling numbers (distributions::get_log_stirling1_row,
//     not vendored); the SAME distribution is the number of tables n customers open in a CRP with
//     concentration a = alpha beta_v, m = 1 + sum_{i=1}^{n-1} Bernoulli(a / (a + i)) -- exact, no table.
//  2. only if every dictionary value has been observed (hyper_kernel.cc:124):
//     a. gamma: grid Gibbs on  V log gamma + lgamma(gamma) - lgamma(gamma + sum aux)   (:127-141)
//     b. (betas, beta0) ~ Dirichlet(aux_1 .. aux_V, gamma), floored at MIN_BETA        (:143-167)
//     c. alpha: grid Gibbs on the data score of all groups (InferShared::done)          (:169-176)
//     Like the reference, b and c only run when the alpha grid is non-empty.
// Uniforms come from the caller: uniform(i) is the i-th draw of this feature's hyper task (Philox
// stream 2, counter a = task id, b = i; include/loom_b200.h).
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

// distributions' DirichletProcessDiscrete::MIN_BETA() is not in the reference tree; stand-in value.
#define LB_DPD_MIN_BETA 1e-6

namespace lb_dpd {

inline int sample_from_scores(double u, std::vector<double> &s) {
    double m = -INFINITY, total = 0;
    for (double v : s) m = v > m ? v : m;
    for (double &v : s) { v = std::exp(v - m); total += v; }
    const double t = u * total;
    double c = 0;
    int last = 0;
    for (int i = 0; i < (int)s.size(); ++i) {
        if (s[i] > 0) last = i;
        c += s[i];
        if (c > t) return i;
    }
    return last;
}

// Marsaglia-Tsang gamma variate (scale 1) from the caller's uniforms
template <class Uniform> double gamma_variate(double shape, Uniform &uniform, uint32_t &draw) {
    if (shape < 1.0) {
        const double u = uniform(draw++);
        return gamma_variate(shape + 1.0, uniform, draw) * std::pow(u, 1.0 / shape);
    }
    const double d = shape - 1.0 / 3.0, c = 1.0 / std::sqrt(9.0 * d);
    for (;;) {
        const double u1 = uniform(draw++), u2 = uniform(draw++);
        const double z = std::sqrt(-2.0 * std::log(1.0 - u1)) * std::cos(6.283185307179586 * u2);
        double v = 1.0 + c * z;
        if (v <= 0) continue;
        v = v * v * v;
        const double u = uniform(draw++);
        if (std::log(1.0 - u) < 0.5 * z * z + d - d * v + d * std::log(v)) return d * v;
    }
}

// counts: [dim + 1][stride] (value rows then the count_sum row), G groups.
// hp = {gamma, alpha, beta0}; betas[dim].  Returns the number of uniforms consumed.
template <class Uniform>
uint32_t hyper(const uint32_t *counts, int dim, int G, size_t stride, float *hp, float *betas, const float *gamma_grid,
               int n_gamma, const float *alpha_grid, int n_alpha, Uniform &&uniform) {
    uint32_t draw = 0;
    if (dim <= 0) return draw;
    std::vector<uint64_t> aux((size_t)dim, 0);
    const double alpha0 = hp[1];
    for (int g = 0; g < G; ++g)
        for (int v = 0; v < dim; ++v) {
            const uint32_t n = counts[(size_t)v * stride + g];
            if (!n) continue;
            const double a = alpha0 * (double)betas[v];
            uint64_t m = 1;
            for (uint32_t i = 1; i < n; ++i) {
                const double u = uniform(draw++);
                m += u * (a + (double)i) < a;
            }
            aux[v] += m;
        }
    int seen = 0;
    uint64_t aux_total = 0;
    for (int v = 0; v < dim; ++v) { seen += aux[v] > 0; aux_total += aux[v]; }
    if (seen != dim) return draw;
    std::vector<double> scores;
    if (n_gamma > 0) {
        scores.resize(n_gamma);
        for (int j = 0; j < n_gamma; ++j) {
            const double gm = gamma_grid[j];
            scores[j] = dim * std::log(gm) + std::lgamma(gm) - std::lgamma(gm + (double)aux_total);
        }
        hp[0] = gamma_grid[sample_from_scores(uniform(draw++), scores)];
    }
    if (n_alpha > 0) {
        std::vector<double> x((size_t)dim + 1);
        double total = 0;
        for (int i = 0; i <= dim; ++i) {
            x[i] = gamma_variate(i < dim ? (double)aux[i] : (double)hp[0], uniform, draw);
            total += x[i];
        }
        double floored = 0;
        for (int i = 0; i <= dim; ++i) { x[i] = std::fmax(x[i] / total, LB_DPD_MIN_BETA); floored += x[i]; }
        for (int v = 0; v < dim; ++v) betas[v] = (float)(x[v] / floored);
        hp[2] = (float)(x[dim] / floored);
        scores.resize(n_alpha);
        for (int j = 0; j < n_alpha; ++j) {
            const double al = alpha_grid[j];
            double score = 0;
            for (int g = 0; g < G; ++g) {
                const uint32_t total_g = counts[(size_t)dim * stride + g];
                if (!total_g) continue;
                for (int v = 0; v < dim; ++v) {
                    const uint32_t n = counts[(size_t)v * stride + g];
                    if (n) score += std::lgamma(al * (double)betas[v] + (double)n) - std::lgamma(al * (double)betas[v]);
                }
                score += std::lgamma(al) - std::lgamma(al + (double)total_g);
            }
            scores[j] = score;
        }
        hp[1] = alpha_grid[sample_from_scores(uniform(draw++), scores)];
    }
    return draw;
}

}  // namespace lb_dpd
