// Test harness: the DEVICE'S marginal-likelihood arithmetic (loom_b200/csrc/lb_marginal.cuh: lb_marginal_group, what K5
// k_prop_score and K7 k_hyper_features sum over groups, and lb_lgamma_rise, K5's vectorised Dirichlet path) compiled for
// the host through tests/cpp/cuda_shim/, one log marginal per input line.  See device_math_host.cc for what the host build
// does and does not establish.
//   stdin, per line: type dim hp0 hp1 hp2 hp3  n_aux aux...  n_ints ints...  n_flts flts...
//   stdout: lb_marginal_group, and for DD / DPD also the sum K5's table walk forms with lb_lgamma_rise
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
#include "../../loom_b200/csrc/lb_marginal.cuh"

int main() {
    FeatDev f;
    int n_aux, n_ints, n_flts;
    while (std::scanf("%d %d %f %f %f %f %d", &f.type, &f.dim, &f.hp[0], &f.hp[1], &f.hp[2], &f.hp[3], &n_aux) == 7) {
        std::vector<float> aux(n_aux + 1);
        for (int i = 0; i < n_aux; ++i) if (std::scanf("%f", &aux[i]) != 1) return 2;
        if (std::scanf("%d", &n_ints) != 1) return 2;
        std::vector<uint32_t> ints(n_ints + 1);
        for (int i = 0; i < n_ints; ++i) if (std::scanf("%u", &ints[i]) != 1) return 2;
        if (std::scanf("%d", &n_flts) != 1) return 2;
        std::vector<double> flts(n_flts + 2);
        for (int i = 0; i < n_flts; ++i) if (std::scanf("%lf", &flts[i]) != 1) return 2;
        f.aux_off = 0;
        f.a_total = 0.f;                                  // as lb_sync_model fills them (lb_engine.cu)
        f.a_scale = 1.f;
        if (f.type == LB_DD) { double t = 0; for (int v = 0; v < f.dim; ++v) t += aux[v]; f.a_total = (float)t; }
        if (f.type == LB_DPD) { f.a_total = f.hp[1]; f.a_scale = f.hp[1]; }
        const double group = lb_marginal_group(f, aux.data(), ints.data(), flts.data(), 1);
        double table = group;
        if (f.type == LB_DD || f.type == LB_DPD) {        // k_prop_score's walk: lgamma(a_v) tabulated by k_prop_lgamma_table
            float part = 0.f;
            for (int v = 0; v <= f.dim; ++v) {
                const bool tot = v == f.dim;
                const float a = tot ? f.a_total : f.a_scale * aux[v];
                const float t = ints[v] ? lb_lgamma_rise(a, lgammaf(a), ints[v]) : 0.f;
                part += tot ? -t : t;
            }
            table = (double)part;
        }
        std::printf("%.17g %.17g\n", group, table);
    }
    return 0;
}
