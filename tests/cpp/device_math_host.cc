// Test harness: the DEVICE'S score arithmetic (loom_b200/csrc/lb_common.cuh: lb_refresh_row builds a group's cached
// scorer entries from its sufficient statistics, lb_score_cell scores one cell from them -- float32, what k_cat_sweep*
// and k_score_rows run) compiled for the host through tests/cpp/cuda_shim/, one score per input line.  libm on the host
// is not CUDA's (logf / log1pf / lgammaf agree to an ulp or two, FMA contraction may differ), so this measures the
// ACCURACY of the formulas against the float64 oracle -- for hypers and group sizes the GPU parity tests do not visit --
// not bit patterns.
//   stdin, per line: type dim hp0 hp1 hp2 hp3  n_aux aux...  n_ints ints...  n_flts flts...  raw
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
#include "../../loom_b200/csrc/lb_common.cuh"

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
        uint32_t raw;
        if (std::scanf("%u", &raw) != 1) return 2;
        f.aux_off = 0;
        f.a_total = 0.f;                                  // as lb_sync_model fills them (lb_engine.cu)
        f.a_scale = 1.f;
        if (f.type == LB_DD) for (int v = 0; v < f.dim; ++v) f.a_total += aux[v];
        if (f.type == LB_DPD) { f.a_total = f.hp[1]; f.a_scale = f.hp[1]; }
        const int n_cache = f.type == LB_BB ? 2 : f.type == LB_GP ? 2 : f.type == LB_NICH ? 5 : f.dim + 1;
        std::vector<float> cache(n_cache + 1, 0.f);
        for (int local = 0; local < n_cache; ++local) lb_refresh_row(f, aux.data(), ints.data(), flts.data(), cache.data(), 1, local);
        const float cell_const = f.type == LB_GP ? lb_log_factorial(raw) : 0.f;
        std::printf("%.9g\n", lb_score_cell(f, ints.data(), cache.data(), 1, raw, cell_const));
    }
    return 0;
}
