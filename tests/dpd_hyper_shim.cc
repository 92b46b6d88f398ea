// Test shim: the PRODUCT's host-side DPD hyper sampler (loom_b200/csrc/lb_dpd_hyper.h, the code
// lb_hyper_run calls on the GPU box) compiled for a CPU box, so tests/test_dpd.py can compare it with
// the oracle's independent restatement draw for draw.  Uniforms are the oracle's Philox
// (lo_uniform) -- the product's device/host Philox is checked against it by the GPU sweep replays.
#include <cstdint>

#include "../loom_b200/csrc/lb_dpd_hyper.h"

extern "C" double lo_uniform(uint64_t seed, uint32_t stream, uint64_t a, uint32_t b);

extern "C" uint32_t dpd_hyper_shim(const uint32_t *counts, int dim, int G, uint64_t stride, float *hp, float *betas,
                                   const float *gamma_grid, int n_gamma, const float *alpha_grid, int n_alpha,
                                   uint64_t seed, uint64_t task) {
    return lb_dpd::hyper(counts, dim, G, (size_t)stride, hp, betas, gamma_grid, n_gamma, alpha_grid, n_alpha,
                         [&](uint32_t i) { return lo_uniform(seed, 2u, task, i); });
}
