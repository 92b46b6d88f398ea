/*
 * loom_oracle.h -- CPU oracle for the loom hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * PARITY UNPINNED against the reference binary: priorknowledge/loom cannot be
 * built here (its arithmetic lives in forcedotcom/distributions 2.0.28, which is
 * not vendored) and ships no golden vectors.  This file
 * restates (a) loom's control flow from the cited reference files and (b) the
 * published conjugate closed forms distributions implements, in float64 with
 * exact libm.  It is pinned instead against scipy known-answer tests and loom's
 * own posterior-enumeration chi-square test (tests/test_oracle_*.py).
 * The pieces of the reference that DO compile here (Makefile target `ref`: the
 * schedules, kind_proposer.cc's block Pitman-Yor sampler, kind_kernel.cc's
 * bookkeeping, differ.cc + product_value.cc, and -- over stand-in mixtures
 * whose per-feature arithmetic is this library's -- product_mixture.cc, a whole
 * kind-structure pass of cross_cat.cc + kind_kernel.cc, hyper_kernel.cc) are
 * pinned against the
 * reference's own output:
 * tests/golden/reference_*.json, tests/test_reference_pins.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product (libloomb200.so)
 * never links or calls it.
 *
 * It exports the same C ABI as include/loom_b200.h with the prefix lo_.
 */
#ifndef LOOM_ORACLE_H_
#define LOOM_ORACLE_H_

#define LB_NAME(x) lo_##x
#include "../include/loom_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Oracle-only helpers used by the tests. */

/* Philox4x32-10 uniform, exposed so tests can pin the generator against the
 * published Random123 known-answer vectors. */
void lo_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
double lo_uniform(uint64_t seed, uint32_t stream, uint64_t a, uint32_t b);

/* Single-feature closed forms (float64) for scipy known-answer tests:
 * log posterior-predictive of `value` and log marginal of a group with the
 * given sufficient statistics.  ints/flts use the per-feature row order of
 * loom_b200.h. */
double lo_feature_score_value(const lb_feature_spec *spec, const float *aux,
                              const uint32_t *ints, const double *flts, uint32_t raw_value);
double lo_feature_score_data(const lb_feature_spec *spec, const float *aux,
                             const uint32_t *ints, const double *flts);
double lo_py_score_counts(double alpha, double d, const uint32_t *counts, int n);

/* Replace the score matrix L[f][k] ([F][K], row f = feature f) that kind_proposer_score computed, keeping the
 * proposer's tables: lets tests/test_reference_pins.py drive kind_run + init_featureless_kinds with the score
 * matrices the compiled reference KindKernel was driven with (oracle/ref_kind_kernel_main.cc).  Checker-only:
 * the product library has no such entry point. */
int lo_set_proposer_scores(lb_handle h, const float *scores);

/* Number of OpenMP threads the kind-parallel sweep / feature-parallel scoring
 * may use (reference threading model: one thread per kind,
 * src/cat_pipeline.cc:103-125; OMP over features src/kind_proposer.cc:339). */
void lo_set_threads(int n);

/* Replay check for the sequential Gibbs kernel: replays `actions` on the
 * oracle's own state but takes the group picked by the device at each ADD
 * (picks[i*K + k] = packed id) instead of sampling, and verifies each pick is
 * consistent with the shared Philox uniform and the oracle's float64 CDF up to
 * `tol` (in probability mass).  Returns the number of inconsistent picks, or
 * <0 on error.  max_score_err receives the largest |device - oracle| score
 * difference (scaled by 1/max(1,|oracle|)) when debug_scores != NULL. */
int64_t lo_cat_replay(lb_handle h, const uint32_t *actions, uint64_t n_actions, uint64_t seed,
                      uint64_t action_offset, const uint32_t *picks,
                      const float *debug_scores, int32_t debug_stride, double tol,
                      double *max_score_err);

#ifdef __cplusplus
}
#endif
#endif
