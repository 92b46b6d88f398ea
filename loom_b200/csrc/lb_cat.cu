// lb_cat.cu -- cat-kernel CUDA kernels for sm_100a:
//   k_cat_sweep      exact sequential single-site Gibbs, one CTA per kind
//                    (CatKernel::process_add_task / process_remove_task,
//                     src/cat_kernel.hpp:199-299; one thread per kind in the
//                     reference, src/cat_pipeline.cc:103-125)
//   k_score_rows     batch ProductMixture::score_value with frozen tables
//                    (src/product_mixture.cc:421-434)
//   k_accumulate     bulk add_value / remove_value for given assignments
//                    (src/product_mixture.cc:165-239)
//   + cached-scorer refresh, table restride, id renumbering, row validation,
//     score_data (src/cross_cat.cc:238-250).
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstring>

#include "lb_accum.cuh"
#include <cub/cub.cuh>

#include "lb_internal.h"

#define FULL 0xffffffffu

// ===========================================================================
// K3: sequential Gibbs sweep.
//
// One CTA per kind walks the action list; the dependent chain per action is
//   ADD    score (all warps) -> bar -> sample (warp 0) -> bar -> update (owner warps)
//   REMOVE lookup (thread 0) -> bar -> update (owner warps)
// Features are split into three classes (categorical = BB/DD/DPD gathers,
// GP, NICH) and every warp owns a fixed, class-pure stripe of features for
// BOTH scoring and updating, so a warp's own table writes are ordered before
// its next reads without a third block barrier, and no warp diverges over
// feature types.  Row cells are prefetched two actions ahead into a 4-slot
// shared-memory ring; the Philox uniform of the next ADD is drawn by an idle
// warp while warp 0 samples.
// ===========================================================================
#ifndef LB_W1_CTAS
#define LB_W1_CTAS 16    // resident one-warp CTAs per SM the register budget is set for
#endif
#ifndef LB_CPB
#define LB_CPB 8
#endif
constexpr int SW_CPB = LB_CPB;    // chains per block when a chain's CTA is a single warp
constexpr int SW_MAX_WARPS = 8;   // warps per (chain, kind) CTA: 8 for latency, fewer when many chains fill the GPU
constexpr int SW_PF = 4;          // row cells prefetched per thread -> nf <= 128 * warps per kind
constexpr int SW_MAX_G = 4096;
enum { CL_CAT = 0, CL_GP = 1, CL_NICH = 2, N_CLASS = 3 };

struct SweepSmem {
    int G, stop;
    int pick[2], flag[2];          // indexed by action parity: no end-of-action barrier
    uint32_t removed_global[2], next_global;
    unsigned long long sample_size;
    float u[2];
    uint32_t pf_assign[4];         // prefetched assignment of the row of each ring slot
    uint32_t pf_pg[4], pf_epoch[4]; // ... its packed group id, valid while no group was swap-removed since
    uint32_t epoch;                // number of swap-removes so far
    int own_off[SW_MAX_WARPS][N_CLASS + 1];   // warp w owns s_own[own_off[w][c] .. own_off[w][c+1]) of class c
};

__device__ __forceinline__ float warp_max(float v) {
    // order-preserving float -> int map, then one redux.sync
    int i = __float_as_int(v);
    i ^= (i >> 31) & 0x7fffffff;
    i = __reduce_max_sync(FULL, i);
    i ^= (i >> 31) & 0x7fffffff;
    return __int_as_float(i);
}

// Draw from softmax(scores[0..G)) with uniform u, CDF in index order
// (distributions::sample_from_scores_overwrite; SURVEY.md 8c "Sampling").
// Executed by one full warp; scores are overwritten with exp(score - max).
__device__ __forceinline__ int warp_sample(float *scores, int G, float u, int lane) {
    float m = -INFINITY;
    for (int g = lane; g < G; g += 32) m = fmaxf(m, scores[g]);
    m = warp_max(m);
    const int per = (G + 31) >> 5;
    const int lo = lane * per, hi = min(G, lo + per);
    float local = 0.f;
    for (int g = lo; g < hi; ++g) {
        const float p = __expf(scores[g] - m);
        scores[g] = p;
        local += p;
    }
    float incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float v = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += v;
    }
    const float total = __shfl_sync(FULL, incl, 31);
    const float t = u * total;
    // Only lanes holding probability mass may be chosen: the scan's rounding can give a
    // lane with an empty group range an inclusive sum one ulp above its predecessor's.
    const unsigned nonzero = __ballot_sync(FULL, local > 0.f);
    const unsigned ballot = __ballot_sync(FULL, incl > t) & nonzero;
    const int src = ballot ? __ffs(ballot) - 1 : 31 - __clz(nonzero);
    int pick = lo;
    if (lane == src) {
        float c = incl - local;
        int last = lo;
        pick = -1;
        for (int g = lo; g < hi; ++g) {
            const float p = scores[g];
            if (p > 0.f) last = g;
            c += p;
            if (c > t) { pick = g; break; }
        }
        if (pick < 0) pick = last;
    }
    return __shfl_sync(FULL, pick, src);
}

__device__ __forceinline__ int feat_class(int type) {
    return type == LB_GP ? CL_GP : (type == LB_NICH ? CL_NICH : CL_CAT);
}

// add the scores of one warp's features of class `cls` (own[0..n)) for groups g0 + lane + 32*jj
template <int NJ>
__device__ __forceinline__ void score_list(int cls, const uint16_t *own, int n, const FeatDev *s_feat,
                                           const uint32_t *c_raw, const uint32_t *c_obs,
                                           const uint32_t *ints, const float *cache, int st,
                                           int zero_row, int g0, int lane, float *acc) {
    if (cls == CL_CAT) {
        // branch-free gathers: unobserved cells read the all-zero cache row
#pragma unroll 4
        for (int k = 0; k < n; ++k) {
            const int j = own[k];
            const FeatDev &f = s_feat[j];
            const uint32_t raw = c_raw[j];
            const bool obs = c_obs[j] != 0u;
            const bool bb = f.type == LB_BB;
            const int vrow = obs ? f.cache_row + (bb ? (raw ? 0 : 1) : (int)raw) : zero_row;
            const int srow = (obs && !bb) ? f.cache_row + f.dim : zero_row;
            const float *cv = cache + (size_t)vrow * st + g0 + lane;
            const float *cs = cache + (size_t)srow * st + g0 + lane;
#pragma unroll
            for (int jj = 0; jj < NJ; ++jj) acc[jj] += cv[32 * jj] - cs[32 * jj];
        }
    } else if (cls == CL_GP) {
        for (int k = 0; k < n; ++k) {
            const int j = own[k];
            if (!c_obs[j]) continue;
            const FeatDev &f = s_feat[j];
            const uint32_t raw = c_raw[j];
            const float lf = lb_log_factorial(raw), x = (float)raw;
            const uint32_t *sum = ints + (size_t)(f.int_row + 1) * st + g0 + lane;
            const float *c0 = cache + (size_t)f.cache_row * st + g0 + lane;
#pragma unroll
            for (int jj = 0; jj < NJ; ++jj) {
                const float a = f.hp[0] + (float)sum[32 * jj];
                acc[jj] += lb_lgamma_diff_int(a, raw) - lf + c0[32 * jj] + x * c0[st + 32 * jj];
            }
        }
    } else {
        for (int k = 0; k < n; ++k) {
            const int j = own[k];
            if (!c_obs[j]) continue;
            const FeatDev &f = s_feat[j];
            const float x = __uint_as_float(c_raw[j]);
            const float *c = cache + (size_t)f.cache_row * st + g0 + lane;
#pragma unroll
            for (int jj = 0; jj < NJ; ++jj) {
                const float dx = (x - c[3 * (size_t)st + 32 * jj]) - c[4 * (size_t)st + 32 * jj];
                acc[jj] += c[32 * jj] - c[st + 32 * jj] * log1pf(c[2 * (size_t)st + 32 * jj] * dx * dx);
            }
        }
    }
}
// all of one warp's features (off[c] .. off[c+1] per class)
template <int NJ>
__device__ __forceinline__ void score_warp(const int *off, const uint16_t *s_own, const FeatDev *s_feat,
                                           const uint32_t *c_raw, const uint32_t *c_obs,
                                           const uint32_t *ints, const float *cache, int st,
                                           int zero_row, int g0, int lane, float *acc) {
#pragma unroll
    for (int jj = 0; jj < NJ; ++jj) acc[jj] = 0.f;
#pragma unroll
    for (int c = 0; c < N_CLASS; ++c)
        if (off[c + 1] > off[c])
            score_list<NJ>(c, s_own + off[c], off[c + 1] - off[c], s_feat, c_raw, c_obs, ints, cache, st, zero_row, g0, lane, acc);
}

// W warps per CTA; registers are capped at 128 per thread for every W
#ifndef LB_LOCKSTEP
#define LB_LOCKSTEP 1
#endif
#ifndef LB_LOCKSTEP_EVERY
#define LB_LOCKSTEP_EVERY 1   // actions between two re-alignments (power of two)
#endif
// W == 1: the warp is the whole "CTA" of its chain, so a warp-level sync is all correctness needs.
// LB_LOCKSTEP >= 2 makes these points block-wide too: pure re-alignment of the block's chains.
// (A chain that stops early simply exits; barriers only wait for warps that have not exited.)
template <int W> __device__ __forceinline__ void sweep_sync() {   // rare paths, start-up
    if (W == 1) __syncwarp(); else __syncthreads();
}
template <int W> __device__ __forceinline__ void phase_sync() {   // the (A) / (B) points of every action
    if (W == 1 && LB_LOCKSTEP < 2) __syncwarp(); else __syncthreads();
}

// W == 1: a block is SW_CPB independent one-warp "CTAs" -- SW_CPB chains of the SAME kind.  They
// never synchronise with each other, but they run the same code over the same action list, so
// they stay close to lockstep and share instruction-cache lines; 16 unrelated one-warp CTAs per
// SM were instruction-fetch-bound (ncu: stall_no_instruction on top).
template <int W>
__global__ void __launch_bounds__(W == 1 ? 32 * SW_CPB : W * 32, W == 1 ? LB_W1_CTAS / SW_CPB : 16 / W)
k_cat_sweep(SweepChain one, const SweepChain *many, int n_chains, int smem_per_chain) {
    constexpr int SW_THREADS = W * 32, SW_WARPS = W;
    extern __shared__ __align__(16) unsigned char smem_all[];
    // blockIdx.y = kind slot (kind_ids are sorted most expensive first and y is the slow grid
    // index, so the long CTAs are dispatched first and the cheap ones fill the tail);
    // chain (independent CrossCat states swept with the same action list) = blockIdx.x, or
    // blockIdx.x * SW_CPB + warp-in-block when W == 1
    const int sub = W == 1 ? (int)(threadIdx.x >> 5) : 0;
    const int chain = W == 1 ? (int)blockIdx.x * SW_CPB + sub : (int)blockIdx.x;
    if (chain >= n_chains) return;
    unsigned char *const smem_raw = smem_all + (size_t)sub * smem_per_chain;
    const SweepChain &ch = many ? many[chain] : one;
    if ((int)blockIdx.y >= ch.n_ids) return;
    KindDev *const kinds = ch.kinds;
    const FeatDev *const feats = ch.feats;
    const int32_t *const kind_feats = ch.kind_feats;
    const float *const aux = ch.aux;
    const RowsDev rows = ch.rows;
    const uint32_t *const actions = ch.actions;
    const uint64_t n_actions = ch.n_actions, seed = ch.seed, offset = ch.offset;
    const int K = ch.K, n_empty = ch.n_empty, dbg_stride = ch.dbg_stride;
    uint32_t *const dbg_picks = ch.dbg_picks;
    float *const dbg_scores = ch.dbg_scores;
    unsigned long long *const prof = ch.prof;
    const int kid = ch.kind_ids[blockIdx.y];
    KindDev &kd = kinds[kid];
    const int tid = W == 1 ? (int)(threadIdx.x & 31) : (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // LB_PROFILE=1: cycles per phase seen by lane 0 of every warp
    // ADD: 0 score | 1 wait at (A) | 2 sample / commit | 3 wait at (B) | 4 update;  REMOVE: 5 lookup | 6 wait at (A) | 7 update
    // (profiling build only, make EXTRA=-DLB_SWEEP_PROFILE: the counters cost 18 registers)
#ifdef LB_SWEEP_PROFILE
    long long pt[8] = {0, 0, 0, 0, 0, 0, 0, 0}, pc = 0;
#define LB_TICK(k) do { if (prof && lane == 0) { const long long now__ = clock64(); pt[k] += now__ - pc; pc = now__; } } while (0)
#else
#define LB_TICK(k) do { } while (0)
#endif
    const int nf = kd.nf, st = kd.Gcap;
    const int n_int_rows = kd.n_int_rows, n_flt_rows = kd.n_flt_rows, n_cache_rows = kd.n_cache_rows;
    const int zero_row = n_cache_rows;
    const uint32_t g2p_cap = kd.g2p_cap;

    // carve shared memory
    SweepSmem *ss = (SweepSmem *)smem_raw;
    FeatDev *s_feat = (FeatDev *)(smem_raw + 256);
    int32_t *s_fid = (int32_t *)(s_feat + nf);
    uint32_t *s_raw = (uint32_t *)(s_fid + nf);          // [4][nf]
    uint32_t *s_obs = s_raw + 4 * nf;                    // [4][nf]
    float *s_score = (float *)(s_obs + 4 * nf);          // [st]
    float *s_part = s_score + st;                        // [SW_WARPS][st]
    uint32_t *s_counts = (uint32_t *)(s_part + SW_WARPS * st);   // [st] clustering counts
    float *s_prior = (float *)(s_counts + st);           // [st] CRP prior per group, empties included
    uint16_t *s_own = (uint16_t *)(s_prior + st);        // [nf] feature indices grouped by (owner warp, class)
    uint8_t *s_owner = (uint8_t *)(s_own + nf);          // [nf]

    uint32_t *const ints = kd.ints;
    double *const flts = kd.flts;
    float *const cache = kd.cache;
    uint32_t *const assign = kd.assign;
    uint32_t *const p2g = kd.p2g;
    uint32_t *const g2p = kd.g2p;
    const int32_t *const cache_row_feat = kd.cache_row_feat;
    const float alpha = kd.alpha, d = kd.d;

    for (int j = tid; j < nf; j += SW_THREADS) {
        const int fid = kind_feats[kd.feat_begin + j];
        s_feat[j] = feats[fid];
        s_fid[j] = fid;
    }
    {   // clustering state lives in shared memory for the whole launch (written back at the end)
        const int G0 = kd.G;
        const float sh_empty0 = logf((alpha + d * (float)(G0 - n_empty)) / (float)n_empty);
        for (int g = tid; g < st; g += SW_THREADS) {
            const uint32_t n = g < G0 ? kd.counts[g] : 0u;
            s_counts[g] = n;
            s_prior[g] = n ? logf((float)n - d) : sh_empty0;
        }
    }
    sweep_sync<W>();
    if (tid == 0) {
        ss->G = kd.G; ss->stop = 0; ss->epoch = 0;
        ss->next_global = kd.next_global; ss->sample_size = kd.sample_size;
        // Features -> warps by cost (GP 4, NICH 3, categorical 1).  A warp owns its features for
        // BOTH scoring and update; its list is sorted by class so the scoring loops stay class-pure.
        // With few warps: longest-processing-time first, warp 0 (the sampler) handicapped.
        const float weight[N_CLASS] = {1.f, 4.f, 3.f};
        int cnt[N_CLASS] = {0, 0, 0}, n_cls = 0;
        for (int j = 0; j < nf; ++j) cnt[feat_class(s_feat[j].type)]++;
        for (int c = 0; c < N_CLASS; ++c) n_cls += cnt[c] > 0;
        if (SW_WARPS >= 2 * n_cls) {
            // enough warps for class-pure ownership: a warp then runs ONE scoring loop, i.e. one
            // round trip of dependent loads per action instead of one per class.  Warps per class
            // follow weight x count; the sampler (warp 0) takes the cheapest class.
            int nw[N_CLASS] = {0, 0, 0}, left = SW_WARPS;
            for (int c = 0; c < N_CLASS; ++c) if (cnt[c]) { nw[c] = 1; --left; }
            while (left > 0) {
                int best = -1; float worst = 0.f;
                for (int c = 0; c < N_CLASS; ++c) {
                    if (!cnt[c] || nw[c] >= cnt[c]) continue;
                    const float l = weight[c] * (float)cnt[c] / (float)nw[c];
                    if (l > worst) { worst = l; best = c; }
                }
                if (best < 0) break;
                nw[best]++; --left;
            }
            int w0 = 0, seen[N_CLASS] = {0, 0, 0};
            int first_warp[N_CLASS];
            for (int c = 0; c < N_CLASS; ++c) { first_warp[c] = w0; w0 += nw[c]; }
            for (int j = 0; j < nf; ++j) {
                const int c = feat_class(s_feat[j].type);
                s_owner[j] = (uint8_t)(first_warp[c] + seen[c]++ % nw[c]);
            }
        } else {
            float load[SW_WARPS];
            for (int w = 0; w < SW_WARPS; ++w) load[w] = 0.f;
            if (SW_WARPS > 1) load[0] = 6.f;
            const int order[N_CLASS] = {CL_GP, CL_NICH, CL_CAT};
            for (int oc = 0; oc < N_CLASS; ++oc)
                for (int j = 0; j < nf; ++j) {
                    if (feat_class(s_feat[j].type) != order[oc]) continue;
                    int best = 0;
                    for (int w = 1; w < SW_WARPS; ++w) if (load[w] < load[best]) best = w;
                    s_owner[j] = (uint8_t)best;
                    load[best] += weight[order[oc]];
                }
        }
        int k = 0;
        for (int w = 0; w < SW_WARPS; ++w) {
            for (int c = 0; c < N_CLASS; ++c) {
                ss->own_off[w][c] = k;
                for (int j = 0; j < nf; ++j)
                    if (s_owner[j] == w && feat_class(s_feat[j].type) == c) s_own[k++] = (uint16_t)j;
            }
            ss->own_off[w][N_CLASS] = k;
        }
    }
    sweep_sync<W>();
    int off[N_CLASS + 1];
#pragma unroll
    for (int c = 0; c <= N_CLASS; ++c) off[c] = ss->own_off[warp][c];

    uint64_t i = kd.progress;
    // row cells of action `idx`: global -> registers, later registers -> ring slot
    uint32_t pf_raw[SW_PF], pf_obs[SW_PF], pf_assign = LB_UNASSIGNED, pf_pg = LB_UNASSIGNED, pf_epoch = 0;
    auto prefetch = [&](uint64_t idx) {
        if (tid == 0 && idx < n_actions) {
            const uint32_t a_ = actions[idx];
            const uint64_t r = a_ >> 1;
            if (r < rows.R) {
                // this row's assignment and, for a REMOVE, its packed group (validated for staleness at use)
                pf_assign = assign[r];
                pf_epoch = ss->epoch;
                pf_pg = ((a_ & 1u) && pf_assign != LB_UNASSIGNED) ? g2p[pf_assign] : LB_UNASSIGNED;
            }
        }
#pragma unroll
        for (int q = 0; q < SW_PF; ++q) {
            if (q * SW_THREADS + warp * 32 >= nf) break;   // warp-uniform: this warp owns no cells of the row
            const int j = tid + q * SW_THREADS;
            pf_raw[q] = 0; pf_obs[q] = 0;
            if (j < nf && idx < n_actions) {
                const uint64_t r = actions[idx] >> 1;
                if (r < rows.R) {
                    const FeatDev &f = s_feat[j];
                    pf_obs[q] = (rows.mask[(uint64_t)s_fid[j] * rows.W + (r >> 5)] >> (r & 31)) & 1u;
                    pf_raw[q] = f.is_wide ? rows.wide[(uint64_t)f.col * rows.R + r]
                                          : (uint32_t)rows.cat8[(uint64_t)f.col * rows.R + r];
                }
            }
        }
    };
    auto commit = [&](int slot) {
        if (tid == 0) { ss->pf_assign[slot] = pf_assign; ss->pf_pg[slot] = pf_pg; ss->pf_epoch[slot] = pf_epoch; }
#pragma unroll
        for (int q = 0; q < SW_PF; ++q) {
            if (q * SW_THREADS + warp * 32 >= nf) break;
            const int j = tid + q * SW_THREADS;
            if (j < nf) { s_raw[slot * nf + j] = pf_raw[q]; s_obs[slot * nf + j] = pf_obs[q]; }
        }
    };
    auto next_uniform = [&](uint64_t idx) {   // drawn by the last warp, off warp 0's critical path
        if (warp == SW_WARPS - 1 && lane == 0 && idx < n_actions && !(actions[idx] & 1u))
            ss->u[idx & 1] = lb_uniform(seed, 0u, offset + idx, (uint32_t)kid);
    };
    prefetch(i); commit((int)(i & 3));
    prefetch(i + 1); commit((int)((i + 1) & 3));
    next_uniform(i);
    sweep_sync<W>();

    // owner-warp update of this warp's feature stripe at group g
    auto update_stripe = [&](const uint32_t *c_raw, const uint32_t *c_obs, int g, int sign) {
        for (int k = off[0] + lane; k < off[N_CLASS]; k += 32) {
            const int j = s_own[k];
            if (!c_obs[j]) continue;
            const FeatDev &f = s_feat[j];
            lb_update_cell(f, aux, ints + (size_t)f.int_row * st + g, flts + (size_t)f.flt_row * st + g,
                           cache + (size_t)f.cache_row * st + g, st, c_raw[j], sign);
        }
        __syncwarp();
    };

    // assignment of the row of action idx: the ring value unless one of the two actions that ran
    // after the prefetch was issued touched the same row (thread 0 only)
    uint64_t prev_r1 = ~0ull, prev_r2 = ~0ull;   // rows of the two previous actions of this launch
    auto row_assignment = [&](uint64_t idx, uint64_t r, bool &fresh) -> uint32_t {
        fresh = r != prev_r1 && r != prev_r2;
        return fresh ? ss->pf_assign[idx & 3] : assign[r];
    };
    // structure changes (rare): add_group / remove_group of product_mixture.cc:178-183,233-238
    auto structure_add_group = [&](int G) {
        // add_group (product_mixture.cc:178-183): append a fresh empty group
        const int ng = G;
        for (int row = tid; row < n_int_rows; row += SW_THREADS) ints[(size_t)row * st + ng] = 0u;
        for (int row = tid; row < n_flt_rows; row += SW_THREADS) flts[(size_t)row * st + ng] = 0.0;
        sweep_sync<W>();
        for (int row = tid; row < n_cache_rows; row += SW_THREADS) {
            const FeatDev &f = s_feat[cache_row_feat[row]];
            lb_refresh_row(f, aux, ints + (size_t)f.int_row * st + ng, flts + (size_t)f.flt_row * st + ng,
                           cache + (size_t)f.cache_row * st + ng, st, row - f.cache_row);
        }
        if (tid == 0) {
            s_counts[ng] = 0u;
            const uint32_t gid = ss->next_global;
            p2g[ng] = gid;
            g2p[gid] = (uint32_t)ng;
            ss->next_global = gid + 1;
            ss->G = ng + 1;
        }
        sweep_sync<W>();
        {   // one more occupied group: refresh the prior of the empty groups
            const float she = logf((alpha + d * (float)(ng + 1 - n_empty)) / (float)n_empty);
            for (int gg = tid; gg <= ng; gg += SW_THREADS) if (!s_counts[gg]) s_prior[gg] = she;
        }
        sweep_sync<W>();
    };
    auto structure_remove_group = [&](int G, int g, uint64_t i) {
        // remove_group (product_mixture.cc:233-238): swap-with-last
        sweep_sync<W>();
        const int last = G - 1;
        if (g != last) {
            for (int row = tid; row < n_int_rows; row += SW_THREADS)
                ints[(size_t)row * st + g] = ints[(size_t)row * st + last];
            for (int row = tid; row < n_flt_rows; row += SW_THREADS)
                flts[(size_t)row * st + g] = flts[(size_t)row * st + last];
            for (int row = tid; row < n_cache_rows; row += SW_THREADS)
                cache[(size_t)row * st + g] = cache[(size_t)row * st + last];
        }
        if (tid == 0) {
            g2p[ss->removed_global[i & 1]] = LB_UNASSIGNED;
            if (g != last) {
                s_counts[g] = s_counts[last];
                s_prior[g] = s_prior[last];
                const uint32_t gid = p2g[last];
                p2g[g] = gid;
                g2p[gid] = (uint32_t)g;
                ss->epoch += 1;
            }
            s_counts[last] = 0u;
            ss->G = last;
        }
        sweep_sync<W>();
        {   // one occupied group fewer: refresh the prior of the empty groups
            const float she = logf((alpha + d * (float)(last - n_empty)) / (float)n_empty);
            for (int gg = tid; gg < last; gg += SW_THREADS) if (!s_counts[gg]) s_prior[gg] = she;
        }
        sweep_sync<W>();
    };
    int status = LB_ST_OK;
    uint64_t cur_r = ~0ull;
    for (; i < n_actions; ++i) {
        const uint32_t act = actions[i];
        const uint64_t r = act >> 1;
        prev_r2 = prev_r1; prev_r1 = cur_r; cur_r = r;
        if (W == 1 && LB_LOCKSTEP >= 1 && (i & (LB_LOCKSTEP_EVERY - 1)) == 0)
            __syncthreads();   // re-align the block's chains (instruction-cache locality)
        const bool is_remove = act & 1u;
        const int slot = (int)(i & 3);
        const uint32_t *c_raw = s_raw + slot * nf, *c_obs = s_obs + slot * nf;
        if (r >= rows.R) { status = LB_ST_BAD_ROW; break; }
        const int G = ss->G;
#ifdef LB_SWEEP_PROFILE
        if (prof && lane == 0) pc = clock64();
#endif
        prefetch(i + 2);   // latency overlaps this action's work

        if (!is_remove) {
            // ---- ADD: model.add_value -> score_value -> sample -> add_value ----
            if (G >= st || ss->next_global >= g2p_cap) { status = LB_ST_NEED_GROW; break; }
            if (tid == 0) { bool fresh; if (row_assignment(i, r, fresh) != LB_UNASSIGNED) ss->stop = LB_ST_DUPLICATE_ROW; }
            if (SW_WARPS <= 2 && off[N_CLASS] > off[CL_GP]) {
                // With one or two warps per chain the class loops run back to back, each behind
                // its own round trip to the tables: start the GP / NICH lines moving before the
                // categorical loop (GP: cache rows 0-1 and the sum row; NICH: cache rows 0-4; two
                // 128-byte lines per row cover 64 groups).
                for (int k = off[CL_GP]; k < off[N_CLASS]; ++k) {
                    const int j = s_own[k];
                    if (!c_obs[j]) continue;
                    const FeatDev &f = s_feat[j];
                    const bool gp = f.type == LB_GP;
                    if (lane < (gp ? 6 : 10)) {
                        const int row = lane >> 1, half = lane & 1;
                        const void *p = (gp && row == 2)
                            ? (const void *)(ints + (size_t)(f.int_row + 1) * st + 32 * half)
                            : (const void *)(cache + (size_t)(f.cache_row + row) * st + 32 * half);
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
                    }
                }
            }
            // 64 groups per trip (two per lane).  ONE tile shape on purpose: every extra
            // instantiation of the scoring loops adds ~10 KB of hot code, and the loop body must
            // stay inside the 32 KB instruction cache when 16 narrow CTAs share an SM.  Columns
            // past G are allocated (st >= G rounded up to 64) and simply ignored.
            for (int g0 = 0; g0 < G; g0 += 64) {
                float acc[2];
                score_warp<2>(off, s_own, s_feat, c_raw, c_obs, ints, cache, st, zero_row, g0, lane, acc);
                s_part[warp * st + g0 + lane] = acc[0];
                s_part[warp * st + g0 + lane + 32] = acc[1];
            }
            LB_TICK(0);
            phase_sync<W>();  // (A) partial scores visible
            if (ss->stop) { status = ss->stop; break; }
            LB_TICK(1);       // after a barrier-dependent read (BAR.SYNC is deferred-blocking)
            if (warp == 0) {
                // clustering.score_value: CRP prior (SURVEY.md 8c Pitman-Yor).  The common
                // -log(n + alpha) term does not affect the draw; it is added for the debug tap only.
                for (int g = lane; g < G; g += 32) {
                    float s = s_prior[g];
#pragma unroll
                    for (int w = 0; w < SW_WARPS; ++w) s += s_part[w * st + g];
                    s_score[g] = s;
                    if (dbg_scores && g < dbg_stride)
                        dbg_scores[((uint64_t)i * K + kid) * dbg_stride + g] = s - logf((float)ss->sample_size + alpha);
                }
                __syncwarp();
                const int pick = warp_sample(s_score, G, ss->u[i & 1], lane);
                if (lane == 0) {
                    // clustering.add_value + push global id (cat_kernel.hpp:222-223)
                    const uint32_t n = s_counts[pick] + 1u;
                    ss->pick[i & 1] = pick;
                    ss->flag[i & 1] = n == 1u;
                    s_counts[pick] = n;
                    s_prior[pick] = logf((float)n - d);
                    ss->sample_size += 1;
                }
            } else {
                commit((int)((i + 2) & 3));
                next_uniform(i + 1);
            }
            LB_TICK(2);
            phase_sync<W>();  // (B) pick visible
            if (warp == 0) {
                commit((int)((i + 2) & 3));
                if (SW_WARPS == 1) next_uniform(i + 1);   // no idle warp to draw it during sampling
            }
            if (tid == 0) {   // after (B): the p2g round trip overlaps the update instead of delaying the barrier
                assign[r] = p2g[ss->pick[i & 1]];
                if (dbg_picks) dbg_picks[(uint64_t)i * K + kid] = (uint32_t)ss->pick[i & 1];
            }
            LB_TICK(3);
        } else {
            // ---- REMOVE: pop global id -> packed; mixture.remove_value ----
            if (tid == 0) {
                bool fresh;
                const uint32_t a = row_assignment(i, r, fresh);
                if (a == LB_UNASSIGNED) ss->stop = LB_ST_REMOVE_UNASSIGNED;
                else {
                    const uint32_t g = (fresh && ss->pf_epoch[i & 3] == ss->epoch) ? ss->pf_pg[i & 3] : g2p[a];
                    const uint32_t n = s_counts[g] - 1u;
                    ss->pick[i & 1] = (int)g;
                    ss->removed_global[i & 1] = a;
                    ss->flag[i & 1] = n == 0u;
                    s_counts[g] = n;
                    if (n) s_prior[g] = logf((float)n - d);
                    ss->sample_size -= 1;
                    assign[r] = LB_UNASSIGNED;
                    if (dbg_picks) dbg_picks[(uint64_t)i * K + kid] = g;
                }
            }
            commit((int)((i + 2) & 3));
            next_uniform(i + 1);
            LB_TICK(5);
            phase_sync<W>();  // (A)
            if (ss->stop) { status = ss->stop; break; }
            LB_TICK(6);
        }
        // ---- both: the owner warps apply the row to group g (one copy of the update code) ----
        const int g = ss->pick[i & 1];
        update_stripe(c_raw, c_obs, g, is_remove ? -1 : +1);
        LB_TICK(is_remove ? 7 : 4);
        if (ss->flag[i & 1]) {
            if (!is_remove) structure_add_group(G);
            else structure_remove_group(G, g, i);
        }
    }
    sweep_sync<W>();
    for (int g = tid; g < st; g += SW_THREADS) {
        const uint32_t n = s_counts[g];
        kd.counts[g] = n;
        kd.shifted[g] = n ? s_prior[g] : 0.f;
    }
    if (tid == 0) {
        kd.G = ss->G;
        kd.sample_size = ss->sample_size;
        kd.next_global = ss->next_global;
        kd.status = status;
        kd.progress = i;
    }
#ifdef LB_SWEEP_PROFILE
    if (prof && lane == 0 && kid < 16)
        for (int k = 0; k < 8; ++k) prof[(kid * SW_WARPS + warp) * 8 + k] += (unsigned long long)pt[k];
#else
    (void)prof;
#endif
#undef LB_TICK
}

#include "lb_sweep_spec.cuh"
#include "lb_sweep_flat.cuh"

static int lb_upload_desc(Engine *e, const void *host, size_t bytes, void **dev);

static size_t sweep_smem_bytes(int nf, int Gcap, int warps) {
    return 256 + (size_t)nf * (sizeof(FeatDev) + 4 + 32 + 4) + (size_t)Gcap * 4 * (3 + warps) + 64;
}

// warps per (chain, kind) CTA: 8 when the launch cannot fill the GPU anyway (lowest latency per
// action), fewer when many chains are swept at once (a CTA spends most of an action waiting on
// dependent loads and barriers, so narrow CTAs, 16 / W resident per SM, finish more actions per
// second in aggregate).  LB_SWEEP_WARPS overrides (tests run every width).
static int sweep_warps(const Engine *e, size_t n_ctas, int max_nf) {
    int w = 8;
    if (n_ctas >= (size_t)e->n_sm * 12) w = 1;
    else if (n_ctas >= (size_t)e->n_sm * 6) w = 2;
    else if (n_ctas >= (size_t)e->n_sm * 3) w = 4;
    if (const char *v = getenv("LB_SWEEP_WARPS")) {
        const int x = atoi(v);
        if (x == 1 || x == 2 || x == 4 || x == 8) w = x;
    }
    while (w < 8 && max_nf > SW_PF * 32 * w) w *= 2;
    return w;
}

// Describe one engine's part of a sweep launch (uploads its active kind ids on ITS stream).
int lb_sweep_chain(Engine *e, const std::vector<int> &kind_ids, const uint32_t *d_actions, uint64_t n_actions,
                   uint64_t seed, uint64_t offset, uint32_t *d_picks, float *d_scores, int stride, SweepChain *out) {
    for (int kid : kind_ids) {
        const KindHost &k = e->kinds[kid];
        if ((int)k.fids.size() > SW_PF * 32 * SW_MAX_WARPS)
            return lb_fail("cat_sweep: kind %d has %zu features (limit %d)", kid, k.fids.size(), SW_PF * 32 * SW_MAX_WARPS);
        // featureless kinds go to the linear-prior kernel (K3f), whose only limit is its count table in shared memory
        const int limit = k.fids.empty() && e->empty_group_count == 1 ? 25600 : SW_MAX_G;
        if (k.Gcap > limit)
            return lb_fail("cat_sweep: kind %d needs %d group slots (limit %d)", kid, k.Gcap, limit);
    }
    if ((int)kind_ids.size() > e->d_kind_ids_cap) {
        if (e->d_kind_ids) cudaFree(e->d_kind_ids);
        e->d_kind_ids = nullptr;
        LB_CUDA(cudaMalloc((void **)&e->d_kind_ids, 4 * (kind_ids.size() + 16)));
        e->d_kind_ids_cap = (int)kind_ids.size() + 16;
    }
    // most expensive kinds first (features x groups): see the grid layout in k_cat_sweep
    std::vector<int> order(kind_ids);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
        const KindHost &ka = e->kinds[a], &kb = e->kinds[b];
        return ka.fids.size() * (size_t)(ka.G + 8) > kb.fids.size() * (size_t)(kb.G + 8);
    });
    LB_CUDA(cudaMemcpyAsync(e->d_kind_ids, order.data(), 4 * order.size(), cudaMemcpyHostToDevice, e->stream));
    SweepChain c;
    c.kinds = e->d_kinds; c.kind_ids = e->d_kind_ids; c.n_ids = (int)kind_ids.size();
    c.feats = e->d_feats; c.kind_feats = e->d_kind_feats; c.aux = e->d_aux;
    c.rows = RowsDev{e->R, e->W, e->d_cat8, e->d_wide, e->d_mask};
    c.actions = d_actions; c.n_actions = n_actions; c.seed = seed; c.offset = offset;
    c.K = (int)e->kinds.size(); c.n_empty = e->empty_group_count;
    c.dbg_picks = d_picks; c.dbg_scores = d_scores; c.dbg_stride = stride; c.prof = e->d_prof;
    *out = c;
    return 0;
}

// K3s launch: the (chain, kind) pairs of a sweep, most expensive first, in three width classes that run side by
// side on three streams: kinds wide enough for 16 warps, for 8 warps, and narrow / featureless kinds that take one
// warp each, SP_CPB to a block.  Returns 1 when a pair does not fit this kernel (the caller falls back).
static int lb_launch_sweep_spec(Engine *leader, const SweepChain *chains, int n, const std::vector<Engine *> &engines,
                                const std::vector<std::vector<int>> &kind_ids) {
    struct Cost { SweepPair p; size_t cost; };
    // 0 narrow (1 warp), 1 wide (8 warps), 2 wider (16 warps), 3 flat: featureless kinds on the linear-prior kernel (K3f)
    std::vector<Cost> cls[4];
    size_t smem[4] = {0, 0, 0, 0};
    bool catonly = true;                           // no GP / NICH column anywhere in the launch: the lean instantiation
    const int warps_of[3] = {1, 8, LB_SPEC_WIDER};
    const char *w16 = getenv("LB_SPEC_WIDE16");
    const bool use16 = !w16 || atoi(w16) != 0;
    const char *fl = getenv("LB_SWEEP_FLAT");
    const bool use_flat = !fl || atoi(fl) != 0;
    const char *wn = getenv("LB_SPEC_WIDE_NF");         // kinds with at least this many columns get an 8-warp CTA
    const int wide_nf = wn ? std::max(1, atoi(wn)) : SP_WIDE_NF;
    for (int c = 0; c < n; ++c)
        for (int kid : kind_ids[c]) {
            const KindHost &k = engines[c]->kinds[kid];
            const int nf = (int)k.fids.size();
            if (nf == 0 && use_flat && engines[c]->empty_group_count == 1 && !chains[c].dbg_scores && 8 * (size_t)k.Gcap <= 200 * 1024) {
                smem[3] = std::max(smem[3], 8 * (size_t)k.Gcap);
                cls[3].push_back(Cost{SweepPair{c, kid}, (size_t)(k.G + 8)});
                continue;
            }
            for (int f : k.fids) {
                const int t = engines[c]->specs[f].type;
                catonly = catonly && (t == LB_BB || t == LB_DD || t == LB_DPD);
            }
            const int w = nf >= SP_WIDER_NF && use16 ? 2 : (nf >= wide_nf ? 1 : 0);
            if (nf > SW_PF * 32 * warps_of[w]) return 1;
            if (k.Gcap > SW_MAX_G) return 1;
            smem[w] = std::max(smem[w], (spec_smem_bytes(nf, k.Gcap, warps_of[w]) + 15) / 16 * 16);
            cls[w].push_back(Cost{SweepPair{c, kid}, (size_t)(nf + 1) * (size_t)(k.G + 8)});
        }
    if (smem[1] > 220 * 1024 || smem[2] > 220 * 1024 || smem[0] * SP_CPB > 220 * 1024) return 1;
    auto by_cost = [](const Cost &a, const Cost &b) { return a.cost > b.cost; };
    for (auto &v : cls) std::stable_sort(v.begin(), v.end(), by_cost);
    // one upload: [chains][wider pairs][wide pairs][narrow pairs][flat pairs]
    const size_t cb = (sizeof(SweepChain) * (size_t)n + 15) / 16 * 16;
    const size_t n_pairs = cls[0].size() + cls[1].size() + cls[2].size() + cls[3].size();
    std::vector<unsigned char> host(cb + sizeof(SweepPair) * n_pairs + 16);
    memcpy(host.data(), chains, sizeof(SweepChain) * (size_t)n);
    SweepPair *hp = (SweepPair *)(host.data() + cb);
    size_t at = 0, first[4];
    for (int w : {2, 1, 0, 3}) { first[w] = at; for (const Cost &c : cls[w]) hp[at++] = c.p; }
    void *d = nullptr;
    LB_TRY(lb_upload_desc(leader, host.data(), host.size(), &d));
    const SweepChain *d_chains = (const SweepChain *)d;
    const SweepPair *d_pairs = (const SweepPair *)((unsigned char *)d + cb);
    if (!leader->stream2) {
        LB_CUDA(cudaStreamCreateWithFlags(&leader->stream2, cudaStreamNonBlocking));
        LB_CUDA(cudaStreamCreateWithFlags(&leader->stream3, cudaStreamNonBlocking));
        LB_CUDA(cudaStreamCreateWithFlags(&leader->stream4, cudaStreamNonBlocking));
        LB_CUDA(cudaEventCreateWithFlags(&leader->ev_fork, cudaEventDisableTiming));
        LB_CUDA(cudaEventCreateWithFlags(&leader->ev_join, cudaEventDisableTiming));
        LB_CUDA(cudaEventCreateWithFlags(&leader->ev_join3, cudaEventDisableTiming));
        LB_CUDA(cudaEventCreateWithFlags(&leader->ev_join4, cudaEventDisableTiming));
    }
    if (const char *v = getenv("LB_SPEC_CATONLY")) catonly = catonly && atoi(v) != 0;      // A/B switch
    auto kernel_of = [&](int w) {
        if (w == 2) return catonly ? k_cat_sweep_spec<LB_SPEC_WIDER, true> : k_cat_sweep_spec<LB_SPEC_WIDER, false>;
        if (w == 1) return catonly ? k_cat_sweep_spec<8, true> : k_cat_sweep_spec<8, false>;
        return catonly ? k_cat_sweep_spec<1, true> : k_cat_sweep_spec<1, false>;
    };
    // the opt-in is per device and cheap: set it at every launch (engines live on any device)
    for (int w = 0; w < 3; ++w)
        if (!cls[w].empty())
            LB_CUDA(cudaFuncSetAttribute(kernel_of(w), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(w ? smem[w] : smem[0] * SP_CPB)));
    // flat class: as many one-warp pairs to a block as their count tables leave room for
    const int flat_ppb = cls[3].empty() ? 1 : (int)std::max<size_t>(1, std::min<size_t>(8, (200 * 1024) / std::max<size_t>(smem[3], 1)));
    if (!cls[3].empty())
        LB_CUDA(cudaFuncSetAttribute(k_cat_sweep_flat, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem[3] * flat_ppb)));
    LB_CUDA(cudaEventRecord(leader->kev0, leader->stream));
    LB_CUDA(cudaEventRecord(leader->ev_fork, leader->stream));
    // the longest-running class goes first on the leader's stream; the others fork off it and join back
    const int lead_cls = !cls[2].empty() ? 2 : (!cls[1].empty() ? 1 : (!cls[0].empty() ? 0 : 3));
    cudaStream_t side[3] = {leader->stream2, leader->stream3, leader->stream4};
    cudaEvent_t joined[3] = {leader->ev_join, leader->ev_join3, leader->ev_join4};
    int n_side = 0;
    for (int w : {2, 1, 0, 3}) {
        if (cls[w].empty()) continue;
        cudaStream_t s = leader->stream;
        if (w != lead_cls) { s = side[n_side]; LB_CUDA(cudaStreamWaitEvent(s, leader->ev_fork, 0)); }
        const int np = (int)cls[w].size();
        const bool timing = getenv("LB_SWEEP_TIMING") != nullptr;
        leader->cls_pairs[w] = timing ? np : 0;
        if (timing) {
            if (!leader->cls_ev[w][0]) { LB_CUDA(cudaEventCreate(&leader->cls_ev[w][0])); LB_CUDA(cudaEventCreate(&leader->cls_ev[w][1])); }
            LB_CUDA(cudaEventRecord(leader->cls_ev[w][0], s));
        }
        if (w == 2) kernel_of(2)<<<(unsigned)np, 32 * LB_SPEC_WIDER, smem[2], s>>>(d_chains, d_pairs + first[2], np, (int)smem[2]);
        else if (w == 1) kernel_of(1)<<<(unsigned)np, 256, smem[1], s>>>(d_chains, d_pairs + first[1], np, (int)smem[1]);
        else if (w == 0) kernel_of(0)<<<(unsigned)((np + SP_CPB - 1) / SP_CPB), 32 * SP_CPB, smem[0] * SP_CPB, s>>>(
                 d_chains, d_pairs + first[0], np, (int)smem[0]);
        else k_cat_sweep_flat<<<(unsigned)((np + flat_ppb - 1) / flat_ppb), 32 * flat_ppb, smem[3] * flat_ppb, s>>>(
                 d_chains, d_pairs + first[3], np, (int)smem[3]);
        LB_CUDA(cudaGetLastError());
        if (timing) LB_CUDA(cudaEventRecord(leader->cls_ev[w][1], s));
        if (w != lead_cls) { LB_CUDA(cudaEventRecord(joined[n_side], s)); ++n_side; }
        leader->launches[LB_L_TOTAL]++; leader->launches[LB_L_SWEEP]++;
    }
    for (int w : {0, 1, 2, 3}) if (cls[w].empty()) leader->cls_pairs[w] = 0;
    for (int q = 0; q < n_side; ++q) LB_CUDA(cudaStreamWaitEvent(leader->stream, joined[q], 0));
    LB_CUDA(cudaEventRecord(leader->kev1, leader->stream));
    leader->kev_pending = true;
    return 0;
}

// which sweep kernel: the speculative one (K3s) unless the launch holds so many (chain, kind) pairs that the
// throughput form (k_cat_sweep with narrow CTAs) wins; LB_SWEEP_KERNEL = spec | classic overrides, and
// LB_SWEEP_WARPS (a classic-kernel width) implies classic
static bool sweep_use_spec(const Engine *e, size_t n_ctas) {
    if (const char *v = getenv("LB_SWEEP_KERNEL")) return strcmp(v, "classic") != 0;
    if (getenv("LB_SWEEP_WARPS")) return false;
    return n_ctas < (size_t)e->n_sm * 6;
}

// One launch for `n` chains on the leader's stream: grid (max active kinds, chains).
int lb_launch_sweep_chains(Engine *leader, const SweepChain *chains, int n, const std::vector<Engine *> &engines,
                           const std::vector<std::vector<int>> &kind_ids) {
    size_t n_ctas = 0, n_featured = 0, smem = 0;
    int max_nf = 0, max_ids = 0;
    for (int c = 0; c < n; ++c) {
        n_ctas += kind_ids[c].size();
        max_ids = std::max(max_ids, (int)kind_ids[c].size());
        for (int kid : kind_ids[c]) {
            max_nf = std::max(max_nf, (int)engines[c]->kinds[kid].fids.size());
            n_featured += !engines[c]->kinds[kid].fids.empty();
        }
    }
    if (!n_ctas) return 0;
    if (sweep_use_spec(leader, std::max<size_t>(n_featured, 1))) {   // featureless kinds are one cheap warp each (K3f)
        const int rc = lb_launch_sweep_spec(leader, chains, n, engines, kind_ids);
        if (rc <= 0) return rc;      // launched (0) or failed (< 0); 1 = does not fit, use the classic kernel
    }
    if (max_ids > 65535) return lb_fail("cat_sweep: %d kinds in one launch (limit 65535)", max_ids);
    for (int c = 0; c < n; ++c)
        for (int kid : kind_ids[c])
            if (engines[c]->kinds[kid].Gcap > SW_MAX_G)
                return lb_fail("cat_sweep: kind %d needs %d group slots (limit %d in the many-chain kernel)", kid, engines[c]->kinds[kid].Gcap, SW_MAX_G);
    int W = sweep_warps(leader, n_ctas, max_nf);
    for (;;) {
        smem = 0;
        for (int c = 0; c < n; ++c)
            for (int kid : kind_ids[c]) {
                const KindHost &k = engines[c]->kinds[kid];
                smem = std::max(smem, sweep_smem_bytes((int)k.fids.size(), k.Gcap, W));
            }
        smem = (smem + 15) / 16 * 16;
        if (W == 1 && smem * SW_CPB > 200 * 1024) { W = 2; continue; }   // SW_CPB chains share a block's shared memory
        break;
    }
    {   // the opt-in is per device and cheap: set it at every launch (engines live on any device)
        const void *fn = W == 1 ? (const void *)k_cat_sweep<1> : W == 2 ? (const void *)k_cat_sweep<2>
                       : W == 4 ? (const void *)k_cat_sweep<4> : (const void *)k_cat_sweep<8>;
        LB_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(W == 1 ? smem * SW_CPB : smem)));
    }
    const SweepChain *d_many = nullptr;
    if (n > 1) {
        void *d = nullptr;
        LB_TRY(lb_upload_desc(leader, chains, sizeof(SweepChain) * (size_t)n, &d));
        d_many = (const SweepChain *)d;
    }
    const dim3 grid((unsigned)(W == 1 ? (n + SW_CPB - 1) / SW_CPB : n), (unsigned)max_ids);
    LB_CUDA(cudaEventRecord(leader->kev0, leader->stream));
    switch (W) {
    case 1: k_cat_sweep<1><<<grid, 32 * SW_CPB, smem * SW_CPB, leader->stream>>>(chains[0], d_many, n, (int)smem); break;
    case 2: k_cat_sweep<2><<<grid, 64, smem, leader->stream>>>(chains[0], d_many, n, (int)smem); break;
    case 4: k_cat_sweep<4><<<grid, 128, smem, leader->stream>>>(chains[0], d_many, n, (int)smem); break;
    default: k_cat_sweep<8><<<grid, 256, smem, leader->stream>>>(chains[0], d_many, n, (int)smem); break;
    }
    LB_CUDA(cudaGetLastError());
    LB_CUDA(cudaEventRecord(leader->kev1, leader->stream));
    leader->kev_pending = true;
    leader->launches[LB_L_TOTAL]++; leader->launches[LB_L_SWEEP]++;
    return 0;
}

int lb_launch_sweep(Engine *e, const std::vector<int> &kind_ids, uint64_t n_actions, uint64_t seed,
                    uint64_t offset, uint32_t *d_picks, float *d_scores, int stride) {
    SweepChain c;
    LB_TRY(lb_sweep_chain(e, kind_ids, e->d_actions, n_actions, seed, offset, d_picks, d_scores, stride, &c));
    return lb_launch_sweep_chains(e, &c, 1, {e}, {kind_ids});
}

// ===========================================================================
// K1: batch score_value with frozen tables: out[(r - b) * G + g].
//
// A CTA takes tiles of SC_TR consecutive rows.  The tile's cells are staged into
// shared memory with coalesced column reads (rows are contiguous inside a feature
// column), then each warp scores SC_TR/8 rows with lanes = groups:
//   * categorical cells and small GP counts are branch-free gathers of one table
//     row (value row minus shift row; unobserved cells read the all-zero row); the
//     GP predictive table for counts < LB_GP_TAB is built by k_gp_table just
//     before the launch, so no lgamma is evaluated per (cell, group);
//   * NICH loads its five per-group parameters once per feature and reuses them
//     for all rows of the warp.
// The 8 x NJ independent gathers per feature give the memory system the
// parallelism the one-row-at-a-time version lacked.
// ===========================================================================
constexpr int SC_THREADS = 256;
constexpr int SC_TR = 64;               // rows per tile
constexpr int SC_RW = SC_TR / 8;        // rows per warp
constexpr int SC_FC = 128;              // features staged per chunk
constexpr int LB_GP_TAB = 16;

// Score table of one kind, built right before a batch-score launch: row 0 is all zero
// ("unobserved"), then for every non-NICH feature one row per value holding the COMPLETE
// log predictive of that value for every group:
//   BB  2 rows (value 0, value 1)          DD/DPD  dim rows, already minus log(A + N)
//   GP  LB_GP_TAB rows (counts 0..15; larger counts are evaluated directly)
// so scoring a cell is one gather + one add per 32 groups.
__global__ void k_score_table(const KindDev *kinds, int kid, const FeatDev *feats, const int32_t *kind_feats,
                              const int32_t *row_base, const int32_t *row_feat, int n_rows, float *tab) {
    const KindDev &kd = kinds[kid];
    const int G = kd.G, st = kd.Gcap;
    const uint64_t total = (uint64_t)n_rows * G;
    for (uint64_t idx = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; idx < total;
         idx += (uint64_t)gridDim.x * blockDim.x) {
        const int row = (int)(idx / G), g = (int)(idx % G);
        float v = 0.f;
        if (row > 0) {
            const int j = row_feat[row];
            const int local = row - row_base[j];
            const FeatDev &f = feats[kind_feats[kd.feat_begin + j]];
            const float *c = kd.cache + (size_t)f.cache_row * st + g;
            if (f.type == LB_BB) v = c[local ? 0 : st];           // cache rows: 0 = log p(1), 1 = log p(0)
            else if (f.type == LB_GP) {
                const float a = f.hp[0] + (float)kd.ints[(size_t)(f.int_row + 1) * st + g];
                v = lb_lgamma_diff_int(a, (uint32_t)local) - lb_log_factorial((uint32_t)local) + c[0] + (float)local * c[st];
            } else v = c[(size_t)local * st] - c[(size_t)f.dim * st];
        }
        tab[(size_t)row * st + g] = v;
    }
}

// LSE = true (one tile of <= 128 groups covers the kind): the scores are not written out; each warp reduces its rows'
// score vectors to log_sum_exp_g in registers and out[r - b] receives (lse_first) or accumulates that scalar -- the
// "scores not materialised" form of K1, what QueryServer Score consumes.
template <int NJ, bool LSE = false>
__global__ void __launch_bounds__(SC_THREADS)
k_score_rows(const KindDev *kinds, int kid, const FeatDev *feats, const int32_t *kind_feats,
             const int32_t *row_base, const float *tab, RowsDev rows, uint64_t b, uint64_t en,
             int n_empty, int g_base, float *out, int lse_first = 0, const int32_t *flist = nullptr, int n_flist = 0) {
    __shared__ FeatDev s_feat[SC_FC];
    __shared__ int32_t s_fid[SC_FC];
    __shared__ uint32_t s_row[SC_FC][SC_TR];     // table row of the cell (0 = unobserved), or raw bits for NICH
    __shared__ uint32_t s_obs[SC_FC][SC_TR / 32];
    __shared__ int32_t s_rb[SC_FC];
    const KindDev &kd = kinds[kid];
    const int nf = flist ? n_flist : kd.nf, st = kd.Gcap, G = kd.G;     // flist: the kind-local features to score (K1d: the untared ones)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *cache = kd.cache;
    const uint64_t n_tiles = (en - b + SC_TR - 1) / SC_TR;
    const float base = -logf((float)kd.sample_size + kd.alpha);
    const float sh_empty = logf((kd.alpha + kd.d * (float)(G - n_empty)) / (float)n_empty);
    float prior[NJ];
#pragma unroll
    for (int jj = 0; jj < NJ; ++jj) {
        const int g = g_base + lane + 32 * jj;
        prior[jj] = g < G ? (kd.counts[g] ? kd.shifted[g] : sh_empty) + base : 0.f;
    }
    const float *tabg = tab + g_base + lane;
    bool live[NJ];                      // this lane's groups that exist: padding columns are never loaded (the gather is
#pragma unroll                          // L2-bandwidth-bound on DD-256 kinds: 68 groups in 128 columns was 47 % wasted sectors)
    for (int jj = 0; jj < NJ; ++jj) live[jj] = g_base + lane + 32 * jj < G;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint64_t r0 = b + tile * SC_TR;
        float acc[SC_RW][NJ];
#pragma unroll
        for (int q = 0; q < SC_RW; ++q)
#pragma unroll
            for (int jj = 0; jj < NJ; ++jj) acc[q][jj] = prior[jj];
        for (int f0 = 0; f0 < nf; f0 += SC_FC) {
            const int fc = min(SC_FC, nf - f0);
            __syncthreads();   // previous chunk fully consumed
            for (int j = tid; j < fc; j += SC_THREADS) {
                const int jl = flist ? flist[f0 + j] : f0 + j;           // kind-local index
                const int fid = kind_feats[kd.feat_begin + jl];
                s_feat[j] = feats[fid];
                s_fid[j] = fid;
                s_rb[j] = row_base[jl];
            }
            __syncthreads();
            for (int idx = tid; idx < fc * (SC_TR / 32); idx += SC_THREADS) {
                const int j = idx / (SC_TR / 32), w = idx % (SC_TR / 32);
                const uint64_t r = r0 + 32 * w;
                uint32_t bits = 0;
                if ((r & 31) == 0 && r + 32 <= en) {
                    bits = rows.mask[(uint64_t)s_fid[j] * rows.W + (r >> 5)];
                } else {
                    for (int k = 0; k < 32; ++k) {
                        const uint64_t rk = r + k;
                        if (rk < en) bits |= ((rows.mask[(uint64_t)s_fid[j] * rows.W + (rk >> 5)] >> (rk & 31)) & 1u) << k;
                    }
                }
                s_obs[j][w] = bits;
            }
            __syncthreads();
            // stage the tile: consecutive threads read consecutive rows of one column and turn
            // the cell into the index of its score-table row
            for (int idx = tid; idx < fc * SC_TR; idx += SC_THREADS) {
                const int j = idx / SC_TR, rr = idx % SC_TR;
                const uint64_t r = r0 + rr;
                const FeatDev &f = s_feat[j];
                uint32_t raw = 0;
                if (r < en) raw = f.is_wide ? rows.wide[(uint64_t)f.col * rows.R + r]
                                            : (uint32_t)rows.cat8[(uint64_t)f.col * rows.R + r];
                const bool obs = (s_obs[j][rr >> 5] >> (rr & 31)) & 1u;
                uint32_t v = raw;
                if (f.type != LB_NICH) {
                    if (!obs) v = 0u;
                    else if (f.type == LB_GP && raw >= (uint32_t)LB_GP_TAB) v = 0x80000000u | raw;  // direct evaluation
                    else v = (uint32_t)s_rb[j] + raw;
                }
                s_row[j][rr] = v;
            }
            __syncthreads();
            for (int j = 0; j < fc; ++j) {
                const FeatDev &f = s_feat[j];
                if (f.type == LB_NICH) {
                    float c0[NJ], c1[NJ], c2[NJ], c3[NJ], c4[NJ];
#pragma unroll
                    for (int jj = 0; jj < NJ; ++jj) {
                        const float *c = cache + (size_t)f.cache_row * st + g_base + lane + 32 * jj;
                        c0[jj] = c[0]; c1[jj] = c[st]; c2[jj] = c[2 * (size_t)st];
                        c3[jj] = c[3 * (size_t)st]; c4[jj] = c[4 * (size_t)st];
                    }
#pragma unroll
                    for (int q = 0; q < SC_RW; ++q) {
                        const int rr = warp * SC_RW + q;
                        if (!((s_obs[j][rr >> 5] >> (rr & 31)) & 1u)) continue;
                        const float x = __uint_as_float(s_row[j][rr]);
#pragma unroll
                        for (int jj = 0; jj < NJ; ++jj) {
                            const float dx = (x - c3[jj]) - c4[jj];
                            acc[q][jj] += c0[jj] - c1[jj] * log1pf(c2[jj] * dx * dx);
                        }
                    }
                } else if (f.type == LB_GP) {
#pragma unroll
                    for (int q = 0; q < SC_RW; ++q) {
                        const uint32_t row = s_row[j][warp * SC_RW + q];
                        if (row & 0x80000000u) {   // rare: large count
                            const uint32_t raw = row & 0x7fffffffu;
                            const float lf = lb_log_factorial(raw);
#pragma unroll
                            for (int jj = 0; jj < NJ; ++jj) {
                                const int g = g_base + lane + 32 * jj;
                                acc[q][jj] += lb_score_cell(f, kd.ints + (size_t)f.int_row * st + g,
                                                            cache + (size_t)f.cache_row * st + g, st, raw, lf);
                            }
                        } else {
#pragma unroll
                            for (int jj = 0; jj < NJ; ++jj) if (live[jj]) acc[q][jj] += tabg[(size_t)row * st + 32 * jj];
                        }
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < SC_RW; ++q) {
                        const uint32_t row = s_row[j][warp * SC_RW + q];
#pragma unroll
                        for (int jj = 0; jj < NJ; ++jj) if (live[jj]) acc[q][jj] += tabg[(size_t)row * st + 32 * jj];
                    }
                }
            }
        }
#pragma unroll
        for (int q = 0; q < SC_RW; ++q) {
            const uint64_t r = r0 + warp * SC_RW + q;
            if (r >= en) continue;          // warp-uniform
            if constexpr (LSE) {
                float m = -INFINITY;
#pragma unroll
                for (int jj = 0; jj < NJ; ++jj)
                    if (g_base + lane + 32 * jj < G) m = fmaxf(m, acc[q][jj]);
                for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL, m, o));
                float t = 0.f;
#pragma unroll
                for (int jj = 0; jj < NJ; ++jj)
                    if (g_base + lane + 32 * jj < G) t += expf(acc[q][jj] - m);
                for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(FULL, t, o);
                if (lane == 0) out[r - b] = (lse_first ? 0.f : out[r - b]) + m + logf(t);
            } else {
#pragma unroll
                for (int jj = 0; jj < NJ; ++jj) {
                    const int g = g_base + lane + 32 * jj;
                    if (g < G) out[(r - b) * (uint64_t)G + g] = acc[q][jj];
                }
            }
        }
    }
}

// acc[r] (+)= log_sum_exp over the G scores of row r; one warp per row
__global__ void k_row_lse_add(const float *scores, uint64_t n_rows, int G, float *acc, int first) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t r = warp; r < n_rows; r += n_warps) {
        const float *s = scores + r * (uint64_t)G;
        float m = -INFINITY;
        for (int g = lane; g < G; g += 32) m = fmaxf(m, s[g]);
        for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL, m, o));
        float t = 0.f;
        for (int g = lane; g < G; g += 32) t += expf(s[g] - m);
        for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(FULL, t, o);
        if (lane == 0) acc[r] = (first ? 0.f : acc[r]) + m + logf(t);
    }
}
int lb_launch_row_lse_add(Engine *e, const float *d_scores, uint64_t n_rows, int G, float *d_acc, int first) {
    const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((n_rows * 32 + 255) / 256, (uint64_t)e->n_sm * 8));
    k_row_lse_add<<<grid, 256, 0, e->stream>>>(d_scores, n_rows, G, d_acc, first);
    LB_CUDA(cudaGetLastError());
    e->launches[LB_L_TOTAL]++; e->launches[LB_L_SCORE]++;
    return 0;
}

// Score table of a kind (k_score_table) in e->d_score_aux: [row_base nf][row_feat n_rows][pad][tab n_rows x Gcap]
static int score_table_build(Engine *e, int kid, std::vector<int32_t> &row_base, int *n_rows_out, int32_t **d_row_base_out,
                             float **d_tab_out) {
    const KindHost &k = e->kinds[kid];
    const int nf = (int)k.fids.size();
    row_base.assign(std::max(nf, 1), 0);
    std::vector<int32_t> row_feat(1, -1);
    for (int j = 0; j < nf; ++j) {
        const lb_feature_spec &s = e->specs[k.fids[j]];
        row_base[j] = (int32_t)row_feat.size();
        const int n = s.type == LB_BB ? 2 : (s.type == LB_GP ? LB_GP_TAB : (s.type == LB_NICH ? 0 : s.dim));
        row_feat.insert(row_feat.end(), n, (int32_t)j);
    }
    const int n_rows = (int)row_feat.size();
    const size_t idx_bytes = ((4 * (row_base.size() + row_feat.size()) + 255) / 256) * 256;
    const size_t need = idx_bytes + 4 * (size_t)n_rows * k.Gcap + 256;
    if (need > e->score_aux_bytes) {
        if (e->d_score_aux) cudaFree(e->d_score_aux);
        e->d_score_aux = nullptr;
        LB_CUDA(cudaMalloc(&e->d_score_aux, need));
        e->score_aux_bytes = need;
    }
    int32_t *d_row_base = (int32_t *)e->d_score_aux;
    int32_t *d_row_feat = d_row_base + row_base.size();
    float *d_tab = (float *)((char *)e->d_score_aux + idx_bytes);
    LB_CUDA(cudaMemcpyAsync(d_row_base, row_base.data(), 4 * row_base.size(), cudaMemcpyHostToDevice, e->stream));
    LB_CUDA(cudaMemcpyAsync(d_row_feat, row_feat.data(), 4 * row_feat.size(), cudaMemcpyHostToDevice, e->stream));
    const uint64_t tot = (uint64_t)n_rows * k.G;
    k_score_table<<<(int)std::max<uint64_t>(1, std::min<uint64_t>((tot + 255) / 256, (uint64_t)e->n_sm * 8)), 256, 0, e->stream>>>(
        e->d_kinds, kid, e->d_feats, e->d_kind_feats, d_row_base, d_row_feat, n_rows, d_tab);
    LB_CUDA(cudaGetLastError());
    e->launches[LB_L_TOTAL]++; e->launches[LB_L_SCORE]++;
    *n_rows_out = n_rows; *d_row_base_out = d_row_base; *d_tab_out = d_tab;
    return 0;
}

// lse: 0 = scores out [rows][G]; 1 / 2 = out[row] set to / increased by the row's log_sum_exp (needs G <= 128)
int lb_launch_score_rows(Engine *e, int kid, uint64_t b, uint64_t en, float *d_out, int lse) {
    const KindHost &k = e->kinds[kid];
    const int nf = (int)k.fids.size();
    if (lse && k.G > 128) return lb_fail("fused score + log-sum-exp needs at most 128 groups (kind %d has %d)", kid, k.G);
    std::vector<int32_t> row_base;
    int n_rows = 0;
    int32_t *d_row_base = nullptr;
    float *d_tab = nullptr;
    LB_TRY(score_table_build(e, kid, row_base, &n_rows, &d_row_base, &d_tab));
    if (!lse) {
        // K1t: a kind made of categorical / Bernoulli columns only can take the tensor-core contraction
        // (lb_score_tc.cu).  LB_K1_TC = 1 always when applicable, 0 never, unset: when a feature has at most 32 values
        // on average (the contraction pays for every VALUE of a column, the gather for every observed CELL).
        const char *env = getenv("LB_K1_TC");
        const int K_total = n_rows - 1;
        const bool want = env ? atoi(env) != 0 : (nf > 0 && K_total <= 32 * nf);
        if (want) {
            const char *pl = getenv("LB_K1_TC_PLANES");
            int used = 0;
            LB_TRY(lb_launch_score_rows_tc(e, kid, b, en, d_tab, row_base.data(), K_total, d_out,
                                           pl ? std::max(1, std::min(3, atoi(pl))) : 3, &used));
            if (used) return 0;
        }
    }
    const uint64_t n_tiles = (en - b + SC_TR - 1) / SC_TR;
    const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>(n_tiles, (uint64_t)e->n_sm * 4));
    RowsDev rows{e->R, e->W, e->d_cat8, e->d_wide, e->d_mask};
    if (lse) {
        const int first = lse == 1;
        if (k.G <= 32)
            k_score_rows<1, true><<<grid, SC_THREADS, 0, e->stream>>>(e->d_kinds, kid, e->d_feats, e->d_kind_feats, d_row_base,
                                                                     d_tab, rows, b, en, e->empty_group_count, 0, d_out, first);
        else if (k.G <= 64)
            k_score_rows<2, true><<<grid, SC_THREADS, 0, e->stream>>>(e->d_kinds, kid, e->d_feats, e->d_kind_feats, d_row_base,
                                                                     d_tab, rows, b, en, e->empty_group_count, 0, d_out, first);
        else
            k_score_rows<4, true><<<grid, SC_THREADS, 0, e->stream>>>(e->d_kinds, kid, e->d_feats, e->d_kind_feats, d_row_base,
                                                                     d_tab, rows, b, en, e->empty_group_count, 0, d_out, first);
        LB_CUDA(cudaGetLastError());
        e->launches[LB_L_TOTAL]++; e->launches[LB_L_SCORE]++;
        return 0;
    }
    for (int g_base = 0; g_base < k.G; g_base += 128) {
        const int left = k.G - g_base;
        if (left <= 32)
            k_score_rows<1><<<grid, SC_THREADS, 0, e->stream>>>(e->d_kinds, kid, e->d_feats, e->d_kind_feats, d_row_base,
                                                               d_tab, rows, b, en, e->empty_group_count, g_base, d_out);
        else if (left <= 64)
            k_score_rows<2><<<grid, SC_THREADS, 0, e->stream>>>(e->d_kinds, kid, e->d_feats, e->d_kind_feats, d_row_base,
                                                               d_tab, rows, b, en, e->empty_group_count, g_base, d_out);
        else
            k_score_rows<4><<<grid, SC_THREADS, 0, e->stream>>>(e->d_kinds, kid, e->d_feats, e->d_kind_feats, d_row_base,
                                                               d_tab, rows, b, en, e->empty_group_count, g_base, d_out);
        LB_CUDA(cudaGetLastError());
        e->launches[LB_L_TOTAL]++; e->launches[LB_L_SCORE]++;
    }
    return 0;
}

// ===========================================================================
// K1d: ProductMixture::score_diff (src/product_mixture.cc:436-463) from the tare-compressed block:
//   scores[g] = prior[g] + (row references the tare ? tare_score[g] : 0) + sum_{pos cells} p(cell | g) - sum_{neg cells} p(tare cell | g)
// tare_score[g] is the kind's TareCache entry (product_mixture.hpp:53-57; _update_tare_cache, cc:58-81): the score of
// the tare row's cells of this kind in group g -- one vector per kind, rebuilt per call from the same score table the
// dense scorer gathers from (frozen tables here; the sweep keeps its own).  Work per row is O(nnz x G), not O(F x G).
// One warp per row, lanes = groups; entries of other kinds are skipped through the feature -> table-row map.
// ===========================================================================
__global__ void k_tare_scores(const KindDev *kinds, int kid, const int32_t *kind_feats, const int32_t *row_base,
                              const float *tab, const uint8_t *tare_obs, const uint32_t *tare_val, float *tare_scores,
                              int32_t *frow, int32_t *fn, const FeatDev *feats) {
    const KindDev &kd = kinds[kid];
    const int st = kd.Gcap;
    // feature -> first table row, for the TARED features of this kind only (the rest stay -1: the dense pass scores them)
    for (int j = threadIdx.x; j < kd.nf; j += blockDim.x) {
        const int f = kind_feats[kd.feat_begin + j];
        if (tare_obs[f]) {
            frow[f] = row_base[j];
            fn[f] = feats[f].type == LB_BB ? 2 : (feats[f].type == LB_GP ? LB_GP_TAB : feats[f].dim);    // tabulated values
        }
    }
    for (int g = threadIdx.x; g < st; g += blockDim.x) {
        float acc = 0.f;
        if (g < kd.G)
            for (int j = 0; j < kd.nf; ++j) {
                const int f = kind_feats[kd.feat_begin + j];
                if (tare_obs[f]) acc += tab[(size_t)(row_base[j] + (int)tare_val[f]) * st + g];
            }
        tare_scores[g] = acc;
    }
}

// out[r][g] += (row references the tare ? tare_score[g] : 0) + sum over the row's pos cells in tared columns of this
// kind - sum over its neg cells.  A warp takes a row: 32 diff entries are read at once (coalesced), each lane turns its
// entry into a table-row offset (or nothing: another kind / an untared column), and the warp walks the hits.
__global__ void __launch_bounds__(256)
k_score_rows_diff(const KindDev *kinds, int kid, const FeatDev *feats, const int32_t *frow, const int32_t *fn, const float *tab,
                  const float *tare_scores,
                  const uint8_t *row_has_tare, const uint32_t *tare_val, const uint64_t *pos_ptr,
                  const uint32_t *pos_feature, const uint32_t *pos_value, const uint64_t *neg_ptr, const uint32_t *neg_feature,
                  uint64_t b, uint64_t en, float *out) {
    const KindDev &kd = kinds[kid];
    const int G = kd.G, st = kd.Gcap;
    const int lane = threadIdx.x & 31;
    const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5, n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (int g0 = 0; g0 < G; g0 += 128) {
        float tare[4];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) tare[jj] = g0 + lane + 32 * jj < G ? tare_scores[g0 + lane + 32 * jj] : 0.f;
        const float *tabg = tab + g0 + lane;
        for (uint64_t r = b + warp; r < en; r += n_warps) {
            float acc[4];
            const bool has = !row_has_tare || row_has_tare[r];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) acc[jj] = has ? tare[jj] : 0.f;
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {           // 0: pos cells (+), 1: neg cells (-)
                const uint64_t lo = pass ? neg_ptr[r] : pos_ptr[r], hi = pass ? neg_ptr[r + 1] : pos_ptr[r + 1];
                for (uint64_t i0 = lo; i0 < hi; i0 += 32) {
                    const uint64_t i = i0 + lane;
                    int off = -1, f = -1;
                    uint32_t raw = 0;
                    if (i < hi) {
                        f = (int)(pass ? neg_feature[i] : pos_feature[i]);
                        const int row0 = frow[f];
                        if (row0 >= 0) {
                            raw = pass ? tare_val[f] : pos_value[i];
                            off = raw < (uint32_t)fn[f] ? (row0 + (int)raw) * st : -2;      // -2: a count beyond the table (GP >= 16)
                        }
                    }
                    unsigned direct = __ballot_sync(FULL, off == -2);
                    while (direct) {                              // rare: evaluate the predictive directly
                        const int src = __ffs(direct) - 1;
                        direct &= direct - 1;
                        const FeatDev &fd = feats[__shfl_sync(FULL, f, src)];
                        const uint32_t x = __shfl_sync(FULL, raw, src);
                        const float lf = lb_log_factorial(x);
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            const int g = g0 + lane + 32 * jj;
                            if (g < G) {
                                const float v = lb_score_cell(fd, kd.ints + (size_t)fd.int_row * st + g, kd.cache + (size_t)fd.cache_row * st + g, st, x, lf);
                                acc[jj] += pass ? -v : v;
                            }
                        }
                    }
                    unsigned hits = __ballot_sync(FULL, off >= 0);
                    while (hits) {
                        const int src = __ffs(hits) - 1;
                        hits &= hits - 1;
                        const float *t = tabg + __shfl_sync(FULL, off, src);
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj)
                            if (g0 + 32 * jj < st) acc[jj] += pass ? -t[32 * jj] : t[32 * jj];
                    }
                }
            }
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const int g = g0 + lane + 32 * jj;
                if (g < G) out[(r - b) * (uint64_t)G + g] += acc[jj];
            }
        }
    }
}

// the pos cells that sit in tared columns: count per row, scan, fill (once per loaded block)
__global__ void k_tpos_count(uint64_t R, const uint64_t *pos_ptr, const uint32_t *pos_feature, const uint8_t *tare_obs, uint64_t *cnt) {
    for (uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; r < R; r += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t n = 0;
        for (uint64_t i = pos_ptr[r]; i < pos_ptr[r + 1]; ++i) n += tare_obs[pos_feature[i]];
        cnt[r + 1] = n;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) cnt[0] = 0;
}
__global__ void k_tpos_fill(uint64_t R, const uint64_t *pos_ptr, const uint32_t *pos_feature, const uint32_t *pos_value,
                            const uint8_t *tare_obs, const uint64_t *tptr, uint32_t *tfeat, uint32_t *tval) {
    for (uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; r < R; r += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t o = tptr[r];
        for (uint64_t i = pos_ptr[r]; i < pos_ptr[r + 1]; ++i)
            if (tare_obs[pos_feature[i]]) { tfeat[o] = pos_feature[i]; tval[o] = pos_value[i]; ++o; }
    }
}
static int tared_pos_build(Engine *e) {
    if (e->tpos_valid) return 0;
    const uint64_t R = e->R;
    const char *diff = (const char *)e->d_diff;
    const uint8_t *tare_obs = (const uint8_t *)(diff + e->diff_off[0]);
    const uint64_t *pos_ptr = (const uint64_t *)(diff + e->diff_off[3]);
    const uint32_t *pos_feature = (const uint32_t *)(diff + e->diff_off[4]), *pos_value = (const uint32_t *)(diff + e->diff_off[5]);
    if (R + 1 > e->tpos_cap_rows) {
        if (e->d_tpos_ptr) cudaFree(e->d_tpos_ptr);
        e->d_tpos_ptr = nullptr; e->tpos_cap_rows = 0;
        LB_CUDA(cudaMalloc((void **)&e->d_tpos_ptr, 8 * (size_t)(R + 1)));
        e->tpos_cap_rows = R + 1;
    }
    const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((R + 255) / 256, (uint64_t)e->n_sm * 16));
    k_tpos_count<<<grid, 256, 0, e->stream>>>(R, pos_ptr, pos_feature, tare_obs, e->d_tpos_ptr);
    size_t scan_bytes = 0;
    LB_CUDA(cub::DeviceScan::InclusiveSum(nullptr, scan_bytes, e->d_tpos_ptr + 1, e->d_tpos_ptr + 1, (int)std::max<uint64_t>(R, 1), e->stream));
    void *d_scan = nullptr;
    LB_CUDA(cudaMallocAsync(&d_scan, scan_bytes + 256, e->stream));
    LB_CUDA(cub::DeviceScan::InclusiveSum(d_scan, scan_bytes, e->d_tpos_ptr + 1, e->d_tpos_ptr + 1, (int)R, e->stream));
    LB_CUDA(cudaFreeAsync(d_scan, e->stream));
    unsigned long long total = 0;
    LB_CUDA(cudaMemcpyAsync(&total, e->d_tpos_ptr + R, 8, cudaMemcpyDeviceToHost, e->stream));
    LB_CUDA(cudaStreamSynchronize(e->stream));
    if (total > e->tpos_cap || !e->d_tpos_feature) {
        if (e->d_tpos_feature) cudaFree(e->d_tpos_feature);
        if (e->d_tpos_value) cudaFree(e->d_tpos_value);
        e->d_tpos_feature = e->d_tpos_value = nullptr; e->tpos_cap = 0;
        LB_CUDA(cudaMalloc((void **)&e->d_tpos_feature, 4 * (size_t)std::max<unsigned long long>(total, 1)));
        LB_CUDA(cudaMalloc((void **)&e->d_tpos_value, 4 * (size_t)std::max<unsigned long long>(total, 1)));
        e->tpos_cap = (size_t)std::max<unsigned long long>(total, 1);
    }
    k_tpos_fill<<<grid, 256, 0, e->stream>>>(R, pos_ptr, pos_feature, pos_value, tare_obs, e->d_tpos_ptr, e->d_tpos_feature, e->d_tpos_value);
    LB_CUDA(cudaGetLastError());
    e->launches[LB_L_TOTAL] += 2; e->launches[LB_L_MISC] += 2;
    e->tpos_valid = true;
    return 0;
}

int lb_launch_score_rows_diff(Engine *e, int kid, uint64_t b, uint64_t en, float *d_out) {
    LB_TRY(tared_pos_build(e));
    const KindHost &k = e->kinds[kid];
    const int nf = (int)k.fids.size();
    std::vector<int32_t> row_base;
    int n_rows = 0;
    int32_t *d_row_base = nullptr;
    float *d_tab = nullptr;
    LB_TRY(score_table_build(e, kid, row_base, &n_rows, &d_row_base, &d_tab));
    // which columns the tare covers is host knowledge of the model file's tare row; mirror the device copy
    std::vector<uint8_t> tare_obs(e->F);
    const char *diff = (const char *)e->d_diff;
    LB_CUDA(cudaMemcpyAsync(tare_obs.data(), diff + e->diff_off[0], (size_t)e->F, cudaMemcpyDeviceToHost, e->stream));
    LB_CUDA(cudaStreamSynchronize(e->stream));
    std::vector<int32_t> flist;                                  // kind-local features the dense pass scores
    for (int j = 0; j < nf; ++j) {
        const int f = k.fids[j];
        if (!tare_obs[f]) flist.push_back(j);
        else if (e->specs[f].type == LB_NICH) return lb_fail("score_rows_diff: feature %d is real-valued and tared (src/differ.cc:121 never tares reals)", f);
    }
    // per-call scratch: [tare_scores Gcap f32][feature -> table row, F i32][tabulated values, F i32][dense feature list nf i32]
    const size_t o_frow = (4 * (size_t)k.Gcap + 255) / 256 * 256;
    const size_t o_fn = o_frow + (4 * (size_t)e->F + 255) / 256 * 256;
    const size_t o_flist = o_fn + (4 * (size_t)e->F + 255) / 256 * 256;
    const size_t need = o_flist + 4 * (size_t)std::max(nf, 1) + 256;
    if (need > e->tc_aux_bytes) {
        if (e->d_tc_aux) cudaFree(e->d_tc_aux);
        e->d_tc_aux = nullptr; e->tc_aux_bytes = 0;
        LB_CUDA(cudaMalloc(&e->d_tc_aux, need));
        e->tc_aux_bytes = need;
    }
    float *d_tare = (float *)e->d_tc_aux;
    int32_t *d_frow = (int32_t *)((char *)e->d_tc_aux + o_frow);
    int32_t *d_fn = (int32_t *)((char *)e->d_tc_aux + o_fn);
    int32_t *d_flist = (int32_t *)((char *)e->d_tc_aux + o_flist);
    const uint32_t *tare_val = (const uint32_t *)(diff + e->diff_off[1]);
    LB_CUDA(cudaMemsetAsync(d_frow, 0xFF, 4 * (size_t)e->F, e->stream));
    if (!flist.empty()) LB_CUDA(cudaMemcpyAsync(d_flist, flist.data(), 4 * flist.size(), cudaMemcpyHostToDevice, e->stream));
    k_tare_scores<<<1, 256, 0, e->stream>>>(e->d_kinds, kid, e->d_kind_feats, d_row_base, d_tab, (const uint8_t *)(diff + e->diff_off[0]),
                                            tare_val, d_tare, d_frow, d_fn, e->d_feats);
    // dense pass: prior + the untared columns, the tile kernel of K1 over the sub-list (an empty list writes the prior)
    {
        const uint64_t n_tiles = (en - b + SC_TR - 1) / SC_TR;
        const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>(n_tiles, (uint64_t)e->n_sm * 4));
        RowsDev rows{e->R, e->W, e->d_cat8, e->d_wide, e->d_mask};
        const int nl = (int)flist.size();
        for (int g_base = 0; g_base < k.G; g_base += 128) {
            const int left = k.G - g_base;
            if (left <= 32)
                k_score_rows<1><<<grid, SC_THREADS, 0, e->stream>>>(e->d_kinds, kid, e->d_feats, e->d_kind_feats, d_row_base, d_tab, rows, b, en,
                                                                   e->empty_group_count, g_base, d_out, 0, d_flist, nl);
            else if (left <= 64)
                k_score_rows<2><<<grid, SC_THREADS, 0, e->stream>>>(e->d_kinds, kid, e->d_feats, e->d_kind_feats, d_row_base, d_tab, rows, b, en,
                                                                   e->empty_group_count, g_base, d_out, 0, d_flist, nl);
            else
                k_score_rows<4><<<grid, SC_THREADS, 0, e->stream>>>(e->d_kinds, kid, e->d_feats, e->d_kind_feats, d_row_base, d_tab, rows, b, en,
                                                                   e->empty_group_count, g_base, d_out, 0, d_flist, nl);
            e->launches[LB_L_TOTAL]++; e->launches[LB_L_SCORE]++;
        }
    }
    const uint64_t n = en - b;
    const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((n * 32 + 255) / 256, (uint64_t)e->n_sm * 8));
    k_score_rows_diff<<<grid, 256, 0, e->stream>>>(
        e->d_kinds, kid, e->d_feats, d_frow, d_fn, d_tab, d_tare,
        e->diff_has_row_flags ? (const uint8_t *)(diff + e->diff_off[2]) : nullptr, tare_val,
        e->d_tpos_ptr, e->d_tpos_feature, e->d_tpos_value,
        (const uint64_t *)(diff + e->diff_off[6]), (const uint32_t *)(diff + e->diff_off[7]), b, en, d_out);
    LB_CUDA(cudaGetLastError());
    e->launches[LB_L_TOTAL] += 2; e->launches[LB_L_SCORE] += 2;
    return 0;
}

// ===========================================================================
// cached scorer refresh for a whole kind: thread per (cache row, group)
// ===========================================================================
__global__ void k_refresh(const KindDev *kinds, int kid, const FeatDev *feats, const int32_t *kind_feats,
                          const float *aux) {
    const KindDev &kd = kinds[kid];
    const int G = kd.G, st = kd.Gcap;
    const uint64_t total = (uint64_t)kd.n_cache_rows * G;
    for (uint64_t idx = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; idx < total + G;
         idx += (uint64_t)gridDim.x * blockDim.x) {
        if (idx >= total) {
            const int g = (int)(idx - total);
            const uint32_t n = kd.counts[g];
            kd.shifted[g] = n ? logf((float)n - kd.d) : 0.f;
            continue;
        }
        const int row = (int)(idx / G), g = (int)(idx % G);
        const FeatDev &f = feats[kind_feats[kd.feat_begin + kd.cache_row_feat[row]]];
        lb_refresh_row(f, aux, kd.ints + (size_t)f.int_row * st + g, kd.flts + (size_t)f.flt_row * st + g,
                       kd.cache + (size_t)f.cache_row * st + g, st, row - f.cache_row);
    }
}

int lb_launch_refresh(Engine *e, int kid) {
    const KindHost &k = e->kinds[kid];
    const uint64_t total = (uint64_t)k.n_cache_rows * k.G + k.G;
    const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((total + 255) / 256, (uint64_t)e->n_sm * 8));
    k_refresh<<<grid, 256, 0, e->stream>>>(e->d_kinds, kid, e->d_feats, e->d_kind_feats, e->d_aux);
    LB_CUDA(cudaGetLastError());
    e->launches[LB_L_TOTAL]++; e->launches[LB_L_REFRESH]++;
    return 0;
}
int lb_refresh_kind(Engine *e, int kid) { return lb_launch_refresh(e, kid); }

// ===========================================================================
// K2 / K4: bulk accumulate with given assignments (kernels in lb_accum.cuh).
// NICH / GP float statistics are accumulated as double sums (sum x, sum x^2,
// sum log x!) and converted back to (mean, count_times_variance) afterwards.
// ===========================================================================
__global__ void k_moments_to_sums(KindDev *kinds, int kid, const FeatDev *feats, const int32_t *kind_feats) {
    const KindDev &kd = kinds[kid];
    const int G = kd.G, st = kd.Gcap;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < kd.nf * G; idx += gridDim.x * blockDim.x) {
        const int j = idx / G, g = idx % G;
        const FeatDev &f = feats[kind_feats[kd.feat_begin + j]];
        if (f.type != LB_NICH) continue;
        const double n = (double)kd.ints[(size_t)f.int_row * st + g];
        double *fl = kd.flts + (size_t)f.flt_row * st + g;
        const double mean = fl[0], ctv = fl[st];
        fl[0] = n * mean;
        fl[st] = ctv + n * mean * mean;
    }
}
__global__ void k_sums_to_moments(KindDev *kinds, int kid, const FeatDev *feats, const int32_t *kind_feats) {
    const KindDev &kd = kinds[kid];
    const int G = kd.G, st = kd.Gcap;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < kd.nf * G; idx += gridDim.x * blockDim.x) {
        const int j = idx / G, g = idx % G;
        const FeatDev &f = feats[kind_feats[kd.feat_begin + j]];
        double *fl = kd.flts + (size_t)f.flt_row * st + g;
        if (f.type == LB_NICH) {
            const double n = (double)kd.ints[(size_t)f.int_row * st + g];
            if (n == 0.0) { fl[0] = 0.0; fl[st] = 0.0; }
            else {
                const double mean = fl[0] / n;
                fl[0] = mean;
                fl[st] = fmax(fl[st] - n * mean * mean, 0.0);
            }
        } else if (f.type == LB_GP) {
            if (kd.ints[(size_t)f.int_row * st + g] == 0u) fl[0] = 0.0;
        }
    }
}

// bytes per packed id for a kind with G groups
int lb_packed_width(int G) { return G < 255 ? 1 : (G < 65535 ? 2 : 4); }

// small device array of launch descriptors (AccTarget / PackSrc), reused
static int lb_upload_desc(Engine *e, const void *host, size_t bytes, void **dev) {
    if (bytes > e->d_desc_cap) {
        if (e->d_desc) cudaFree(e->d_desc);
        e->d_desc = nullptr;
        LB_CUDA(cudaMalloc(&e->d_desc, 2 * bytes + 4096));
        e->d_desc_cap = 2 * bytes + 4096;
    }
    // pageable source: the runtime stages it before returning, so `host` may die after the call
    LB_CUDA(cudaMemcpyAsync(e->d_desc, host, bytes, cudaMemcpyHostToDevice, e->stream));
    *dev = e->d_desc;
    return 0;
}

int lb_launch_pack_assign(Engine *e, const std::vector<PackSrc> &src, uint64_t b, uint64_t en, int launch_family) {
    if (en <= b || src.empty()) return 0;
    const unsigned n = (unsigned)src.size();
    const uint64_t want = (uint64_t)e->n_sm * 8 / n + 1;
    const unsigned pg = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((en - b + 255) / 256, want));
    const PackSrc *d_many = nullptr;
    if (n > 1) {
        void *d = nullptr;
        LB_TRY(lb_upload_desc(e, src.data(), sizeof(PackSrc) * n, &d));
        d_many = (const PackSrc *)d;
    }
    k_pack_assign<<<dim3(pg, n), 256, 0, e->stream>>>(src[0], d_many, b, en);
    LB_CUDA(cudaGetLastError());
    e->launches[LB_L_TOTAL]++; e->launches[launch_family]++;
    return 0;
}

// Launch k_accum_feature over `fids` (global feature ids) x `targets` (one per kind), grouped by
// shared-memory need so that small tables (BB / GP / NICH / low-dimensional DD) keep a high
// occupancy.  All targets go into ONE grid (blockIdx.z): a proposal round is a handful of
// launches, each many waves deep, instead of one under-filled launch per kind.
int lb_launch_accum_features(Engine *e, const FeatDev *d_feats, const std::vector<int> &fids,
                             const std::vector<AccTarget> &targets, uint64_t b, uint64_t en, int sign,
                             int launch_family) {
    if (en <= b || fids.empty() || targets.empty()) return 0;
    const size_t kBig = 200 * 1024, kSmall = 16 * 1024;
    LB_CUDA(cudaFuncSetAttribute(k_accum_feature, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBig));
    int G = 1;
    for (const AccTarget &t : targets) G = std::max(G, t.G);
    std::vector<int32_t> cls[3];
    size_t big_need = 0;
    for (int fid : fids) {
        const lb_feature_spec &s = e->specs[fid];
        const size_t need = lb_acc_smem_bytes(s.type, s.dim, G);
        if (need <= kSmall) cls[0].push_back(fid);
        else if (need <= kBig) { cls[1].push_back(fid); big_need = std::max(big_need, need); }
        else cls[2].push_back(fid);
    }
    const size_t total = cls[0].size() + cls[1].size() + cls[2].size();
    if ((int)total > e->d_fid_list_cap) {
        if (e->d_fid_list) cudaFree(e->d_fid_list);
        e->d_fid_list = nullptr;
        LB_CUDA(cudaMalloc((void **)&e->d_fid_list, 4 * (total + 64)));
        e->d_fid_list_cap = (int)total + 64;
    }
    const unsigned nt = (unsigned)targets.size();
    const AccTarget *d_many = nullptr;
    if (nt > 1) {
        void *d = nullptr;
        LB_TRY(lb_upload_desc(e, targets.data(), sizeof(AccTarget) * nt, &d));
        d_many = (const AccTarget *)d;
    }
    const uint64_t n = en - b;
    RowsDev rows{e->R, e->W, e->d_cat8, e->d_wide, e->d_mask};
    size_t off = 0;
    const size_t smem_of[3] = {kSmall, big_need, 0};
    for (int c = 0; c < 3; ++c) {
        if (cls[c].empty()) continue;
        // A CTA must see many more cells than its table has entries, or the flush dominates: big
        // (DD-256-sized) tables take >= 64k-row chunks, small ones >= 8k-row chunks.  Beyond that,
        // cut the rows only as far as needed to fill the machine a few times over.
        const uint64_t min_rows = c == 1 ? 65536 : 8192;
        const uint64_t pairs = (uint64_t)cls[c].size() * nt;
        const uint64_t want_chunks = std::max<uint64_t>(1, ((uint64_t)e->n_sm * 16 + pairs - 1) / pairs);
        uint64_t rows_per_cta = std::max<uint64_t>(min_rows, (n + want_chunks - 1) / want_chunks);
        rows_per_cta = (rows_per_cta + ACC_THREADS - 1) / ACC_THREADS * ACC_THREADS;
        const unsigned chunks = (unsigned)((n + rows_per_cta - 1) / rows_per_cta);
        int32_t *d_list = e->d_fid_list + off;
        LB_CUDA(cudaMemcpyAsync(d_list, cls[c].data(), 4 * cls[c].size(), cudaMemcpyHostToDevice, e->stream));
        dim3 grid(chunks, (unsigned)cls[c].size(), nt);
        k_accum_feature<<<grid, ACC_THREADS, smem_of[c], e->stream>>>(d_feats, d_list, rows, targets[0], d_many, b, en,
                                                                   rows_per_cta, sign, (int)smem_of[c]);
        LB_CUDA(cudaGetLastError());
        e->launches[LB_L_TOTAL]++; e->launches[launch_family]++;
        off += cls[c].size();
    }
    return 0;
}

int lb_launch_accumulate(Engine *e, int kid, uint64_t b, uint64_t en, int sign, bool sums_only) {
    const KindHost &k = e->kinds[kid];
    const int nf = (int)k.fids.size();
    const int items = std::max(1, nf * k.G);
    const int g1 = std::min((items + 255) / 256, e->n_sm * 4);
    if (nf) k_moments_to_sums<<<g1, 256, 0, e->stream>>>(e->d_kinds, kid, e->d_feats, e->d_kind_feats);
    if (en > b) {
        LB_TRY(lb_ensure_scratch(e, 4 * (size_t)e->R));
        void *d_packed = e->d_scratch;
        const int pg = (int)std::max<uint64_t>(1, std::min<uint64_t>((en - b + 255) / 256, (uint64_t)e->n_sm * 8));
        const int width = lb_packed_width(k.G);
        LB_TRY(lb_launch_pack_assign(e, {PackSrc{k.assign, k.g2p, d_packed, width}}, b, en, LB_L_ACCUM));
        // clustering.add_value
        unsigned long long *d_ss = (unsigned long long *)((char *)&e->d_kinds[kid] + offsetof(KindDev, sample_size));
        k_accum_counts<<<std::min(pg, e->n_sm * 2), ACC_THREADS, 4 * (size_t)k.G + 16, e->stream>>>(
            d_packed, width, b, en, sign, k.counts, d_ss, k.G, e->d_flag);
        e->launches[LB_L_TOTAL]++; e->launches[LB_L_ACCUM]++;
        LB_TRY(lb_launch_accum_features(e, e->d_feats, k.fids, {AccTarget{d_packed, k.ints, k.flts, k.Gcap, k.G, width}},
                                        b, en, sign, LB_L_ACCUM));
    }
    if (sums_only) { LB_CUDA(cudaGetLastError()); e->launches[LB_L_TOTAL] += nf ? 1 : 0; e->launches[LB_L_ACCUM] += nf ? 1 : 0; return 0; }
    if (nf) k_sums_to_moments<<<g1, 256, 0, e->stream>>>(e->d_kinds, kid, e->d_feats, e->d_kind_feats);
    LB_CUDA(cudaGetLastError());
    e->launches[LB_L_TOTAL] += nf ? 2 : 0; e->launches[LB_L_ACCUM] += nf ? 2 : 0;
    return lb_launch_refresh(e, kid);
}

// the second half of a row-sharded accumulate: the tables hold SUMS (sum x, sum x^2 for NICH; every other statistic
// is a sum anyway) that have been added up over the ranks; back to moments, cached scorers rebuilt
int lb_launch_sums_to_moments(Engine *e, int kid) {
    const KindHost &k = e->kinds[kid];
    const int nf = (int)k.fids.size();
    const int g1 = std::min((std::max(1, nf * k.G) + 255) / 256, e->n_sm * 4);
    if (nf) k_sums_to_moments<<<g1, 256, 0, e->stream>>>(e->d_kinds, kid, e->d_feats, e->d_kind_feats);
    LB_CUDA(cudaGetLastError());
    e->launches[LB_L_TOTAL] += nf ? 1 : 0; e->launches[LB_L_ACCUM] += nf ? 1 : 0;
    return lb_launch_refresh(e, kid);
}

// ===========================================================================
// table restride (capacity growth) and global-id renumbering
// ===========================================================================
template <class T>
__global__ void k_restride(const T *src, int src_st, T *dst, int dst_st, int rows, int G) {
    const uint64_t total = (uint64_t)rows * G;
    for (uint64_t idx = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; idx < total;
         idx += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t row = idx / G, g = idx % G;
        dst[row * dst_st + g] = src[row * src_st + g];
    }
}

int lb_launch_restride(Engine *e, const KindHost &s, KindHost &d) {
    auto grid = [&](uint64_t n) { return (int)std::max<uint64_t>(1, std::min<uint64_t>((n + 255) / 256, (uint64_t)e->n_sm * 8)); };
    const int G = s.G;
    k_restride<uint32_t><<<grid(G), 256, 0, e->stream>>>(s.counts, s.Gcap, d.counts, d.Gcap, 1, G);
    k_restride<float><<<grid(G), 256, 0, e->stream>>>(s.shifted, s.Gcap, d.shifted, d.Gcap, 1, G);
    k_restride<uint32_t><<<grid(G), 256, 0, e->stream>>>(s.p2g, s.Gcap, d.p2g, d.Gcap, 1, G);
    if (s.n_int_rows)
        k_restride<uint32_t><<<grid((uint64_t)s.n_int_rows * G), 256, 0, e->stream>>>(s.ints, s.Gcap, d.ints, d.Gcap, s.n_int_rows, G);
    if (s.n_flt_rows)
        k_restride<double><<<grid((uint64_t)s.n_flt_rows * G), 256, 0, e->stream>>>(s.flts, s.Gcap, d.flts, d.Gcap, s.n_flt_rows, G);
    if (s.n_cache_rows)
        k_restride<float><<<grid((uint64_t)s.n_cache_rows * G), 256, 0, e->stream>>>(s.cache, s.Gcap, d.cache, d.Gcap, s.n_cache_rows, G);
    LB_CUDA(cudaGetLastError());
    e->launches[LB_L_TOTAL] += 6; e->launches[LB_L_MISC] += 6;
    return 0;
}

__global__ void k_renumber_assign(uint32_t *assign, const uint32_t *g2p, uint64_t R) {
    for (uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; r < R; r += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t a = assign[r];
        if (a != LB_UNASSIGNED) assign[r] = g2p[a];
    }
}
__global__ void k_iota_ids(uint32_t *p2g, uint32_t *g2p, int G) {
    for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < G; g += gridDim.x * blockDim.x) {
        p2g[g] = (uint32_t)g;
        g2p[g] = (uint32_t)g;
    }
}

// global ids := packed ids (invisible outside: dumps remap ids anyway,
// src/cross_cat.cc:212-236), so the global->packed map never outgrows its buffer
int lb_launch_renumber(Engine *e, int kid) {
    KindHost &k = e->kinds[kid];
    const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((e->R + 255) / 256, (uint64_t)e->n_sm * 8));
    k_renumber_assign<<<grid, 256, 0, e->stream>>>(k.assign, k.g2p, e->R);
    LB_CUDA(cudaMemsetAsync(k.g2p, 0xFF, 4 * (size_t)k.g2p_cap, e->stream));
    k_iota_ids<<<(k.G + 255) / 256, 256, 0, e->stream>>>(k.p2g, k.g2p, k.G);
    LB_CUDA(cudaGetLastError());
    k.next_global = (uint32_t)k.G;
    e->launches[LB_L_TOTAL] += 2; e->launches[LB_L_MISC] += 2;
    return 0;
}

// ===========================================================================
// row validation: observed categorical values must index inside their tables
// ===========================================================================
__global__ void k_validate_rows(const FeatDev *feats, int F, RowsDev rows, int32_t *flag) {
    const int fid = blockIdx.y;
    const FeatDev &f = feats[fid];
    uint32_t limit;
    if (f.type == LB_BB) limit = 2u;
    else if (f.type == LB_DD || f.type == LB_DPD) limit = (uint32_t)f.dim;
    else return;
    for (uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; r < rows.R; r += (uint64_t)gridDim.x * blockDim.x) {
        if (!((rows.mask[(uint64_t)fid * rows.W + (r >> 5)] >> (r & 31)) & 1u)) continue;
        const uint32_t raw = f.is_wide ? rows.wide[(uint64_t)f.col * rows.R + r]
                                       : (uint32_t)rows.cat8[(uint64_t)f.col * rows.R + r];
        if (raw >= limit) *flag = 1;
    }
}

int lb_launch_validate_rows(Engine *e, int *bad) {
    LB_CUDA(cudaMemsetAsync(e->d_flag, 0, 4, e->stream));
    RowsDev rows{e->R, e->W, e->d_cat8, e->d_wide, e->d_mask};
    if (e->R) {
        dim3 grid((unsigned)std::min<uint64_t>((e->R + 255) / 256, 1024), (unsigned)e->F);
        k_validate_rows<<<grid, 256, 0, e->stream>>>(e->d_feats, e->F, rows, e->d_flag);
        LB_CUDA(cudaGetLastError());
        e->launches[LB_L_TOTAL]++; e->launches[LB_L_MISC]++;
    }
    LB_CUDA(cudaMemcpyAsync(bad, e->d_flag, 4, cudaMemcpyDeviceToHost, e->stream));
    LB_CUDA(cudaStreamSynchronize(e->stream));
    return 0;
}

// ===========================================================================
// score_data: sum over kinds / features / groups of log marginals
// ===========================================================================
__device__ double lb_marginal(const FeatDev &f, const float *aux, const uint32_t *in, const double *fl, int st) {
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

__device__ double block_sum(double v, double *s_red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0;
    if (threadIdx.x < 32) {
        t = threadIdx.x < (blockDim.x >> 5) ? s_red[threadIdx.x] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(FULL, t, o);
    }
    __syncthreads();
    return t;
}

// one block per kind; out[kid] = clustering score_counts + feature marginals
__global__ void __launch_bounds__(256)
k_score_data(const KindDev *kinds, const FeatDev *feats, const int32_t *kind_feats, const float *aux, double *out) {
    __shared__ double s_red[8];
    const KindDev &kd = kinds[blockIdx.x];
    const int G = kd.G, st = kd.Gcap;
    double acc = 0;
    if (kd.nf) {
        for (int idx = threadIdx.x; idx < kd.nf * G; idx += blockDim.x) {
            const int j = idx / G, g = idx % G;
            const FeatDev &f = feats[kind_feats[kd.feat_begin + j]];
            acc += lb_marginal(f, aux, kd.ints + (size_t)f.int_row * st + g, kd.flts + (size_t)f.flt_row * st + g, st);
        }
        // Pitman-Yor score_counts
        const double a = kd.alpha, d = kd.d;
        for (int g = threadIdx.x; g < G; g += blockDim.x) {
            const uint32_t n = kd.counts[g];
            if (n) acc += lgamma((double)n - d) - lgamma(1.0 - d);
        }
        if (threadIdx.x == 0) {
            int m = 0;
            double total = 0;
            for (int g = 0; g < G; ++g) if (kd.counts[g]) { acc += log(a + d * m); ++m; total += kd.counts[g]; }
            acc += lgamma(a) - lgamma(a + total);
        }
    }
    const double t = block_sum(acc, s_red);
    if (threadIdx.x == 0) out[blockIdx.x] = t;
}

int lb_launch_score_data(Engine *e, double *out) {
    const int K = (int)e->kinds.size();
    LB_TRY(lb_ensure_scratch(e, 8 * (size_t)K));
    k_score_data<<<K, 256, 0, e->stream>>>(e->d_kinds, e->d_feats, e->d_kind_feats, e->d_aux, (double *)e->d_scratch);
    LB_CUDA(cudaGetLastError());
    e->launches[LB_L_TOTAL]++; e->launches[LB_L_MISC]++;
    std::vector<double> part(K);
    LB_CUDA(cudaMemcpyAsync(part.data(), e->d_scratch, 8 * (size_t)K, cudaMemcpyDeviceToHost, e->stream));
    LB_CUDA(cudaStreamSynchronize(e->stream));
    double score = 0;
    // topology.score_counts over the feature counts of kinds that have features
    const double a = e->topo[0], d = e->topo[1];
    int m = 0;
    double total = 0;
    for (int k = 0; k < K; ++k) {
        const size_t nf = e->kinds[k].fids.size();
        if (!nf) continue;
        score += part[k];
        score += std::log(a + d * m) + std::lgamma((double)nf - d) - std::lgamma(1.0 - d);
        ++m; total += (double)nf;
    }
    score += std::lgamma(a) - std::lgamma(a + total);
    *out = score;
    return 0;
}
