// lb_sweep_spec.cuh -- K3s: the latency-optimised form of the sequential Gibbs sweep
// (CatKernel::process_add_task / process_remove_task, src/cat_kernel.hpp:199-299), included by
// lb_cat.cu after the scoring helpers.  Same chain, same draws, same tables as k_cat_sweep; what
// changes is WHEN the scores of an ADD are computed.
//
// An ADD's score vector depends on every earlier action of its kind, but each action changes ONE
// column (group) of the kind's tables.  So the feature part of the NEXT ADD's scores is gathered
// one action ahead, against the tables as they stand, while the current action samples; after the
// current action's update only the column it touched is recomputed for the pending row
// (lanes = features, one warp reduction) and written over the gathered value.  The gather (one
// L2 round trip over 2 x G floats per observed cell) leaves the dependent chain; what remains per
// action is   sample -> barrier -> update column g (+ recompute that column for the pending row)
// -> barrier.  Structure changes keep the pending vector valid the same way: a new group appends
// one recomputed column, a removed group copies the last column over the freed one.
//
// W = 8: warp 0 samples and keeps the books, warps 1..7 own class-pure feature stripes for the
//        gather, the update and the column recompute (the thread that updates a feature is the
//        thread that re-reads it: program order, no fence).
// W = 1: one warp is the whole CTA (kinds with fewer than SP_WIDE_NF features, featureless
//        ephemeral kinds included); SP_CPB such (chain, kind) pairs share a block and never
//        synchronise with each other.
//
// Exactness: scores differ from k_cat_sweep's only in the order of fp32 additions (the column
// recompute sums over lanes); the replay test checks every pick against the float64 oracle CDF.
#pragma once

constexpr int SP_RING = 8;        // ring slots of compiled rows
constexpr int SP_AHEAD = 3;       // the cells of action i + 3 are fetched while action i runs
constexpr int SP_LOOK = 2;        // the next ADD is scored ahead when it is at most this far away
constexpr int SP_CPB = 8;         // (chain, kind) pairs per block when W == 1
constexpr int SP_WIDE_NF = 8;     // kinds with at least this many features get a multi-warp CTA
constexpr int SP_WIDER_NF = 64;   // ... and with at least this many, 16 warps
constexpr int SP_HIST = 3;        // actions whose effect on assign[] may be newer than a ring prefetch (= SP_AHEAD)
constexpr int SP_MAXW = 16;
#ifndef LB_SPEC_WIDER
#define LB_SPEC_WIDER 16          // warps of the widest CTA class (kinds with >= SP_WIDER_NF features)
#endif

#ifndef LB_CO_GATHER
#define LB_CO_GATHER 1
#endif
#ifndef LB_CO_COLUMN
#define LB_CO_COLUMN 1
#endif
#ifndef LB_CO_PHASEB
#define LB_CO_PHASEB 1
#endif
#ifndef LB_CO_ADDGROUP
#define LB_CO_ADDGROUP 1
#endif
#ifndef LB_SP_PREFETCH
#define LB_SP_PREFETCH 1
#endif
#ifndef LB_SP_UNROLL
#define LB_SP_UNROLL 9
#endif
constexpr int SP_UNROLL = LB_SP_UNROLL;      // cells whose table rows a gathering warp requests before it uses the first

struct SweepPair { int32_t chain, kid; };

template <int PER> __device__ __forceinline__ void sp_ldv(const float *p, float (&o)[PER]);
template <> __device__ __forceinline__ void sp_ldv<1>(const float *p, float (&o)[1]) { o[0] = *p; }
template <> __device__ __forceinline__ void sp_ldv<2>(const float *p, float (&o)[2]) {
    const float2 t = *(const float2 *)p; o[0] = t.x; o[1] = t.y;
}
template <> __device__ __forceinline__ void sp_ldv<4>(const float *p, float (&o)[4]) {
    const float4 t = *(const float4 *)p; o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
}
template <int PER> __device__ __forceinline__ void sp_stv(float *p, const float (&v)[PER]);
template <> __device__ __forceinline__ void sp_stv<1>(float *p, const float (&v)[1]) { *p = v[0]; }
template <> __device__ __forceinline__ void sp_stv<2>(float *p, const float (&v)[2]) { *(float2 *)p = make_float2(v[0], v[1]); }
template <> __device__ __forceinline__ void sp_stv<4>(float *p, const float (&v)[4]) { *(float4 *)p = make_float4(v[0], v[1], v[2], v[3]); }

struct SpecSmem {
    int G, stop, pick, flag;
    uint32_t removed_global, next_global, epoch, pad0;
    unsigned long long sample_size;
    float u[32];                                   // uniforms of actions [base, base + 32), base = i & ~31
    uint32_t pf_assign[SP_RING], pf_pg[SP_RING], pf_epoch[SP_RING];
    float ovr[SP_MAXW];                            // per-warp sums of a recomputed column
    int own_off[SP_MAXW][N_CLASS + 1];
};
static_assert(sizeof(SpecSmem) <= 768, "SpecSmem must fit its 768-byte slot");

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

// softmax draw with the group scores in registers: lane owns the contiguous groups [lane * PER, lane * PER + PER).
// Same CDF order and tie rules as warp_sample (distributions::sample_from_scores_overwrite).
template <int PER>
__device__ __forceinline__ int warp_sample_regs(const float *feat, const float *prior, int G, float u, int lane,
                                                float *dbg, float dbg_shift, int dbg_stride) {
    const int lo = lane * PER;
    float v[PER];
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        const int g = lo + j;
        v[j] = g < G ? feat[g] + prior[g] : -INFINITY;
        m = fmaxf(m, v[j]);
    }
    if (dbg) {
#pragma unroll
        for (int j = 0; j < PER; ++j) if (lo + j < G && lo + j < dbg_stride) dbg[lo + j] = v[j] - dbg_shift;
    }
    m = warp_max(m);
    float local = 0.f;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        v[j] = lo + j < G ? __expf(v[j] - m) : 0.f;
        local += v[j];
    }
    float incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float t = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += t;
    }
    const float total = __shfl_sync(FULL, incl, 31);
    const float t = u * total;
    const unsigned nonzero = __ballot_sync(FULL, local > 0.f);
    const unsigned ballot = __ballot_sync(FULL, incl > t) & nonzero;
    const int src = ballot ? __ffs(ballot) - 1 : 31 - __clz(nonzero);
    int pick = lo;
    if (lane == src) {
        float c = incl - local;
        int last = lo;
        pick = -1;
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            if (lo + j < G && pick < 0) {
                if (v[j] > 0.f) last = lo + j;
                c += v[j];
                if (c > t) pick = lo + j;
            }
        }
        if (pick < 0) pick = last;
    }
    return __shfl_sync(FULL, pick, src);
}

// CATONLY: every kind of the launch holds Bernoulli / categorical columns only (C3): the gamma-Poisson and
// normal-inverse-chi-squared paths (lgamma, float64 moments) are compiled out, which takes the hot loop from ~280 KB
// of SASS to a fraction of that -- it has to stream through the instruction cache once per row action.
template <int W, bool CATONLY = false>
__global__ void __launch_bounds__(W == 1 ? 32 * SP_CPB : W * 32, 1)
k_cat_sweep_spec(const SweepChain *chains, const SweepPair *pairs, int n_pairs, int smem_per_pair) {
    constexpr int T = W * 32;
    constexpr int W0 = W > 1 ? 1 : 0;            // first warp that owns features
    extern __shared__ __align__(16) unsigned char spec_smem_all[];
    const int sub = W == 1 ? (int)(threadIdx.x >> 5) : 0;
    const int pair = W == 1 ? (int)blockIdx.x * SP_CPB + sub : (int)blockIdx.x;
    if (pair >= n_pairs) return;
    unsigned char *const smem_raw = spec_smem_all + (size_t)sub * smem_per_pair;
    const SweepChain &ch = chains[pairs[pair].chain];
    const int kid = pairs[pair].kid;
    KindDev &kd = ch.kinds[kid];
    const FeatDev *const feats = ch.feats;
    const int32_t *const kind_feats = ch.kind_feats;
    const float *const aux = ch.aux;
    const RowsDev rows = ch.rows;
    const uint32_t *const actions = ch.actions;
    const uint64_t n_actions = ch.n_actions, seed = ch.seed, offset = ch.offset;
    const int K = ch.K, n_empty = ch.n_empty, dbg_stride = ch.dbg_stride;
    uint32_t *const dbg_picks = ch.dbg_picks;
    float *const dbg_scores = ch.dbg_scores;
    const int tid = W == 1 ? (int)(threadIdx.x & 31) : (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nf = kd.nf, st = kd.Gcap;
    const int n_int_rows = kd.n_int_rows, n_flt_rows = kd.n_flt_rows, n_cache_rows = kd.n_cache_rows;
    const uint32_t zoff = (uint32_t)n_cache_rows * (uint32_t)st;      // the all-zero cache row
    const uint32_t g2p_cap = kd.g2p_cap;

    // Everything per-feature is indexed by POSITION k: features are reordered so that the positions a warp owns
    // for one class are contiguous (off[c] .. off[c + 1]); no indirection is left in the action loop.
    SpecSmem *ss = (SpecSmem *)smem_raw;
    FeatDev *s_feat = (FeatDev *)(smem_raw + 768);       // [nf] by position
    int32_t *s_fid = (int32_t *)(s_feat + nf);           // [nf] global feature id (mask row)
    uint32_t *s_Noff = (uint32_t *)(s_fid + nf);         // [nf] ints offset of the feature's total row (DD / DPD)
    float *s_A = (float *)(s_Noff + nf);                 // [nf] a_total
    uint32_t *s_voff = (uint32_t *)(s_A + nf);           // [SP_RING][nf] cache offset of the value row (zero row: unobserved)
    uint32_t *s_soff = s_voff + SP_RING * nf;            // [SP_RING][nf] cache offset of the shift row (zero row: none)
    uint32_t *s_ioff = s_soff + SP_RING * nf;            // [SP_RING][nf] ints offset of the count row the cell bumps
    float *s_av = (float *)(s_ioff + SP_RING * nf);      // [SP_RING][nf] the value's concentration a_v
    uint32_t *s_raw = (uint32_t *)(s_av + SP_RING * nf); // [SP_RING][nf]
    uint32_t *s_obs = s_raw + SP_RING * nf;              // [SP_RING][nf]
    float *s_part = (float *)(((uintptr_t)(s_obs + SP_RING * nf) + 15) & ~(uintptr_t)15);   // [W][st] per-warp partial feature scores of a gather (vector stores)
    float *s_a = s_part + W * st;                        // [st] feature scores of the ADD being sampled / pending
    float *s_b = s_a + st;                               // [st]
    float *s_score = s_b + st;                           // [st] scratch of the large-G sampler
    uint32_t *s_counts = (uint32_t *)(s_score + st);     // [st] clustering counts
    float *s_prior = (float *)(s_counts + st);           // [st] CRP prior per group, empties included
    uint32_t *s_p2g = (uint32_t *)(s_prior + st);        // [st] packed -> global
    uint16_t *s_own = (uint16_t *)(s_p2g + st);          // [nf] set-up only: kind-local feature index of position k
    uint8_t *s_owner = (uint8_t *)(s_own + nf);          // [nf] set-up only
    uint8_t *s_type = s_owner + nf;                      // [nf] set-up only: type by kind-local index

    uint32_t *const ints = kd.ints;
    double *const flts = kd.flts;
    float *const cache = kd.cache;
    uint32_t *const assign = kd.assign;
    uint32_t *const g2p = kd.g2p;
    const int32_t *const cache_row_feat = kd.cache_row_feat;
    const float alpha = kd.alpha, d = kd.d;

    for (int j = tid; j < nf; j += T) s_type[j] = (uint8_t)feats[kind_feats[kd.feat_begin + j]].type;
    {
        const int G0 = kd.G;
        const float sh_empty0 = logf((alpha + d * (float)(G0 - n_empty)) / (float)n_empty);
        for (int g = tid; g < st; g += T) {
            const uint32_t n = g < G0 ? kd.counts[g] : 0u;
            s_counts[g] = n;
            s_prior[g] = n ? logf((float)n - d) : sh_empty0;
            s_p2g[g] = kd.p2g[g];
        }
    }
    sweep_sync<W>();
    if (tid == 0) {
        ss->G = kd.G; ss->stop = 0; ss->epoch = 0; ss->flag = 0; ss->pick = 0;
        ss->next_global = kd.next_global; ss->sample_size = kd.sample_size;
        // features -> owner warps W0 .. W-1 by cost (GP 4, NICH 3, categorical 1); class-pure
        // stripes when there are enough warps, longest-processing-time first otherwise
        const float weight[N_CLASS] = {1.f, 4.f, 3.f};
        constexpr int NW = W - W0;
        int cnt[N_CLASS] = {0, 0, 0}, n_cls = 0;
        for (int j = 0; j < nf; ++j) cnt[feat_class(s_type[j])]++;
        for (int c = 0; c < N_CLASS; ++c) n_cls += cnt[c] > 0;
        if (NW >= 2 * n_cls && n_cls > 0) {
            int nw[N_CLASS] = {0, 0, 0}, left = NW;
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
            int w0 = W0, seen[N_CLASS] = {0, 0, 0}, first_warp[N_CLASS];
            for (int c = 0; c < N_CLASS; ++c) { first_warp[c] = w0; w0 += nw[c]; }
            for (int j = 0; j < nf; ++j) {
                const int c = feat_class(s_type[j]);
                s_owner[j] = (uint8_t)(first_warp[c] + seen[c]++ % nw[c]);
            }
        } else {
            float load[W];
            for (int w = 0; w < W; ++w) load[w] = 0.f;
            const int order[N_CLASS] = {CL_GP, CL_NICH, CL_CAT};
            for (int oc = 0; oc < N_CLASS; ++oc)
                for (int j = 0; j < nf; ++j) {
                    if (feat_class(s_type[j]) != order[oc]) continue;
                    int best = W0;
                    for (int w = W0 + 1; w < W; ++w) if (load[w] < load[best]) best = w;
                    s_owner[j] = (uint8_t)best;
                    load[best] += weight[order[oc]];
                }
        }
        int k = 0;
        for (int w = 0; w < W; ++w) {
            for (int c = 0; c < N_CLASS; ++c) {
                ss->own_off[w][c] = k;
                if (w >= W0)
                    for (int j = 0; j < nf; ++j)
                        if (s_owner[j] == w && feat_class(s_type[j]) == c) s_own[k++] = (uint16_t)j;
            }
            ss->own_off[w][N_CLASS] = k;
        }
    }
    sweep_sync<W>();
    for (int k = tid; k < nf; k += T) {
        const int fid = kind_feats[kd.feat_begin + s_own[k]];
        const FeatDev f = feats[fid];
        s_feat[k] = f;
        s_fid[k] = fid;
        s_Noff[k] = (uint32_t)(f.int_row + f.dim) * (uint32_t)st;
        s_A[k] = f.a_total;
    }
    int off[N_CLASS + 1];
#pragma unroll
    for (int c = 0; c <= N_CLASS; ++c) off[c] = ss->own_off[warp][c];
    sweep_sync<W>();

    uint64_t i = kd.progress;

    // ---- compiled rows: global -> registers (SP_AHEAD actions early) -> ring slot ----
    // Thread t owns positions t + q T.  A cell is fetched as (raw, observed) and "compiled" at commit time into
    // the offsets the action loop uses directly: value row / shift row of the cached scorer (the all-zero row
    // when there is nothing to add), the count row it bumps, its concentration.
    // (thread 0 also fetches the row's assignment; the dependent global -> packed lookup of a REMOVE is issued
    // at commit time, in phase B, so that it never delays the sampler)
    uint32_t pf_raw[SW_PF], pf_obs[SW_PF], pf_assign = LB_UNASSIGNED, pf_pg = LB_UNASSIGNED, pf_epoch = 0, pf_is_remove = 0;
    auto prefetch = [&](uint64_t idx, uint32_t a_) {      // a_ = actions[idx] (garbage when idx >= n_actions)
        const uint64_t r = a_ >> 1;
        const bool ok = idx < n_actions && r < rows.R;
        if (tid == 0) {
            pf_assign = LB_UNASSIGNED; pf_pg = LB_UNASSIGNED; pf_epoch = ss->epoch; pf_is_remove = 0;
            if (ok) { pf_assign = assign[r]; pf_is_remove = a_ & 1u; }
        }
#pragma unroll
        for (int q = 0; q < SW_PF; ++q) {
            if (q * T + warp * 32 >= nf) break;
            const int k = tid + q * T;
            pf_raw[q] = 0; pf_obs[q] = 0;
            if (k < nf && ok) {
                const FeatDev &f = s_feat[k];
                pf_obs[q] = (rows.mask[(uint64_t)s_fid[k] * rows.W + (r >> 5)] >> (r & 31)) & 1u;
                pf_raw[q] = f.is_wide ? rows.wide[(uint64_t)f.col * rows.R + r]
                                      : (uint32_t)rows.cat8[(uint64_t)f.col * rows.R + r];
            }
        }
    };
    // commit in two halves: `begin` compiles the cell, stores what needs no further load and ISSUES the loads the
    // rest depends on (the value's concentration, thread 0's global -> packed lookup); `end` stores those, a
    // phase later, so that no warp ever waits on them
    float pf_av[SW_PF];
    auto commit_begin = [&](int slot) {
#pragma unroll
        for (int q = 0; q < SW_PF; ++q) {
            if (q * T + warp * 32 >= nf) break;
            const int k = tid + q * T;
            pf_av[q] = 0.f;
            if (k < nf) {
                const FeatDev &f = s_feat[k];
                const uint32_t raw = pf_raw[q], obs = pf_obs[q];
                uint32_t vo = zoff, so = zoff, io = 0;
                if (obs) {
                    if (f.type == LB_BB) {
                        vo = (uint32_t)(f.cache_row + (raw ? 0 : 1)) * (uint32_t)st;
                    } else if (f.type == LB_DD || f.type == LB_DPD) {
                        vo = (uint32_t)(f.cache_row + (int)raw) * (uint32_t)st;
                        so = (uint32_t)(f.cache_row + f.dim) * (uint32_t)st;
                        io = (uint32_t)(f.int_row + (int)raw) * (uint32_t)st;
                        pf_av[q] = f.a_scale * aux[f.aux_off + (int)raw];
                    }
                }
                const int at = slot * nf + k;
                s_voff[at] = vo; s_soff[at] = so; s_ioff[at] = io; s_raw[at] = raw; s_obs[at] = obs;
#if LB_SP_PREFETCH
                // the value row this cell will gather SP_AHEAD actions from now: ask the L2 for it today.  With many
                // chains the tables (10 x 25 MB of value rows on C3) do not fit the L2 and a sweep's time follows its
                // DRAM bytes (8.7 s at 399 GB, 6.6 s at 223 GB): the misses cost latency inside the gather, not bandwidth.
                if (vo != zoff) {
                    const float *row = cache + vo;
                    const int Gn = ss->G;
                    for (int g0 = 0; g0 < Gn; g0 += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(row + g0));
#if LB_SP_PREFETCH >= 2
                    // ... and the count row the update will bump (its column is not known yet: the whole row).  Measured: 1 %
                    // faster, twice the DRAM traffic (434 GB against 223 GB per sweep) -- off by default
                    if (so != zoff) {
                        const uint32_t *crow = ints + io;
                        for (int g0 = 0; g0 < Gn; g0 += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(crow + g0));
                    }
#endif
                }
#endif
            }
        }
        if (tid == 0) pf_pg = (pf_is_remove && pf_assign != LB_UNASSIGNED) ? g2p[pf_assign] : LB_UNASSIGNED;
    };
    auto commit_end = [&](int slot) {
#pragma unroll
        for (int q = 0; q < SW_PF; ++q) {
            if (q * T + warp * 32 >= nf) break;
            const int k = tid + q * T;
            if (k < nf) s_av[slot * nf + k] = pf_av[q];
        }
        if (tid == 0) { ss->pf_assign[slot] = pf_assign; ss->pf_pg[slot] = pf_pg; ss->pf_epoch[slot] = pf_epoch; }
    };
    // Philox uniforms, 32 actions at a time across the lanes of warp 0
    auto draw_uniforms = [&](uint64_t base) {
        if (warp == 0) ss->u[lane] = lb_uniform(seed, 0u, offset + base + (uint64_t)lane, (uint32_t)kid);
    };
    auto action_at = [&](uint64_t idx) -> uint32_t { return idx < n_actions ? actions[idx] : 1u; };   // past the end: a REMOVE
    for (int q = 0; q < SP_AHEAD; ++q) {
        prefetch(i + q, action_at(i + q));
        commit_begin((int)((i + q) % SP_RING));
        commit_end((int)((i + q) % SP_RING));
    }
    draw_uniforms(i & ~31ull);
    sweep_sync<W>();

    // feature part of the scores of a compiled row (ring slot) against the tables as they stand: per-warp
    // partial sums into s_part.  A lane owns PER consecutive groups (the sampler's ownership), so one table row is ONE
    // vector load per warp for up to 128 groups, and SP_UNROLL cells' value and shift rows are all requested before
    // the first is used: a gather is one L2 round trip per SP_UNROLL cells of the warp, not one per cell.
    // (CATONLY: the GP / NICH loops get bounds the compiler can see are empty -- `if constexpr` inside the generic
    // lambda miscompiled the categorical loop above it with nvcc 12.9)
    const int gp_end = CATONLY ? off[CL_GP] : off[CL_GP + 1];
    const int nich_begin = CATONLY ? 0 : off[CL_NICH], nich_end = CATONLY ? 0 : off[CL_NICH + 1];
    auto gather_per = [&](int slot, int G, auto per_tag) {
        constexpr int PER = decltype(per_tag)::value;
        const uint32_t *rv = s_voff + slot * nf, *rs = s_soff + slot * nf, *rr = s_raw + slot * nf, *ro = s_obs + slot * nf;
        for (int g0 = 0; g0 < G; g0 += 32 * PER) {
            const int gl = g0 + PER * lane;
            if (gl >= G) continue;                         // this lane's groups do not exist: no table traffic for padding columns
            float acc[PER];
#pragma unroll
            for (int q = 0; q < PER; ++q) acc[q] = 0.f;
            const float *cg = cache + gl;
            const int cend = off[CL_CAT + 1];
            for (int k0 = off[CL_CAT]; k0 < cend; k0 += SP_UNROLL) {   // branch-free: absent cells read the zero row
                float v[SP_UNROLL][PER], sh[SP_UNROLL][PER];
#pragma unroll
                for (int j = 0; j < SP_UNROLL; ++j) {
                    const int k = k0 + j;
                    const uint32_t vo = k < cend ? rv[k] : zoff, so = k < cend ? rs[k] : zoff;
                    sp_ldv<PER>(cg + vo, v[j]);
                    sp_ldv<PER>(cg + so, sh[j]);
                }
#pragma unroll
                for (int j = 0; j < SP_UNROLL; ++j)
#pragma unroll
                    for (int q = 0; q < PER; ++q) acc[q] += v[j][q] - sh[j][q];
            }
            for (int k = off[CL_GP]; k < gp_end; ++k) {
                if (!ro[k]) continue;
                const FeatDev &f = s_feat[k];
                const uint32_t raw = rr[k];
                const float lf = lb_log_factorial(raw), x = (float)raw;
                const uint32_t *sum = ints + (size_t)(f.int_row + 1) * st + gl;
                const float *c0 = cg + (size_t)f.cache_row * st;
#pragma unroll
                for (int q = 0; q < PER; ++q)
                    acc[q] += lb_lgamma_diff_int(f.hp[0] + (float)sum[q], raw) - lf + c0[q] + x * c0[st + q];
            }
            for (int k = nich_begin; k < nich_end; ++k) {
                if (!ro[k]) continue;
                const FeatDev &f = s_feat[k];
                const float x = __uint_as_float(rr[k]);
                const float *c = cg + (size_t)f.cache_row * st;
#pragma unroll
                for (int q = 0; q < PER; ++q) {
                    const float dx = (x - c[3 * (size_t)st + q]) - c[4 * (size_t)st + q];
                    acc[q] += c[q] - c[st + q] * log1pf(c[2 * (size_t)st + q] * dx * dx);
                }
            }
            sp_stv<PER>(s_part + warp * st + gl, acc);
        }
    };
    auto gather = [&](int slot, int G) {
        if (warp < W0) return;
        if (G <= 32) gather_per(slot, G, std::integral_constant<int, 1>());
        else if (G <= 64) gather_per(slot, G, std::integral_constant<int, 2>());
        else gather_per(slot, G, std::integral_constant<int, 4>());
    };
    auto reduce_parts = [&](float *dst, int G) {
        const int G64 = min(st, (G + 63) & ~63);
        for (int g = tid; g < G64; g += T) {
            float s = 0.f;
#pragma unroll
            for (int w = W0; w < W; ++w) s += s_part[w * st + g];
            dst[g] = s;
        }
    };
    // this warp's share of the feature scores of a compiled row at ONE column, lanes = positions -> ss->ovr[warp]
    auto column_part = [&](int slot, int col) {
        const uint32_t *rv = s_voff + slot * nf, *rs = s_soff + slot * nf, *rr = s_raw + slot * nf, *ro = s_obs + slot * nf;
        float part = 0.f;
        for (int k = off[CL_CAT] + lane; k < off[CL_CAT + 1]; k += 32) part += cache[rv[k] + col] - cache[rs[k] + col];
        for (int k = off[CL_GP] + lane; k < (CATONLY ? off[CL_GP] : off[N_CLASS]); k += 32) {
            if (!ro[k]) continue;
            const FeatDev &f = s_feat[k];
            const uint32_t raw = rr[k];
            part += lb_score_cell(f, ints + (size_t)f.int_row * st + col, cache + (size_t)f.cache_row * st + col, st, raw,
                                  f.type == LB_GP ? lb_log_factorial(raw) : 0.f);
        }
        part = warp_sum_f(part);
        if (lane == 0) ss->ovr[warp] = part;
    };
    // warp 0, after a barrier: pending[col] = sum of the per-warp column parts (one lane, W - W0 independent loads)
    auto apply_column = [&](float *pending, int col, bool add) {
        if (tid == 0) {
            float v = 0.f;
#pragma unroll
            for (int w = W0; w < W; ++w) v += ss->ovr[w];
            pending[col] = add ? pending[col] + v : v;
        }
        if (warp == 0) __syncwarp();
    };

    float *s_cur = s_a, *s_next = s_b;
    uint64_t next_idx = ~0ull;        // ADD action whose feature scores sit in s_next (complete up to the last update)
    // thread 0: the last SP_HIST actions' effect on assign[] (newer than any ring prefetch)
    uint32_t h_row[SP_HIST], h_val[SP_HIST], h_pg[SP_HIST], h_epoch[SP_HIST];
#pragma unroll
    for (int q = 0; q < SP_HIST; ++q) { h_row[q] = 0xFFFFFFFFu; h_val[q] = LB_UNASSIGNED; h_pg[q] = 0; h_epoch[q] = 0; }
    // the action words of i .. i + 3 live in registers; the word of i + 4 is loaded one action early
    uint32_t act0 = action_at(i), act1 = action_at(i + 1), act2 = action_at(i + 2), act3 = action_at(i + 3);

    // make prof: cycles per phase seen by lane 0 of every warp (clock64), summed over the launch
    // ADD: 0 gather / sample | 1 wait at (A) | 2 phase-B work | 3 wait at (B) | 4 apply, structure
    // REMOVE: 5 lookup / gather | 6 wait at (A) + phase-B work | 7 wait at (B) + rest
#ifdef LB_SWEEP_PROFILE
    unsigned long long *const prof = ch.prof;
    long long pt[8] = {0, 0, 0, 0, 0, 0, 0, 0}, pc = clock64();
#define SP_TICK(k) do { if (prof && lane == 0) { const long long now__ = clock64(); pt[k] += now__ - pc; pc = now__; } } while (0)
#else
#define SP_TICK(k) do { } while (0)
#endif
    int status = LB_ST_OK;
    for (; i < n_actions; ++i) {
        const uint32_t act = act0;
        const uint32_t r = act >> 1;
        const bool is_remove = act & 1u;
        const int slot = (int)(i % SP_RING);
        if (r >= rows.R) { status = LB_ST_BAD_ROW; break; }
        const int G = ss->G;
        const uint32_t act4 = action_at(i + 4);          // consumed next iteration
        prefetch(i + SP_AHEAD, act3);

        // the next ADD after this action, if it is close enough to have its cells in the ring
        const bool have_p = !(act1 & 1u) ? (i + 1 < n_actions) : (!(act2 & 1u) && i + 2 < n_actions);
        const uint64_t p = !(act1 & 1u) ? i + 1 : i + 2;

        if (!is_remove) {
            if (G >= st || ss->next_global >= g2p_cap) { status = LB_ST_NEED_GROW; break; }
            if (next_idx != i) {
                // no scores were gathered ahead for this ADD (first action of a launch): gather them now
                gather(slot, G);
                sweep_sync<W>();
                reduce_parts(s_next, G);
                sweep_sync<W>();
            }
            float *t = s_cur; s_cur = s_next; s_next = t;
            next_idx = ~0ull;
        }
        const bool do_gather = have_p && next_idx != p;
        if (do_gather) gather((int)(p % SP_RING), G);

        if (!is_remove) {
            // ---- ADD: duplicate check, sample (warp 0) ----
            if (tid == 0) {
                uint32_t a = ss->pf_assign[slot];
#pragma unroll
                for (int q = 0; q < SP_HIST; ++q) if (h_row[q] == r) a = h_val[q];   // oldest first: the most recent wins
                if (a != LB_UNASSIGNED) ss->stop = LB_ST_DUPLICATE_ROW;
            }
            if (warp == 0) {
                const float u = ss->u[i & 31];
                float *dbg = dbg_scores ? dbg_scores + ((uint64_t)i * K + kid) * dbg_stride : nullptr;
                const float dbg_shift = dbg ? logf((float)ss->sample_size + alpha) : 0.f;
                int pick;
                if (G <= 32) pick = warp_sample_regs<1>(s_cur, s_prior, G, u, lane, dbg, dbg_shift, dbg_stride);
                else if (G <= 64) pick = warp_sample_regs<2>(s_cur, s_prior, G, u, lane, dbg, dbg_shift, dbg_stride);
                else if (G <= 128) pick = warp_sample_regs<4>(s_cur, s_prior, G, u, lane, dbg, dbg_shift, dbg_stride);
                else {
                    for (int g = lane; g < G; g += 32) {
                        const float s = s_cur[g] + s_prior[g];
                        s_score[g] = s;
                        if (dbg && g < dbg_stride) dbg[g] = s - dbg_shift;
                    }
                    __syncwarp();
                    pick = warp_sample(s_score, G, u, lane);
                }
                if (lane == 0) {
                    // clustering.add_value (cat_kernel.hpp:222-223); the prior entry is refreshed in phase B
                    const uint32_t n = s_counts[pick] + 1u;
                    ss->pick = pick;
                    ss->flag = n == 1u;
                    s_counts[pick] = n;
                    ss->sample_size += 1;
                }
            }
        } else if (tid == 0) {
            // ---- REMOVE: pop global id -> packed (cat_kernel.hpp:289-290) ----
            uint32_t a = ss->pf_assign[slot], pg = ss->pf_pg[slot], ep = ss->pf_epoch[slot];
#pragma unroll
            for (int q = 0; q < SP_HIST; ++q) if (h_row[q] == r) { a = h_val[q]; pg = h_pg[q]; ep = h_epoch[q]; }
            if (a == LB_UNASSIGNED) ss->stop = LB_ST_REMOVE_UNASSIGNED;
            else {
                const uint32_t g = (ep == ss->epoch && pg != LB_UNASSIGNED) ? pg : g2p[a];
                const uint32_t n = s_counts[g] - 1u;
                ss->pick = (int)g;
                ss->removed_global = a;
                ss->flag = n == 0u;
                s_counts[g] = n;
                ss->sample_size -= 1;
                assign[r] = LB_UNASSIGNED;
                if (dbg_picks) dbg_picks[(uint64_t)i * K + kid] = g;
            }
        }
        SP_TICK(is_remove ? 5 : 0);
        sweep_sync<W>();   // (A) pick visible; the gather's partial sums are in s_part
        if (ss->stop) { status = ss->stop; break; }
        const int g = ss->pick;
        const int flag = ss->flag;
        if (!is_remove) SP_TICK(1);

        // ---- phase B: apply the row to column g; bring the pending ADD's scores up to date ----
        // The pending row's score at column g changes by what THIS update changes: for every feature both rows
        // observe, the old entries of column g (requested together with the counts the update needs, one round trip)
        // against the new ones, which the update holds in registers.  Nothing is re-read after the stores.
        commit_begin((int)((i + SP_AHEAD) % SP_RING));
        if (do_gather) { reduce_parts(s_next, G); next_idx = p; }
        const bool pending = next_idx != ~0ull;
        const int nslot = (int)(next_idx % SP_RING);
        const bool patch = pending && !(flag && is_remove);
        {
            const int sign = is_remove ? -1 : +1;
            const uint32_t *rv = s_voff + slot * nf, *rs = s_soff + slot * nf, *ri = s_ioff + slot * nf;
            const uint32_t *rr = s_raw + slot * nf, *ro = s_obs + slot * nf;
            const float *ra = s_av + slot * nf;
            const uint32_t *pv = s_voff + nslot * nf, *pr = s_raw + nslot * nf, *po = s_obs + nslot * nf;
            float delta = 0.f;
            for (int k = off[CL_CAT] + lane; k < off[N_CLASS]; k += 32) {
                if (!ro[k]) continue;
                const bool both = patch && po[k];
                const uint32_t so = rs[k];
                if (so != zoff) {
                    // DD / DPD: Group::add_value / remove_value on the two count rows + their cached scorer entries
                    const uint32_t io = ri[k] + (uint32_t)g, no = s_Noff[k] + (uint32_t)g;
                    const uint32_t n0 = ints[io], t0 = ints[no];
                    float old_s = 0.f, pend_v = 0.f;
                    const bool same = both && pv[k] == rv[k];
                    if (both) old_s = cache[so + g];                    // same round trip as the counts
                    if (same) pend_v = cache[rv[k] + g];
                    const uint32_t n = n0 + (uint32_t)sign, tot = t0 + (uint32_t)sign;
                    const float new_v = logf(ra[k] + (float)n), new_s = logf(s_A[k] + (float)tot);
                    ints[io] = n;
                    ints[no] = tot;
                    cache[rv[k] + g] = new_v;
                    cache[so + g] = new_s;
                    if (both) delta += (same ? new_v - pend_v : 0.f) - (new_s - old_s);
                } else if constexpr (CATONLY && LB_CO_PHASEB) {
                    // Bernoulli: heads / tails and the two cached log-probabilities of this group
                    const FeatDev &f = s_feat[k];
                    uint32_t *in = ints + (size_t)f.int_row * st + g;
                    float *c = cache + (size_t)f.cache_row * st + g;
                    uint32_t h = in[0], t = in[st];
                    float old = 0.f;
                    if (both) old = c[pr[k] ? 0 : st];
                    if (rr[k]) { h += (uint32_t)sign; in[0] = h; } else { t += (uint32_t)sign; in[st] = t; }
                    float c_new[2];
                    lb_refresh_bb_v(f, h, t, c_new, 1);
                    c[0] = c_new[0]; c[st] = c_new[1];
                    if (both) delta += c_new[pr[k] ? 0 : 1] - old;
                } else {
                    const FeatDev &f = s_feat[k];
                    const uint32_t praw = pr[k];
                    const float lf = both && f.type == LB_GP ? lb_log_factorial(praw) : 0.f;
                    uint32_t *in = ints + (size_t)f.int_row * st + g;
                    float *c = cache + (size_t)f.cache_row * st + g;
                    float old = 0.f;
                    if (both) old = lb_score_cell(f, in, c, st, praw, lf);
                    uint32_t in_new[2];
                    float c_new[5];
                    lb_update_cell_x(f, aux, in, flts + (size_t)f.flt_row * st + g, c, st, rr[k], sign, in_new, c_new);
                    if (both) delta += lb_score_cell(f, in_new, c_new, 1, praw, lf) - old;
                }
            }
            if (patch) {
                delta = warp_sum_f(delta);
                if (lane == 0) ss->ovr[warp] = delta;
            }
        }
        commit_end((int)((i + SP_AHEAD) % SP_RING));
        if (warp == 0) {
            if ((i & 31) == 31) draw_uniforms(i + 1);
            if (lane == 0) {
                const uint32_t n = s_counts[g];
                if (n) s_prior[g] = logf((float)n - d);
                uint32_t gid = LB_UNASSIGNED;
                if (!is_remove) {
                    gid = s_p2g[g];
                    assign[r] = gid;
                    if (dbg_picks) dbg_picks[(uint64_t)i * K + kid] = (uint32_t)g;
                }
#pragma unroll
                for (int q = 0; q < SP_HIST - 1; ++q) { h_row[q] = h_row[q + 1]; h_val[q] = h_val[q + 1]; h_pg[q] = h_pg[q + 1]; h_epoch[q] = h_epoch[q + 1]; }
                h_row[SP_HIST - 1] = r; h_val[SP_HIST - 1] = gid; h_pg[SP_HIST - 1] = (uint32_t)g; h_epoch[SP_HIST - 1] = ss->epoch;
            }
        }
        act0 = act1; act1 = act2; act2 = act3; act3 = act4;
        SP_TICK(is_remove ? 6 : 2);
        sweep_sync<W>();   // (B) updates visible to the next gather; column parts in ss->ovr
        if (!is_remove) SP_TICK(3);
        if (patch) apply_column(s_next, g, true);

        if (flag) {
            if (!is_remove) {
                // add_group (product_mixture.cc:178-183): append a fresh empty group
                const int ng = G;
                for (int row = tid; row < n_int_rows; row += T) ints[(size_t)row * st + ng] = 0u;
                for (int row = tid; row < n_flt_rows; row += T) flts[(size_t)row * st + ng] = 0.0;
                sweep_sync<W>();
                for (int row = tid; row < n_cache_rows; row += T) {
                    const FeatDev f = feats[kind_feats[kd.feat_begin + cache_row_feat[row]]];
                    if constexpr (CATONLY && LB_CO_ADDGROUP) {       // a fresh group: every count is zero
                        const int local = row - f.cache_row;
                        float v;
                        if (f.type == LB_BB) v = logf((local ? f.hp[1] : f.hp[0]) / (f.hp[0] + f.hp[1]));
                        else v = logf(local < f.dim ? f.a_scale * aux[f.aux_off + local] : f.a_total);
                        cache[(size_t)row * st + ng] = v;
                    } else {
                        lb_refresh_row(f, aux, ints + (size_t)f.int_row * st + ng, flts + (size_t)f.flt_row * st + ng,
                                       cache + (size_t)f.cache_row * st + ng, st, row - f.cache_row);
                    }
                }
                if (tid == 0) {
                    s_counts[ng] = 0u;
                    const uint32_t gid = ss->next_global;
                    s_p2g[ng] = gid;
                    g2p[gid] = (uint32_t)ng;
                    ss->next_global = gid + 1;
                    ss->G = ng + 1;
                }
                sweep_sync<W>();
                {
                    const float she = logf((alpha + d * (float)(ng + 1 - n_empty)) / (float)n_empty);
                    for (int gg = tid; gg <= ng; gg += T) if (!s_counts[gg]) s_prior[gg] = she;
                }
                if (pending) {   // the pending ADD sees one more column
                    column_part(nslot, ng);
                    sweep_sync<W>();
                    apply_column(s_next, ng, false);
                }
                sweep_sync<W>();
            } else {
                // remove_group (product_mixture.cc:233-238): swap-with-last
                const int last = G - 1;
                if (g != last) {
                    for (int row = tid; row < n_int_rows; row += T)
                        ints[(size_t)row * st + g] = ints[(size_t)row * st + last];
                    for (int row = tid; row < n_flt_rows; row += T)
                        flts[(size_t)row * st + g] = flts[(size_t)row * st + last];
                    for (int row = tid; row < n_cache_rows; row += T)
                        cache[(size_t)row * st + g] = cache[(size_t)row * st + last];
                }
                if (tid == 0) {
                    g2p[ss->removed_global] = LB_UNASSIGNED;
                    if (g != last) {
                        s_counts[g] = s_counts[last];
                        s_prior[g] = s_prior[last];
                        const uint32_t gid = s_p2g[last];
                        s_p2g[g] = gid;
                        g2p[gid] = (uint32_t)g;
                        ss->epoch += 1;
                        if (pending) s_next[g] = s_next[last];   // the pending ADD's column moves with the group
                    }
                    s_counts[last] = 0u;
                    ss->G = last;
                }
                sweep_sync<W>();
                {
                    const float she = logf((alpha + d * (float)(last - n_empty)) / (float)n_empty);
                    for (int gg = tid; gg < last; gg += T) if (!s_counts[gg]) s_prior[gg] = she;
                }
                sweep_sync<W>();
            }
        }
        SP_TICK(is_remove ? 7 : 4);
    }
#ifdef LB_SWEEP_PROFILE
    if (prof && lane == 0 && kid < 16 && W > 1 && warp < 8)
        for (int k = 0; k < 8; ++k) atomicAdd(&prof[(kid * 8 + warp) * 8 + k], (unsigned long long)pt[k]);
#endif
#undef SP_TICK
    sweep_sync<W>();
    for (int g = tid; g < st; g += T) {
        const uint32_t n = s_counts[g];
        kd.counts[g] = n;
        kd.shifted[g] = n ? s_prior[g] : 0.f;
        kd.p2g[g] = s_p2g[g];
    }
    if (tid == 0) {
        kd.G = ss->G;
        kd.sample_size = ss->sample_size;
        kd.next_global = ss->next_global;
        kd.status = status;
        kd.progress = i;
    }
}

static size_t spec_smem_bytes(int nf, int Gcap, int warps) {
    return 768 + (size_t)nf * (sizeof(FeatDev) + 12 + 24 * SP_RING + 4) + (size_t)Gcap * 4 * (warps + 6) + 64 + 16;
}
