// lb_sweep_flat.cuh -- K3f: the sweep of a FEATURELESS kind (the ephemeral kinds of the kind kernel,
// src/kind_kernel.cc:159-227: 32 of a C3 chain's 41 kinds), included by lb_cat.cu.
//
// Without features, CatKernel::process_add_task (src/cat_kernel.hpp:199-224) samples from the clustering prior alone,
// and the Pitman-Yor prior is LINEAR in the counts: weight n_g - d for an occupied group, (alpha + d m) / e for an
// empty one (m occupied groups, e empty), total N + alpha in closed form.  So no logarithm, exponential or max is
// needed: the draw is an inverse CDF over integer counts.  A lane owns a contiguous block of groups and keeps the
// block's (sum of counts, number of occupied groups) as integers in registers; a draw is one warp scan over the 32
// block weights plus a walk inside one block; an update touches one shared-memory count and one lane's registers.
// The general kernel spends a score vector, a max, an exp and a scan of G terms per ADD on the same kinds, and the grid
// draws of (alpha, d) give some of them thousands of groups (alpha = 50, d = 0.36, 10^6 rows: ~4800) -- they, not the
// kinds with features, set the length of a C3 sweep.
//
// One warp per (chain, kind), several to a block; the chain, its draws (Philox keyed by seed / action / kind) and the
// tables it leaves (counts, id maps, assignments, `shifted` = log(n - d)) are the general kernel's.  Exactly one empty
// group (loom's default empty_group_count; other settings take the general kernel), no score taps.
#pragma once

__global__ void __launch_bounds__(256)
k_cat_sweep_flat(const SweepChain *chains, const SweepPair *pairs, int n_pairs, int smem_per_pair) {
    extern __shared__ __align__(16) unsigned char flat_smem_all[];
    const int sub = (int)(threadIdx.x >> 5), lane = (int)(threadIdx.x & 31);
    const int pair = (int)blockIdx.x * (int)(blockDim.x >> 5) + sub;
    if (pair >= n_pairs) return;
    const SweepChain &ch = chains[pairs[pair].chain];
    const int kid = pairs[pair].kid;
    KindDev &kd = ch.kinds[kid];
    const int st = kd.Gcap, K = ch.K;
    uint32_t *s_counts = (uint32_t *)(flat_smem_all + (size_t)sub * smem_per_pair);
    uint32_t *s_p2g = s_counts + st;
    uint32_t *const assign = kd.assign, *const g2p = kd.g2p;
    uint32_t *const dbg_picks = ch.dbg_picks;
    const uint32_t *const actions = ch.actions;
    const uint64_t n_actions = ch.n_actions, seed = ch.seed, offset = ch.offset, R = ch.rows.R;
    const float alpha = kd.alpha, d = kd.d;
    const uint32_t g2p_cap = kd.g2p_cap;
    const int PER = (st + 31) / 32;                         // groups per lane block
    const int lo = lane * PER, hi = min(st, lo + PER);

    int G = kd.G;
    uint32_t next_global = kd.next_global;
    unsigned long long N = kd.sample_size;
    // this lane's block: sum of counts, occupied groups; the one empty group
    unsigned long long I = 0;
    int O = 0, ge_local = 0x7fffffff;
    for (int g = lo; g < hi; ++g) {
        const uint32_t n = g < G ? kd.counts[g] : 0u;
        s_counts[g] = n;
        s_p2g[g] = kd.p2g[g];
        I += n;
        O += n != 0u;
        if (g < G && n == 0u) ge_local = min(ge_local, g);
    }
    int m = O, ge = ge_local;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m += __shfl_xor_sync(FULL, m, o);
        ge = min(ge, __shfl_xor_sync(FULL, ge, o));
    }
    __syncwarp();

    uint64_t i = kd.progress;
    int status = LB_ST_OK;
    uint32_t acts = 1u, pa = LB_UNASSIGNED, pg = LB_UNASSIGNED;     // this lane's action of the current batch + prefetched ids
    float us = 0.f;
    bool stale = true;
    uint64_t batch = ~0ull;
    uint32_t h_r = 0xFFFFFFFFu, h_a = LB_UNASSIGNED;          // the last row whose assignment this sweep wrote, and what it wrote
    for (; i < n_actions; ++i) {
        if ((i & ~31ull) != batch) {
            // 32 actions at a time: lane j holds action batch + j, its uniform, and (REMOVE) the row's global and packed id
            batch = i & ~31ull;
            const uint64_t idx = batch + (uint64_t)lane;
            acts = idx < n_actions ? actions[idx] : 1u;
            us = lb_uniform(seed, 0u, offset + idx, (uint32_t)kid);
            pa = LB_UNASSIGNED; pg = LB_UNASSIGNED; stale = true;
            if (idx >= i && idx < n_actions && (acts >> 1) < R) {
                pa = assign[acts >> 1];
                if ((acts & 1u) && pa != LB_UNASSIGNED) pg = g2p[pa];
                stale = false;
            }
        }
        const int j = (int)(i & 31);
        const uint32_t act = __shfl_sync(FULL, acts, j);
        const uint32_t r = act >> 1;
        if (r >= R) { status = LB_ST_BAD_ROW; break; }
        uint32_t a = __shfl_sync(FULL, pa, j), gpk = __shfl_sync(FULL, pg, j);
        if (__shfl_sync(FULL, (int)stale, j)) {              // the row or the id map changed inside this batch: read again
            a = r == h_r ? h_a : assign[r];                    // (REMOVE r, ADD r: the id was cleared one action ago)
            gpk = (act & 1u) && a != LB_UNASSIGNED ? g2p[a] : LB_UNASSIGNED;
        }
        if (!(act & 1u)) {
            // ---- ADD (cat_kernel.hpp:199-224 with an empty feature list) ----
            if (G >= st || next_global >= g2p_cap) { status = LB_ST_NEED_GROW; break; }
            if (a != LB_UNASSIGNED) { status = LB_ST_DUPLICATE_ROW; break; }
            const float we = alpha + d * (float)m;              // the empty group's weight (one empty group)
            const float B = ((float)I - d * (float)O) + (ge >= lo && ge < hi ? we : 0.f);
            float incl = B;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const float t = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += t;
            }
            const float total = __shfl_sync(FULL, incl, 31);
            const float t = __shfl_sync(FULL, us, j) * total;
            const unsigned nonzero = __ballot_sync(FULL, B > 0.f);
            const unsigned over = __ballot_sync(FULL, incl > t) & nonzero;
            const int src = over ? __ffs(over) - 1 : 31 - __clz(nonzero);
            // second level: the 32 lanes share the chosen block, P2 consecutive groups each
            const int blo = src * PER, bhi = min(st, blo + PER);
            const float base = __shfl_sync(FULL, incl - B, src);
            const int P2 = (PER + 31) / 32;
            const int l2 = min(bhi, blo + lane * P2), h2 = min(bhi, l2 + P2);
            float w2 = 0.f;
            for (int g = l2; g < h2; ++g) {
                const uint32_t n = s_counts[g];
                w2 += n ? (float)n - d : (g == ge ? we : 0.f);
            }
            float incl2 = w2;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const float x = __shfl_up_sync(FULL, incl2, o);
                if (lane >= o) incl2 += x;
            }
            const unsigned nz2 = __ballot_sync(FULL, w2 > 0.f);
            const unsigned over2 = __ballot_sync(FULL, base + incl2 > t) & nz2;
            const int src2 = over2 ? __ffs(over2) - 1 : 31 - __clz(nz2);
            int pick = -1;
            if (lane == src2) {
                float c = base + incl2 - w2;
                int last = l2;
                for (int g = l2; g < h2 && pick < 0; ++g) {
                    const uint32_t n = s_counts[g];
                    const float w = n ? (float)n - d : (g == ge ? we : 0.f);
                    if (w > 0.f) last = g;
                    c += w;
                    if (c > t) pick = g;
                }
                if (pick < 0) pick = last;
            }
            pick = __shfl_sync(FULL, pick, src2);
            const bool was_empty = pick == ge;
            if (pick >= lo && pick < hi) { s_counts[pick] += 1u; I += 1; O += was_empty; }
            N += 1;
            const uint32_t gid = s_p2g[pick];
            if (lane == 0) {
                assign[r] = gid;
                if (dbg_picks) dbg_picks[i * (uint64_t)K + kid] = (uint32_t)pick;
            }
            h_r = r; h_a = gid;
            if (was_empty) {                                   // add_group (product_mixture.cc:178-183): a fresh empty group at the end
                m += 1;
                if (lane == 0) { s_counts[G] = 0u; s_p2g[G] = next_global; g2p[next_global] = (uint32_t)G; }
                ge = G;
                next_global += 1;
                G += 1;
            }
            if ((acts >> 1) == r && lane > j) stale = true;     // a later action of this batch names the row just assigned
            __syncwarp();
        } else {
            // ---- REMOVE (cat_kernel.hpp:277-299) ----
            if (a == LB_UNASSIGNED || gpk == LB_UNASSIGNED) { status = LB_ST_REMOVE_UNASSIGNED; break; }
            const int g = (int)gpk;
            uint32_t n_after = 0;
            if (g >= lo && g < hi) { n_after = s_counts[g] - 1u; s_counts[g] = n_after; I -= 1; }
            n_after = __shfl_sync(FULL, n_after, g / PER);
            N -= 1;
            if (lane == 0) {
                assign[r] = LB_UNASSIGNED;
                if (dbg_picks) dbg_picks[i * (uint64_t)K + kid] = (uint32_t)g;
            }
            h_r = r; h_a = LB_UNASSIGNED;
            if (n_after == 0u) {                               // remove_group (product_mixture.cc:233-238): swap-with-last
                const int last = G - 1;
                if (g >= lo && g < hi) O -= 1;
                __syncwarp();
                const uint32_t cl = s_counts[last], gid_l = s_p2g[last];
                __syncwarp();
                if (lane == 0) g2p[a] = LB_UNASSIGNED;
                if (g != last) {
                    if (last == ge) ge = g;                    // the empty group moves into the freed slot
                    else {
                        if (last >= lo && last < hi) { I -= cl; O -= 1; }
                        if (g >= lo && g < hi) { I += cl; O += 1; }
                    }
                    if (lane == 0) { s_counts[g] = cl; s_counts[last] = 0u; s_p2g[g] = gid_l; g2p[gid_l] = (uint32_t)g; }
                    stale = true;                              // packed ids moved: every prefetched lookup of this batch is suspect
                }
                m -= 1;
                G -= 1;
            }
            if ((acts >> 1) == r && lane > j) stale = true;
            __syncwarp();
        }
    }
    __syncwarp();
    for (int g = lo; g < hi; ++g) {
        const uint32_t n = s_counts[g];
        kd.counts[g] = n;
        kd.shifted[g] = n ? logf((float)n - d) : 0.f;
        kd.p2g[g] = s_p2g[g];
    }
    if (lane == 0) {
        kd.G = G;
        kd.sample_size = N;
        kd.next_global = next_global;
        kd.status = status;
        kd.progress = i;
    }
}
