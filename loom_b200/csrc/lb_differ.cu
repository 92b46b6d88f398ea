// lb_differ.cu -- loom's tare and sparsify passes (loom::Differ, src/differ.hpp, src/differ.cc) as a
// columnar pre-pass over the device-resident row block (sm_100a).
//
//   add_rows      per column a 16-bin histogram of the observed raw values (BooleanSummary /
//                 CountSummary, src/differ.hpp:49-95): one CTA per (row tile, column), the bins live in
//                 shared memory, a warp merges equal values with match_any before it touches a bin --
//                 the interesting columns are exactly the ones where nearly every row holds the same value
//   compress_rows _abs_to_rel (src/differ.cc:248-298): thread per row walks the columns (adjacent threads
//                 read adjacent rows of a column: coalesced), counts its pos / neg entries, an exclusive
//                 scan turns counts into CSR offsets, a second walk (a warp per 32 rows, cells transposed through
//                 shared memory so that a row's entries are stored by consecutive lanes) writes the entries
// Both are HBM streams over the block: 1 B (BB, DD) or 4 B (DPD, GP) per cell + 1 mask bit; reals are
// only read by compress_rows.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cub/device/device_scan.cuh>

#include "lb_internal.h"

#define DIFFER_BINS 16   // CountSummary::max_count (src/differ.hpp:65); booleans use bins 0 and 1

struct DifferState {
    unsigned long long *d_hist = nullptr;   // [F][16]
    int F = 0;
    uint64_t rows_added = 0;
    // the last compress_rows result
    unsigned long long *d_pos_ptr = nullptr, *d_neg_ptr = nullptr;   // [R + 1]
    uint32_t *d_pos_feature = nullptr, *d_pos_value = nullptr, *d_neg_feature = nullptr;
    uint64_t rows = 0, n_pos = 0, n_neg = 0;
    uint64_t cap_rows = 0, cap_pos = 0, cap_neg = 0;    // buffers are kept between calls and only grow
    bool have_diff = false;
};

static void free_diff(DifferState *s) {
    cudaFree(s->d_pos_ptr); cudaFree(s->d_neg_ptr);
    cudaFree(s->d_pos_feature); cudaFree(s->d_pos_value); cudaFree(s->d_neg_feature);
    s->d_pos_ptr = s->d_neg_ptr = nullptr;
    s->d_pos_feature = s->d_pos_value = s->d_neg_feature = nullptr;
    s->cap_rows = s->cap_pos = s->cap_neg = 0;
    s->have_diff = false;
}

void lb_release_differ(Engine *e) {
    DifferState *s = (DifferState *)e->differ;
    if (!s) return;
    cudaFree(s->d_hist);
    free_diff(s);
    delete s;
    e->differ = nullptr;
}

// dict_off[f] = first entry of feature f's dictionary in dict_values, or -1 when the cell IS the raw value
__global__ void k_differ_hist(const FeatDev *feats, RowsDev rows, const int32_t *dict_off, const uint32_t *dict_values,
                              unsigned long long *hist) {
    const int f = blockIdx.y;
    const FeatDev fd = feats[f];
    if (fd.type == LB_NICH) return;                        // "do not sparsify reals" (src/differ.cc:121)
    __shared__ unsigned int s_hist[DIFFER_BINS];
    if (threadIdx.x < DIFFER_BINS) s_hist[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t *mask = rows.mask + (uint64_t)f * rows.W;
    const int32_t off = dict_off[f];
    const unsigned lane = threadIdx.x & 31;
    for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x; base < rows.R; base += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = base + threadIdx.x;
        uint32_t bin = DIFFER_BINS;
        if (r < rows.R && ((mask[r >> 5] >> (r & 31)) & 1u)) {
            uint32_t v = fd.is_wide ? rows.wide[(uint64_t)fd.col * rows.R + r] : (uint32_t)rows.cat8[(uint64_t)fd.col * rows.R + r];
            if (off >= 0) v = dict_values[off + (int32_t)v];
            if (v < DIFFER_BINS) bin = v;
        }
        const unsigned peers = __match_any_sync(0xffffffffu, bin);
        if (bin < DIFFER_BINS && lane == (unsigned)(__ffs(peers) - 1)) atomicAdd(&s_hist[bin], (unsigned)__popc(peers));
    }
    __syncthreads();
    if (threadIdx.x < DIFFER_BINS && s_hist[threadIdx.x])
        atomicAdd(&hist[(size_t)f * DIFFER_BINS + threadIdx.x], (unsigned long long)s_hist[threadIdx.x]);
}

// One row against the tare.  FILL = false: count the entries (written to ptr[r + 1]); FILL = true: write them at ptr[r].
template <bool FILL>
__global__ void k_differ_rows(const FeatDev *feats, int F, RowsDev rows, const uint8_t *tare_obs, const uint32_t *tare_val,
                              unsigned long long *pos_ptr, unsigned long long *neg_ptr, uint32_t *pos_feature,
                              uint32_t *pos_value, uint32_t *neg_feature) {
    for (uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; r < rows.R; r += (uint64_t)gridDim.x * blockDim.x) {
        unsigned long long np = FILL ? pos_ptr[r] : 0ull, nn = FILL ? neg_ptr[r] : 0ull;
        for (int f = 0; f < F; ++f) {
            const bool obs = (rows.mask[(uint64_t)f * rows.W + (r >> 5)] >> (r & 31)) & 1u;
            uint32_t v = 0;
            if (obs) {
                const int col = feats[f].col;
                v = feats[f].is_wide ? rows.wide[(uint64_t)col * rows.R + r] : (uint32_t)rows.cat8[(uint64_t)col * rows.R + r];
            }
            bool to_pos, to_neg;
            if (tare_obs[f]) {
                const bool differs = obs && v != tare_val[f];
                to_pos = differs;
                to_neg = !obs || differs;
            } else {
                to_pos = obs;
                to_neg = false;
            }
            if (to_pos) {
                if (FILL) { pos_feature[np] = (uint32_t)f; pos_value[np] = v; }
                ++np;
            }
            if (to_neg) {
                if (FILL) neg_feature[nn] = (uint32_t)f;
                ++nn;
            }
        }
        if (!FILL) { pos_ptr[r + 1] = np; neg_ptr[r + 1] = nn; }
    }
}

// The first walk (counts only), eight columns at a time: the thread-per-row walk above issues one mask word and one
// cell per column behind a branch, so a warp has one load in flight; here the eight mask words are requested together,
// then the eight cells (predicated, no branch between them), then the entries are counted.  Same predicate as
// k_differ_rows / k_differ_fill, so the offsets the scan makes of these counts are the ones the fill pass consumes.
#define DC_U 8
__global__ void __launch_bounds__(128)
k_differ_count(const FeatDev *feats, int F, RowsDev rows, const uint8_t *tare_obs, const uint32_t *tare_val,
               unsigned long long *pos_ptr, unsigned long long *neg_ptr) {
    for (uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; r < rows.R; r += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t np = 0, nn = 0;
        const uint64_t word_at = r >> 5;
        const uint32_t bit = (uint32_t)(r & 31);
        for (int f0 = 0; f0 < F; f0 += DC_U) {
            bool obs[DC_U];
            uint32_t v[DC_U];
#pragma unroll
            for (int u = 0; u < DC_U; ++u)
                obs[u] = f0 + u < F && ((rows.mask[(uint64_t)(f0 + u) * rows.W + word_at] >> bit) & 1u);
            int col[DC_U];
            bool is_wide[DC_U];
#pragma unroll
            for (int u = 0; u < DC_U; ++u) {                       // descriptors without a branch (cached, the same for every row)
                const int f = min(f0 + u, F - 1);
                col[u] = feats[f].col;
                is_wide[u] = feats[f].is_wide != 0;
            }
#pragma unroll
            for (int u = 0; u < DC_U; ++u) {                       // one predicated load each: all eight go out together
                uint32_t w = 0u, b = 0u;
                if (obs[u] && is_wide[u]) w = rows.wide[(uint64_t)col[u] * rows.R + r];
                if (obs[u] && !is_wide[u]) b = rows.cat8[(uint64_t)col[u] * rows.R + r];
                v[u] = w | b;
            }
#pragma unroll
            for (int u = 0; u < DC_U; ++u) {
                if (f0 + u >= F) break;
                if (tare_obs[f0 + u]) {
                    const bool differs = obs[u] && v[u] != tare_val[f0 + u];
                    np += differs;
                    nn += !obs[u] || differs;
                } else {
                    np += obs[u];
                }
            }
        }
        pos_ptr[r + 1] = np;
        neg_ptr[r + 1] = nn;
    }
}

// The second walk with coalesced stores.  k_differ_rows<true> keeps a thread per row, so the 32 rows of a warp write 32
// different places of the CSR arrays 4 bytes at a time: sectors leave the L2 half-filled and come back to be completed
// (ncu, 10^6 rows x 100 columns: 576 MB read and 770 MB written for 322 MB of output; profiles/r2_differ_dram.md).
// Here a warp takes 32 consecutive rows and 32 columns at a time: the cells are loaded with lane = row (coalesced along
// the feature-major block, as before), transposed through shared memory, and emitted row by row with lane = column -- a
// ballot and a popcount place every entry, so a row's entries are stored by consecutive lanes at consecutive addresses.
// Entry order inside a row is ascending column, as _abs_to_rel_type produces it (src/differ.cc:248-298).
#define DF_WARPS 4
__global__ void __launch_bounds__(32 * DF_WARPS)
k_differ_fill(const FeatDev *feats, int F, RowsDev rows, const uint8_t *tare_obs, const uint32_t *tare_val,
              const unsigned long long *pos_ptr, const unsigned long long *neg_ptr, uint32_t *pos_feature,
              uint32_t *pos_value, uint32_t *neg_feature) {
    __shared__ uint32_t s_val[DF_WARPS][32][33];    // [column of the chunk][row of the tile], padded: both walks conflict-free
    __shared__ uint32_t s_obs[DF_WARPS][32];        // per column: bit i = row i of the tile observes it
    const unsigned FULL = 0xFFFFFFFFu;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t n_tiles = (rows.R + 31) >> 5;
    for (uint64_t tile = (uint64_t)blockIdx.x * DF_WARPS + warp; tile < n_tiles; tile += (uint64_t)gridDim.x * DF_WARPS) {
        const uint64_t r0 = tile << 5, r = r0 + lane;
        const int n_rows = (int)min((uint64_t)32, rows.R - r0);
        const uint32_t live = n_rows == 32 ? FULL : ((1u << n_rows) - 1u);
        unsigned long long np = lane < n_rows ? pos_ptr[r] : 0ull, nn = lane < n_rows ? neg_ptr[r] : 0ull;   // lane = row
        for (int f0 = 0; f0 < F; f0 += 32) {
            const int nf = min(32, F - f0);
            for (int j0 = 0; j0 < nf; j0 += 8) {                             // lane = row; 8 columns' loads in flight at once
                uint32_t word[8], v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u)                                  // the tile's rows are one mask word per column
                    word[u] = j0 + u < nf ? rows.mask[(uint64_t)(f0 + j0 + u) * rows.W + tile] & live : 0u;
#pragma unroll
                for (int u = 0; u < 8; ++u) {                                // descriptor without a branch, then one predicated load
                    const int f = min(f0 + j0 + u, F - 1);
                    const int col = feats[f].col;
                    const bool is_wide = feats[f].is_wide != 0, on = (word[u] >> lane) & 1u;
                    uint32_t w = 0u, b = 0u;
                    if (on && is_wide) w = rows.wide[(uint64_t)col * rows.R + r];
                    if (on && !is_wide) b = rows.cat8[(uint64_t)col * rows.R + r];
                    v[u] = w | b;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    if (j0 + u < nf) {
                        s_val[warp][j0 + u][lane] = v[u];
                        if (lane == 0) s_obs[warp][j0 + u] = word[u];
                    }
            }
            __syncwarp();
            const int f = f0 + lane;                                         // lane = column
            const bool have = lane < nf;
            const bool t_obs = have && tare_obs[f];
            const uint32_t t_val = have ? tare_val[f] : 0u;
            const uint32_t obs_word = have ? s_obs[warp][lane] : 0u;
            const uint32_t below = (1u << lane) - 1u;
            for (int i = 0; i < n_rows; ++i) {
                const bool obs = (obs_word >> i) & 1u;
                const uint32_t v = s_val[warp][lane][i];
                bool to_pos, to_neg;
                if (t_obs) {
                    const bool differs = obs && v != t_val;
                    to_pos = differs;
                    to_neg = !obs || differs;
                } else {
                    to_pos = obs;                                            // obs is false for lanes past the last column
                    to_neg = false;
                }
                const uint32_t mp = __ballot_sync(FULL, to_pos), mn = __ballot_sync(FULL, to_neg);
                const unsigned long long bp = __shfl_sync(FULL, np, i), bn = __shfl_sync(FULL, nn, i);
                if (to_pos) {
                    const unsigned long long at = bp + (unsigned long long)__popc(mp & below);
                    pos_feature[at] = (uint32_t)f;
                    pos_value[at] = v;
                }
                if (to_neg) neg_feature[bn + (unsigned long long)__popc(mn & below)] = (uint32_t)f;
                if (lane == i) { np += (unsigned long long)__popc(mp); nn += (unsigned long long)__popc(mn); }
            }
            __syncwarp();
        }
    }
}

static DifferState *state(Engine *e) {
    if (!e->differ) e->differ = new DifferState;
    return (DifferState *)e->differ;
}

extern "C" int lb_differ_reset(lb_handle h) {
    Engine *e = (Engine *)h;
    LB_CUDA(cudaSetDevice(e->device));
    if (!e->F) return lb_fail("differ_reset before set_model");
    DifferState *s = state(e);
    if (s->F != e->F) {
        cudaFree(s->d_hist);
        s->d_hist = nullptr;
        LB_CUDA(cudaMalloc(&s->d_hist, 8 * (size_t)e->F * DIFFER_BINS));
        s->F = e->F;
    }
    LB_CUDA(cudaMemsetAsync(s->d_hist, 0, 8 * (size_t)e->F * DIFFER_BINS, e->stream));
    s->rows_added = 0;
    return 0;
}

// per feature: the offset of its dictionary, -1 = identity
static std::vector<int32_t> dictionary_offsets(const Engine *e, const int32_t *dpd_offsets) {
    std::vector<int32_t> off(e->F, -1);
    int ord = 0;
    for (int f = 0; f < e->F; ++f)
        if (e->specs[f].type == LB_DPD) {
            if (dpd_offsets) off[f] = dpd_offsets[ord];
            ++ord;
        }
    return off;
}

extern "C" int lb_differ_add_rows(lb_handle h, const int32_t *dpd_offsets, const uint32_t *dpd_values) {
    Engine *e = (Engine *)h;
    LB_CUDA(cudaSetDevice(e->device));
    DifferState *s = (DifferState *)e->differ;
    if (!s || !s->d_hist || s->F != e->F) return lb_fail("differ_add_rows before differ_reset");
    if (!e->R) return 0;
    LB_TRY(lb_sync_model(e));
    const std::vector<int32_t> off = dictionary_offsets(e, dpd_offsets);
    int n_dict = 0;
    if (dpd_offsets)
        for (int f = 0, ord = 0; f < e->F; ++f)
            if (e->specs[f].type == LB_DPD) {
                if (dpd_offsets[ord + 1] - dpd_offsets[ord] < e->specs[f].dim)
                    return lb_fail("differ_add_rows: dictionary of feature %d is shorter than its dim", f);
                n_dict = dpd_offsets[++ord];
            }
    const size_t off_bytes = (4 * (size_t)e->F + 255) / 256 * 256;
    LB_TRY(lb_ensure_scratch(e, off_bytes + 4 * (size_t)std::max(n_dict, 1) + 256));
    int32_t *d_off = (int32_t *)e->d_scratch;
    uint32_t *d_dict = (uint32_t *)((char *)e->d_scratch + off_bytes);
    LB_CUDA(cudaMemcpyAsync(d_off, off.data(), 4 * (size_t)e->F, cudaMemcpyHostToDevice, e->stream));
    if (n_dict) LB_CUDA(cudaMemcpyAsync(d_dict, dpd_values, 4 * (size_t)n_dict, cudaMemcpyHostToDevice, e->stream));
    RowsDev rows{e->R, e->W, e->d_cat8, e->d_wide, e->d_mask};
    // row tiles x columns; enough CTAs per column to fill the SMs whatever F is
    const unsigned tiles = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((e->R + 255) / 256, (uint64_t)std::max(1, e->n_sm * 8 / e->F + 1)));
    k_differ_hist<<<dim3(tiles, (unsigned)e->F), 256, 0, e->stream>>>(e->d_feats, rows, d_off, d_dict, s->d_hist);
    LB_CUDA(cudaGetLastError());
    e->launches[LB_L_TOTAL]++; e->launches[LB_L_MISC]++;
    LB_CUDA(cudaStreamSynchronize(e->stream));       // the staging copies above came from pageable memory
    s->rows_added += e->R;
    return 0;
}

extern "C" int lb_differ_make_tare(lb_handle h, const int32_t *dpd_offsets, const uint32_t *dpd_values,
                                   uint8_t *tare_observed, uint32_t *tare_values) {
    Engine *e = (Engine *)h;
    LB_CUDA(cudaSetDevice(e->device));
    DifferState *s = (DifferState *)e->differ;
    if (!s || !s->d_hist || s->F != e->F) return lb_fail("differ_make_tare before differ_reset");
    std::vector<unsigned long long> hist((size_t)e->F * DIFFER_BINS);
    LB_CUDA(cudaMemcpyAsync(hist.data(), s->d_hist, 8 * hist.size(), cudaMemcpyDeviceToHost, e->stream));
    LB_CUDA(cudaStreamSynchronize(e->stream));
    // _make_tare_type (src/differ.cc:191-206): F comparisons, on the host
    const float count_threshold = (float)(0.5 * (double)s->rows_added);
    int ord = 0;
    for (int f = 0; f < e->F; ++f) {
        const bool is_dpd = e->specs[f].type == LB_DPD;
        const unsigned long long *counts = hist.data() + (size_t)f * DIFFER_BINS;
        uint32_t mode = 0;
        for (uint32_t i = 0; i < DIFFER_BINS; ++i)
            if (counts[i] > counts[mode]) mode = i;
        tare_observed[f] = e->specs[f].type != LB_NICH && (float)counts[mode] > count_threshold;
        tare_values[f] = 0;
        if (tare_observed[f]) {
            uint32_t cell = mode;
            if (is_dpd && dpd_offsets) {                   // back to the dictionary index the row block stores
                cell = LB_UNASSIGNED;
                for (int32_t i = dpd_offsets[ord]; i < dpd_offsets[ord + 1]; ++i)
                    if (dpd_values[i] == mode) cell = (uint32_t)(i - dpd_offsets[ord]);
                if (cell == LB_UNASSIGNED) return lb_fail("differ_make_tare: mode of feature %d is not in its dictionary", f);
            }
            tare_values[f] = cell;
        }
        ord += is_dpd;
    }
    return 0;
}

extern "C" int lb_differ_compress_rows(lb_handle h, const uint8_t *tare_observed, const uint32_t *tare_values,
                                       uint64_t *n_pos, uint64_t *n_neg) {
    Engine *e = (Engine *)h;
    LB_CUDA(cudaSetDevice(e->device));
    if (!e->F) return lb_fail("differ_compress_rows before set_model");
    DifferState *s = state(e);
    s->have_diff = false;
    LB_TRY(lb_sync_model(e));
    const uint64_t R = e->R;
    if (R + 1 > s->cap_rows) {
        cudaFree(s->d_pos_ptr); cudaFree(s->d_neg_ptr);
        s->d_pos_ptr = s->d_neg_ptr = nullptr;
        s->cap_rows = 0;
        LB_CUDA(cudaMalloc(&s->d_pos_ptr, 8 * (size_t)(R + 1)));
        LB_CUDA(cudaMalloc(&s->d_neg_ptr, 8 * (size_t)(R + 1)));
        s->cap_rows = R + 1;
    }
    LB_CUDA(cudaMemsetAsync(s->d_pos_ptr, 0, 8, e->stream));
    LB_CUDA(cudaMemsetAsync(s->d_neg_ptr, 0, 8, e->stream));
    size_t scan_bytes = 0;
    LB_CUDA(cub::DeviceScan::InclusiveSum(nullptr, scan_bytes, s->d_pos_ptr, s->d_pos_ptr, (int)std::max<uint64_t>(R, 1), e->stream));
    const size_t obs_bytes = ((size_t)e->F + 255) / 256 * 256, val_bytes = (4 * (size_t)e->F + 255) / 256 * 256;
    LB_TRY(lb_ensure_scratch(e, obs_bytes + val_bytes + scan_bytes + 256));
    uint8_t *d_obs = (uint8_t *)e->d_scratch;
    uint32_t *d_val = (uint32_t *)((char *)e->d_scratch + obs_bytes);
    void *d_scan = (char *)e->d_scratch + obs_bytes + val_bytes;
    LB_CUDA(cudaMemcpyAsync(d_obs, tare_observed, (size_t)e->F, cudaMemcpyHostToDevice, e->stream));
    LB_CUDA(cudaMemcpyAsync(d_val, tare_values, 4 * (size_t)e->F, cudaMemcpyHostToDevice, e->stream));
    RowsDev rows{R, e->W, e->d_cat8, e->d_wide, e->d_mask};
    const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((R + 127) / 128, (uint64_t)e->n_sm * 16));
    unsigned long long total[2] = {0, 0};
    if (R) {
        static const bool plain_count = getenv("LB_DIFFER_FILL") && !strcmp(getenv("LB_DIFFER_FILL"), "rows");   // A/B switch
        if (plain_count)
            k_differ_rows<false><<<grid, 128, 0, e->stream>>>(e->d_feats, e->F, rows, d_obs, d_val, s->d_pos_ptr, s->d_neg_ptr,
                                                             nullptr, nullptr, nullptr);
        else
            k_differ_count<<<grid, 128, 0, e->stream>>>(e->d_feats, e->F, rows, d_obs, d_val, s->d_pos_ptr, s->d_neg_ptr);
        LB_CUDA(cudaGetLastError());
        LB_CUDA(cub::DeviceScan::InclusiveSum(d_scan, scan_bytes, s->d_pos_ptr + 1, s->d_pos_ptr + 1, (int)R, e->stream));
        LB_CUDA(cub::DeviceScan::InclusiveSum(d_scan, scan_bytes, s->d_neg_ptr + 1, s->d_neg_ptr + 1, (int)R, e->stream));
        LB_CUDA(cudaMemcpyAsync(&total[0], s->d_pos_ptr + R, 8, cudaMemcpyDeviceToHost, e->stream));
        LB_CUDA(cudaMemcpyAsync(&total[1], s->d_neg_ptr + R, 8, cudaMemcpyDeviceToHost, e->stream));
        LB_CUDA(cudaStreamSynchronize(e->stream));
    }
    if (total[0] > s->cap_pos || !s->d_pos_feature) {
        cudaFree(s->d_pos_feature); cudaFree(s->d_pos_value);
        s->d_pos_feature = s->d_pos_value = nullptr;
        s->cap_pos = 0;
        LB_CUDA(cudaMalloc(&s->d_pos_feature, 4 * (size_t)std::max<unsigned long long>(total[0], 1)));
        LB_CUDA(cudaMalloc(&s->d_pos_value, 4 * (size_t)std::max<unsigned long long>(total[0], 1)));
        s->cap_pos = total[0];
    }
    if (total[1] > s->cap_neg || !s->d_neg_feature) {
        cudaFree(s->d_neg_feature);
        s->d_neg_feature = nullptr;
        s->cap_neg = 0;
        LB_CUDA(cudaMalloc(&s->d_neg_feature, 4 * (size_t)std::max<unsigned long long>(total[1], 1)));
        s->cap_neg = total[1];
    }
    if (R) {
        static const bool thread_per_row = getenv("LB_DIFFER_FILL") && !strcmp(getenv("LB_DIFFER_FILL"), "rows");   // A/B switch
        if (thread_per_row) {
            k_differ_rows<true><<<grid, 128, 0, e->stream>>>(e->d_feats, e->F, rows, d_obs, d_val, s->d_pos_ptr, s->d_neg_ptr,
                                                            s->d_pos_feature, s->d_pos_value, s->d_neg_feature);
        } else {
            const uint64_t tiles = (R + 31) / 32;
            const int fill_grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((tiles + DF_WARPS - 1) / DF_WARPS, (uint64_t)e->n_sm * 16));
            k_differ_fill<<<fill_grid, 32 * DF_WARPS, 0, e->stream>>>(e->d_feats, e->F, rows, d_obs, d_val, s->d_pos_ptr, s->d_neg_ptr,
                                                                     s->d_pos_feature, s->d_pos_value, s->d_neg_feature);
        }
        LB_CUDA(cudaGetLastError());
        e->launches[LB_L_TOTAL] += 4; e->launches[LB_L_MISC] += 4;
        LB_CUDA(cudaStreamSynchronize(e->stream));
    }
    s->rows = R;
    s->n_pos = total[0];
    s->n_neg = total[1];
    s->have_diff = true;
    *n_pos = s->n_pos;
    *n_neg = s->n_neg;
    return 0;
}

extern "C" int lb_differ_fetch(lb_handle h, uint64_t *pos_ptr, uint32_t *pos_feature, uint32_t *pos_value,
                               uint64_t *neg_ptr, uint32_t *neg_feature) {
    Engine *e = (Engine *)h;
    LB_CUDA(cudaSetDevice(e->device));
    DifferState *s = (DifferState *)e->differ;
    if (!s || !s->have_diff || s->rows != e->R) return lb_fail("differ_fetch before differ_compress_rows");
    LB_CUDA(cudaMemcpyAsync(pos_ptr, s->d_pos_ptr, 8 * (size_t)(s->rows + 1), cudaMemcpyDeviceToHost, e->stream));
    LB_CUDA(cudaMemcpyAsync(neg_ptr, s->d_neg_ptr, 8 * (size_t)(s->rows + 1), cudaMemcpyDeviceToHost, e->stream));
    if (s->n_pos) {
        LB_CUDA(cudaMemcpyAsync(pos_feature, s->d_pos_feature, 4 * (size_t)s->n_pos, cudaMemcpyDeviceToHost, e->stream));
        LB_CUDA(cudaMemcpyAsync(pos_value, s->d_pos_value, 4 * (size_t)s->n_pos, cudaMemcpyDeviceToHost, e->stream));
    }
    if (s->n_neg) LB_CUDA(cudaMemcpyAsync(neg_feature, s->d_neg_feature, 4 * (size_t)s->n_neg, cudaMemcpyDeviceToHost, e->stream));
    LB_CUDA(cudaStreamSynchronize(e->stream));
    return 0;
}
