// lb_accum.cuh -- bulk sufficient-statistic accumulation shared by K2
// (ProductMixture::add_value for given assignments, src/product_mixture.cc:165-184)
// and K4 (KindKernel::add_to_kind_proposer, src/kind_kernel.hpp:193-217).
//
// One CTA owns one (feature, row chunk).  The feature's per-group table
// ([values][G] uint32 counts, or count/sum + float64 moment sums) is privatised
// in shared memory: a thread takes one row (coalesced column / assignment /
// mask reads) and does ONE shared-memory atomic; the table is flushed to the
// global table once per CTA, skipping zeros, and DD's count_sum row is derived
// from the counts during the flush.  Tables that do not fit the shared-memory
// budget fall back to global atomics.
#pragma once
#include "lb_common.cuh"
#include "lb_internal.h"

constexpr int ACC_THREADS = 256;
constexpr int ACC_UNROLL = 4;     // rows per thread per trip, scalar path
constexpr int ACC_VUNROLL = 2;    // row quads per thread per trip, vector path

__host__ __device__ inline int lb_acc_int_rows(int type, int dim) {
    return type == LB_BB ? 2 : ((type == LB_DD || type == LB_DPD) ? dim + 1 : (type == LB_GP ? 2 : 1));
}
__host__ __device__ inline int lb_acc_flt_rows(int type) {
    return type == LB_GP ? 1 : (type == LB_NICH ? 2 : 0);
}
__host__ __device__ inline size_t lb_acc_smem_bytes(int type, int dim, int G) {
    return (size_t)G * (8 * (size_t)lb_acc_flt_rows(type) + 4 * (size_t)lb_acc_int_rows(type, dim)) + 16;
}

// AccTarget (lb_internal.h): the packed group id of every row (1, 2 or 4 bytes each, all-ones =
// unassigned) and the [rows][st] tables the statistics go to.  Narrow ids matter: the id column
// is re-read once per (feature, kind) and is the largest stream of the proposer accumulate.
__device__ __forceinline__ uint32_t lb_load_packed(const void *p, int width, uint64_t r) {
    if (width == 1) { const uint32_t v = ((const uint8_t *)p)[r]; return v == 0xFFu ? LB_UNASSIGNED : v; }
    if (width == 2) { const uint32_t v = ((const uint16_t *)p)[r]; return v == 0xFFFFu ? LB_UNASSIGNED : v; }
    return ((const uint32_t *)p)[r];
}
// ids of rows r .. r+3 (r % 4 == 0, buffer 16-byte aligned) in one load
__device__ __forceinline__ void lb_load_packed4(const void *p, int width, uint64_t r, uint32_t id[4]) {
    if (width == 1) {
        const uint32_t w = *(const uint32_t *)((const uint8_t *)p + r);
#pragma unroll
        for (int i = 0; i < 4; ++i) { const uint32_t v = (w >> (8 * i)) & 0xFFu; id[i] = v == 0xFFu ? LB_UNASSIGNED : v; }
    } else if (width == 2) {
        const uint2 w = *(const uint2 *)((const uint16_t *)p + r);
        const uint32_t v[4] = {w.x & 0xFFFFu, w.x >> 16, w.y & 0xFFFFu, w.y >> 16};
#pragma unroll
        for (int i = 0; i < 4; ++i) id[i] = v[i] == 0xFFFFu ? LB_UNASSIGNED : v[i];
    } else {
        const uint4 w = *(const uint4 *)((const uint32_t *)p + r);
        id[0] = w.x; id[1] = w.y; id[2] = w.z; id[3] = w.w;
    }
}

// one cell into the shared-memory table; CLS: LB_BB, LB_DD (also DPD), LB_GP, LB_NICH
template <int CLS>
__device__ __forceinline__ void lb_acc_cell(uint32_t *si, double *sf, int G, uint32_t g, uint32_t raw) {
    if (CLS == LB_BB) atomicAdd(&si[(raw ? 0 : G) + g], 1u);
    else if (CLS == LB_DD) atomicAdd(&si[raw * (uint32_t)G + g], 1u);
    else if (CLS == LB_GP) {
        atomicAdd(&si[g], 1u);
        atomicAdd(&si[G + g], raw);
        atomicAdd(&sf[g], lb_log_factorial_d(raw));
    } else {
        const double x = (double)__uint_as_float(raw);
        atomicAdd(&si[g], 1u);
        atomicAdd(&sf[g], x);
        atomicAdd(&sf[G + g], x * x);
    }
}

// Rows [r_lo, r_hi) of one column into the shared-memory table.  VEC: a thread takes four
// consecutive rows per load (one 32-bit word of 8-bit values / ids, one mask word) -- ~8
// instructions per row instead of ~50 -- and needs r_lo % 4 == 0 and R % 4 == 0.  Either way a
// trip issues all of its loads before using any: the id, mask and value of a row are independent
// (unobserved cells are loaded and ignored), so nothing serialises on an L2 round trip.
template <int CLS, bool VEC>
__device__ __forceinline__ void lb_acc_rows(uint32_t *si, double *sf, int G, const void *packed, int width,
                                            const uint32_t *mask, const uint8_t *col8, const uint32_t *col32,
                                            uint64_t r_lo, uint64_t r_hi) {
    if (VEC) {
        const uint64_t nq = (r_hi - r_lo) >> 2;
        for (uint64_t q0 = 0; q0 < nq; q0 += (uint64_t)ACC_VUNROLL * ACC_THREADS) {
            uint32_t ids[ACC_VUNROLL][4], raws[ACC_VUNROLL][4], obs[ACC_VUNROLL];
#pragma unroll
            for (int j = 0; j < ACC_VUNROLL; ++j) {
                const uint64_t q = q0 + (uint64_t)j * ACC_THREADS + threadIdx.x;
                obs[j] = 0u;
                if (q < nq) {
                    const uint64_t r = r_lo + 4 * q;
                    lb_load_packed4(packed, width, r, ids[j]);
                    obs[j] = (mask[r >> 5] >> (r & 31)) & 0xFu;
                    if (col32) {
                        const uint4 w = *(const uint4 *)(col32 + r);
                        raws[j][0] = w.x; raws[j][1] = w.y; raws[j][2] = w.z; raws[j][3] = w.w;
                    } else {
                        const uint32_t w = *(const uint32_t *)(col8 + r);
#pragma unroll
                        for (int i = 0; i < 4; ++i) raws[j][i] = (w >> (8 * i)) & 0xFFu;
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < ACC_VUNROLL; ++j)
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (((obs[j] >> i) & 1u) && ids[j][i] != LB_UNASSIGNED) lb_acc_cell<CLS>(si, sf, G, ids[j][i], raws[j][i]);
        }
        r_lo += 4 * nq;     // at most three rows left
    }
    for (uint64_t base = r_lo; base < r_hi; base += (uint64_t)ACC_UNROLL * ACC_THREADS) {
        uint32_t pgs[ACC_UNROLL], mws[ACC_UNROLL], raws[ACC_UNROLL];
#pragma unroll
        for (int j = 0; j < ACC_UNROLL; ++j) {
            const uint64_t r = base + (uint64_t)j * ACC_THREADS + threadIdx.x;
            const bool in = r < r_hi;
            pgs[j] = in ? lb_load_packed(packed, width, r) : LB_UNASSIGNED;
            mws[j] = in ? mask[r >> 5] : 0u;
            raws[j] = in ? (col32 ? col32[r] : (uint32_t)col8[r]) : 0u;
        }
#pragma unroll
        for (int j = 0; j < ACC_UNROLL; ++j) {
            const uint64_t r = base + (uint64_t)j * ACC_THREADS + threadIdx.x;
            if (pgs[j] != LB_UNASSIGNED && ((mws[j] >> (r & 31)) & 1u)) lb_acc_cell<CLS>(si, sf, G, pgs[j], raws[j]);
        }
    }
}

// feats: FeatDev table whose int_row / flt_row address the TARGET tables; fid_list maps
// blockIdx.y to a global feature id (nullptr: identity); blockIdx.z picks the target
// (many[blockIdx.z], or `one` when many == nullptr).
// Moment rows of the target hold SUMS while accumulating (sum x, sum x^2 / sum log x!).
__global__ void __launch_bounds__(ACC_THREADS)
k_accum_feature(const FeatDev *feats, const int32_t *fid_list, RowsDev rows, AccTarget one, const AccTarget *many,
                uint64_t b, uint64_t en, uint64_t rows_per_cta, int sign, int use_smem) {
    const AccTarget tg = many ? many[blockIdx.z] : one;
    const void *packed = tg.packed;
    const int st = tg.st, G = tg.G, width = tg.width;
    extern __shared__ __align__(16) unsigned char acc_smem[];
    const int fid = fid_list ? fid_list[blockIdx.y] : (int)blockIdx.y;
    const FeatDev f = feats[fid];
    const int nir = lb_acc_int_rows(f.type, f.dim), nfr = lb_acc_flt_rows(f.type);
    const uint64_t r_lo = b + blockIdx.x * rows_per_cta;
    const uint64_t r_hi = r_lo + rows_per_cta < en ? r_lo + rows_per_cta : en;
    if (r_lo >= r_hi) return;
    const uint32_t *mask = rows.mask + (uint64_t)fid * rows.W;
    uint32_t *gi = tg.ints + (size_t)f.int_row * st;
    double *gf = tg.flts + (size_t)f.flt_row * st;
    const uint8_t *col8 = f.is_wide ? nullptr : rows.cat8 + (uint64_t)f.col * rows.R;
    const uint32_t *col32 = f.is_wide ? rows.wide + (uint64_t)f.col * rows.R : nullptr;
    const bool smem_ok = use_smem && lb_acc_smem_bytes(f.type, f.dim, G) <= (size_t)use_smem;
    if (!smem_ok) {     // table larger than shared memory: global atomics, one row per thread
        for (uint64_t r = r_lo + threadIdx.x; r < r_hi; r += ACC_THREADS) {
            const uint32_t g = lb_load_packed(packed, width, r);
            if (g == LB_UNASSIGNED || !((mask[r >> 5] >> (r & 31)) & 1u)) continue;
            const uint32_t raw = col32 ? col32[r] : (uint32_t)col8[r];
            uint32_t *in = gi + g;
            double *fl = gf + g;
            switch (f.type) {
            case LB_BB: atomicAdd(&in[raw ? 0 : st], (uint32_t)sign); break;
            case LB_DD:
            case LB_DPD:
                atomicAdd(&in[(size_t)raw * st], (uint32_t)sign);
                atomicAdd(&in[(size_t)f.dim * st], (uint32_t)sign);
                break;
            case LB_GP:
                atomicAdd(&in[0], (uint32_t)sign);
                atomicAdd(&in[st], (uint32_t)(sign * (int64_t)raw));
                atomicAdd(&fl[0], sign * lb_log_factorial_d(raw));
                break;
            case LB_NICH: {
                const double x = (double)__uint_as_float(raw);
                atomicAdd(&in[0], (uint32_t)sign);
                atomicAdd(&fl[0], sign * x);
                atomicAdd(&fl[st], sign * x * x);
                break;
            }
            }
        }
        return;
    }
    double *sf = (double *)acc_smem;                       // [nfr][G]
    uint32_t *si = (uint32_t *)(sf + (size_t)nfr * G);     // [nir][G]
    for (int i = threadIdx.x; i < nfr * G; i += ACC_THREADS) sf[i] = 0.0;
    for (int i = threadIdx.x; i < nir * G; i += ACC_THREADS) si[i] = 0u;
    __syncthreads();
    const bool vec = !(rows.R & 3) && !(r_lo & 3);
    const bool cat = f.type == LB_DD || f.type == LB_DPD;
#define LB_ACC_GO(CLS) \
    (vec ? lb_acc_rows<CLS, true>(si, sf, G, packed, width, mask, col8, col32, r_lo, r_hi) \
         : lb_acc_rows<CLS, false>(si, sf, G, packed, width, mask, col8, col32, r_lo, r_hi))
    if (cat) LB_ACC_GO(LB_DD);
    else if (f.type == LB_BB) LB_ACC_GO(LB_BB);
    else if (f.type == LB_GP) LB_ACC_GO(LB_GP);
    else LB_ACC_GO(LB_NICH);
#undef LB_ACC_GO
    __syncthreads();
    const int value_rows = cat ? f.dim : nir;
    for (int i = threadIdx.x; i < value_rows * G; i += ACC_THREADS) {
        const uint32_t v = si[i];
        if (v) atomicAdd(&gi[(size_t)(i / G) * st + (i % G)], sign > 0 ? v : 0u - v);
    }
    if (cat) {   // count_sum = sum over values
        for (int g = threadIdx.x; g < G; g += ACC_THREADS) {
            uint32_t tot = 0;
            for (int v = 0; v < f.dim; ++v) tot += si[(size_t)v * G + g];
            if (tot) atomicAdd(&gi[(size_t)f.dim * st + g], sign > 0 ? tot : 0u - tot);
        }
    }
    for (int i = threadIdx.x; i < nfr * G; i += ACC_THREADS) {
        const double v = sf[i];
        if (v != 0.0) atomicAdd(&gf[(size_t)(i / G) * st + (i % G)], sign > 0 ? v : -v);
    }
}

// packed ids: out[r] = g2p[assign[r]] narrowed to `width` bytes; blockIdx.y picks the kind
__global__ void k_pack_assign(PackSrc one, const PackSrc *many, uint64_t b, uint64_t en) {
    const PackSrc s = many ? many[blockIdx.y] : one;
    for (uint64_t r = b + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; r < en; r += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t a = s.assign[r];
        const uint32_t g = a == LB_UNASSIGNED ? LB_UNASSIGNED : s.g2p[a];
        if (s.width == 1) ((uint8_t *)s.out)[r] = (uint8_t)g;
        else if (s.width == 2) ((uint16_t *)s.out)[r] = (uint16_t)g;
        else ((uint32_t *)s.out)[r] = g;
    }
}

// clustering counts of the rows [b, en): histogram of the packed ids
__global__ void __launch_bounds__(ACC_THREADS)
k_accum_counts(const void *packed, int width, uint64_t b, uint64_t en, int sign, uint32_t *counts,
               unsigned long long *sample_size, int G, int32_t *flag) {
    extern __shared__ __align__(16) unsigned char acc_smem[];
    uint32_t *sc = (uint32_t *)acc_smem;
    for (int g = threadIdx.x; g < G; g += ACC_THREADS) sc[g] = 0u;
    __syncthreads();
    for (uint64_t r = b + blockIdx.x * (uint64_t)ACC_THREADS + threadIdx.x; r < en; r += (uint64_t)gridDim.x * ACC_THREADS) {
        const uint32_t g = lb_load_packed(packed, width, r);
        if (g != LB_UNASSIGNED) atomicAdd(&sc[g], 1u);
    }
    __syncthreads();
    unsigned long long local = 0;
    for (int g = threadIdx.x; g < G; g += ACC_THREADS) {
        const uint32_t v = sc[g];
        if (!v) continue;
        local += v;
        const uint32_t old = atomicAdd(&counts[g], sign > 0 ? v : 0u - v);
        if (sign < 0 && old < v) *flag = 1;
    }
    if (local) atomicAdd(sample_size, sign > 0 ? local : 0ull - local);
}
