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

constexpr int ACC_THREADS = 256;
constexpr int ACC_UNROLL = 4;

__host__ __device__ inline int lb_acc_int_rows(int type, int dim) {
    return type == LB_BB ? 2 : ((type == LB_DD || type == LB_DPD) ? dim + 1 : (type == LB_GP ? 2 : 1));
}
__host__ __device__ inline int lb_acc_flt_rows(int type) {
    return type == LB_GP ? 1 : (type == LB_NICH ? 2 : 0);
}
__host__ __device__ inline size_t lb_acc_smem_bytes(int type, int dim, int G) {
    return (size_t)G * (8 * (size_t)lb_acc_flt_rows(type) + 4 * (size_t)lb_acc_int_rows(type, dim)) + 16;
}

// feats: FeatDev table whose int_row / flt_row address the TARGET tables; fid_list maps
// blockIdx.y to a global feature id (nullptr: identity).  packed: packed group id per row.
// Moment rows of the target hold SUMS while accumulating (sum x, sum x^2 / sum log x!).
// PT: storage type of the packed group ids (uint8_t when G < 255, uint16_t when G < 65535, else
// uint32_t; the all-ones value means "unassigned").  Narrow ids matter: the id column is re-read
// once per (feature, kind) and is the largest stream of the proposer accumulate.
template <typename PT>
__global__ void __launch_bounds__(ACC_THREADS)
k_accum_feature(const FeatDev *feats, const int32_t *fid_list, RowsDev rows, const PT *packed,
                uint64_t b, uint64_t en, uint64_t rows_per_cta, int sign, uint32_t *ints, double *flts,
                int st, int G, int use_smem) {
    extern __shared__ __align__(16) unsigned char acc_smem[];
    const int fid = fid_list ? fid_list[blockIdx.y] : (int)blockIdx.y;
    const FeatDev f = feats[fid];
    const int nir = lb_acc_int_rows(f.type, f.dim), nfr = lb_acc_flt_rows(f.type);
    const uint64_t r_lo = b + blockIdx.x * rows_per_cta;
    const uint64_t r_hi = r_lo + rows_per_cta < en ? r_lo + rows_per_cta : en;
    if (r_lo >= r_hi) return;
    const uint32_t *mask = rows.mask + (uint64_t)fid * rows.W;
    uint32_t *gi = ints + (size_t)f.int_row * st;
    double *gf = flts + (size_t)f.flt_row * st;
    const bool smem_ok = use_smem && lb_acc_smem_bytes(f.type, f.dim, G) <= (size_t)use_smem;
    double *sf = (double *)acc_smem;                       // [nfr][G]
    uint32_t *si = (uint32_t *)(sf + (size_t)nfr * G);     // [nir][G]
    if (smem_ok) {
        for (int i = threadIdx.x; i < nfr * G; i += ACC_THREADS) sf[i] = 0.0;
        for (int i = threadIdx.x; i < nir * G; i += ACC_THREADS) si[i] = 0u;
        __syncthreads();
    }
    // ACC_UNROLL rows per thread per trip, all loads issued before any is used: the id, mask and
    // value loads of one row are independent (unobserved cells are loaded and ignored), so a trip
    // keeps 3 * ACC_UNROLL requests in flight instead of three serialised L2 round trips per row.
    const uint8_t *col8 = f.is_wide ? nullptr : rows.cat8 + (uint64_t)f.col * rows.R;
    const uint32_t *col32 = f.is_wide ? rows.wide + (uint64_t)f.col * rows.R : nullptr;
    for (uint64_t base = r_lo; base < r_hi; base += (uint64_t)ACC_UNROLL * ACC_THREADS) {
        PT pgs[ACC_UNROLL];
        uint32_t mws[ACC_UNROLL], raws[ACC_UNROLL];
#pragma unroll
        for (int j = 0; j < ACC_UNROLL; ++j) {
            const uint64_t r = base + (uint64_t)j * ACC_THREADS + threadIdx.x;
            const bool in = r < r_hi;
            pgs[j] = in ? packed[r] : (PT)~(PT)0;
            mws[j] = in ? mask[r >> 5] : 0u;
            raws[j] = in ? (col32 ? col32[r] : (uint32_t)col8[r]) : 0u;
        }
#pragma unroll
        for (int j = 0; j < ACC_UNROLL; ++j) {
            const uint64_t r = base + (uint64_t)j * ACC_THREADS + threadIdx.x;
            if (pgs[j] == (PT)~(PT)0) continue;
            if (!((mws[j] >> (r & 31)) & 1u)) continue;
            const uint32_t g = (uint32_t)pgs[j];
            const uint32_t raw = raws[j];
            if (smem_ok) {
                switch (f.type) {
                case LB_BB: atomicAdd(&si[(raw ? 0 : 1) * G + g], 1u); break;
                case LB_DD:
                case LB_DPD: atomicAdd(&si[(size_t)raw * G + g], 1u); break;
                case LB_GP:
                    atomicAdd(&si[g], 1u);
                    atomicAdd(&si[G + g], raw);
                    atomicAdd(&sf[g], lb_log_factorial_d(raw));
                    break;
                case LB_NICH: {
                    const double x = (double)__uint_as_float(raw);
                    atomicAdd(&si[g], 1u);
                    atomicAdd(&sf[g], x);
                    atomicAdd(&sf[G + g], x * x);
                    break;
                }
                }
            } else {
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
        }
    }
    if (!smem_ok) return;
    __syncthreads();
    const bool cat = f.type == LB_DD || f.type == LB_DPD;
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

// packed ids of one kind: g2p[assign[r]] narrowed to PT
template <typename PT>
__global__ void k_pack_assign_t(const uint32_t *assign, const uint32_t *g2p, PT *packed, uint64_t b, uint64_t en) {
    for (uint64_t r = b + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; r < en; r += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t a = assign[r];
        packed[r] = a == LB_UNASSIGNED ? (PT)~(PT)0 : (PT)g2p[a];
    }
}

// clustering counts of the rows [b, en): histogram of the packed ids
template <typename PT>
__global__ void __launch_bounds__(ACC_THREADS)
k_accum_counts(const PT *packed, uint64_t b, uint64_t en, int sign, uint32_t *counts,
               unsigned long long *sample_size, int G, int32_t *flag) {
    extern __shared__ __align__(16) unsigned char acc_smem[];
    uint32_t *sc = (uint32_t *)acc_smem;
    for (int g = threadIdx.x; g < G; g += ACC_THREADS) sc[g] = 0u;
    __syncthreads();
    for (uint64_t r = b + blockIdx.x * (uint64_t)ACC_THREADS + threadIdx.x; r < en; r += (uint64_t)gridDim.x * ACC_THREADS) {
        const PT g = packed[r];
        if (g != (PT)~(PT)0) atomicAdd(&sc[(uint32_t)g], 1u);
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

// host helper: launch k_accum_feature for a list of features, grouped by shared-memory need
struct AccumLaunch {
    const FeatDev *feats;        // device
    RowsDev rows;
    const uint32_t *packed;      // device
    uint32_t *ints;
    double *flts;
    int st, G;
};
