// lb_score_tc.cu -- K1t: the categorical / Bernoulli block of ProductMixture::score_value
// (src/product_mixture.cc:403-434) as a dense contraction on the 5th-generation tensor cores:
//
//     scores[r][g] = prior[g] + sum_k onehot[r][k] * T[k][g]
//
// k runs over the (feature, value) pairs of the kind -- BB 2, DD / DPD dim per feature: the rows of the score table
// k_score_table builds -- onehot[r][k] = 1 when row r holds that value, T[k][g] is the complete log predictive of the
// value in group g.  fp32 accuracy on bf16 tensor cores by SPLIT PRECISION: T = hi + mid + lo with three bf16 planes
// (8 + 8 + 8 mantissa bits), one-hot entries are exactly 0 / 1, products accumulate in fp32 in TMEM, so the result
// differs from the fp32 gather only in summation order.
//
// One persistent CTA per SM, 10 warps:
//   warps 0-3  thread = row of a 128-row tile.  Stage the row's cells (coalesced column reads), then for every
//              64-wide K chunk write that chunk of the one-hot operand into a ring stage (canonical K-major
//              no-swizzle core-matrix layout) and publish it (fence.proxy.async + mbarrier); later the epilogue:
//              tcgen05.ld the row's accumulators, add the CRP prior, transpose through a swizzled shared tile and
//              store coalesced.
//   warp 4     one thread issues tcgen05.mma (M = 128, N = groups padded to 16, K = 16 per instruction):
//              planes x 4 per chunk; tcgen05.commit frees the ring stage / publishes the accumulator.
//   warps 6-9  cell loaders: the NEXT tile's cells (one 32-bit load = 4 rows of a column, TC_LD_UNROLL columns in
//              flight per lane) become operand column indices in a double-buffered shared array, so the row block
//              streams from HBM under the current tile's MMAs and epilogue.
//   warp 5     one thread streams the table chunks (all planes of a chunk are contiguous: k_score_table_tc writes
//              them in the operand layout) with ONE bulk TMA copy per stage (cp.async.bulk + mbarrier complete_tx);
//              the warp also owns the TMEM allocation.
// Two accumulator stages in TMEM: the MMAs of tile t + 1 run under the epilogue of tile t.
//
// Cost model: M x N x 16 MACs every max(128, M) * N / 256 cycles -> a chunk costs planes * 4 * N / 2
// cycles per 128 rows whatever the number of cells in it, against one gathered line per observed cell for the gather
// kernel.  The contraction wins where (values per feature / density) is small -- Bernoulli and low-cardinality
// columns -- and loses on DD-256, where 255 of 256 operand columns are zero; lb_launch_score_rows picks per kind.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include <cuda_bf16.h>

#include "lb_internal.h"

namespace {

constexpr int TC_ROWS = 128;            // M
constexpr int TC_KC = 64;               // K chunk per ring stage (4 MMAs of K = 16)
constexpr int TC_A_BYTES = TC_ROWS * TC_KC * 2;
constexpr int TC_THREADS = 320;         // 4 producer / epilogue warps, MMA, table loader, 4 cell loaders
constexpr int TC_LD_UNROLL = 8;

struct TcFeat {                         // what the producers need per feature
    int32_t col, is_wide, fid, row_base;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// shared-memory matrix descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor, version 1): the operand is a grid
// of 8 x 16-byte core matrices; lbo = byte distance between the two K halves of one MMA, sbo = between 8-row groups
__device__ __forceinline__ uint64_t tc_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// The table operand: chunk-major, inside a chunk [plane][k / 8][n][k % 8] bf16 -- exactly the shared-memory image of
// the K-major B operand, so one bulk copy fills a ring stage.  plane 0 = bf16(T), 1 = bf16(T - p0), 2 = the rest.
__global__ void k_score_table_tc(const float *tab, int st, int G, int K_total, int n_pad, int planes, int n_chunks,
                                 __nv_bfloat16 *out) {
    const uint64_t per_plane = (uint64_t)TC_KC * n_pad;
    const uint64_t total = (uint64_t)n_chunks * per_plane;
    for (uint64_t idx = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; idx < total; idx += (uint64_t)gridDim.x * blockDim.x) {
        const int c = (int)(idx / per_plane);
        const int rem = (int)(idx % per_plane);
        const int kc = rem / (n_pad * 8), n = (rem / 8) % n_pad, kk = rem % 8;
        const int k = c * TC_KC + kc * 8 + kk;
        float v = (k < K_total && n < G) ? tab[(size_t)(k + 1) * st + n] : 0.f;     // table row 0 is the all-zero row
        for (int p = 0; p < planes; ++p) {
            const __nv_bfloat16 h = __float2bfloat16_rn(v);
            out[((uint64_t)c * planes + p) * per_plane + rem] = h;
            v -= __bfloat162float(h);
        }
    }
}

__global__ void __launch_bounds__(TC_THREADS, 1)
k_score_rows_tc(const KindDev *kinds, int kid, const TcFeat *tfeat, const uint16_t *chunk_lo, const uint16_t *chunk_hi,
                const __nv_bfloat16 *btab, RowsDev rows, uint64_t b, uint64_t en, int n_empty, int K_total, int n_pad,
                int planes, int n_stages, int tmem_cols, int n8, float *out) {
    extern __shared__ __align__(128) unsigned char tc_smem[];
    const uint32_t acc_stride = (uint32_t)(n_pad + 31) & ~31u;     // TMEM columns of one accumulator stage
    const KindDev &kd = kinds[kid];
    const int nf = kd.nf, G = kd.G;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_chunks = (K_total + TC_KC - 1) / TC_KC;
    const uint32_t b_stage_bytes = (uint32_t)planes * n_pad * TC_KC * 2;

    // carve: [A stages][B stages][epilogue transpose 4 x 32 x 32 f32][cells nf x 128 u16][prior n_pad f32][TcFeat nf][barriers]
    unsigned char *sA = tc_smem;
    unsigned char *sB = sA + (size_t)n_stages * TC_A_BYTES;
    float *s_tr = (float *)(sB + (size_t)n_stages * b_stage_bytes);
    uint16_t *s_cell = (uint16_t *)(s_tr + 4 * 32 * 32);
    const int cell_stride = ((nf + 1) & ~1) * TC_ROWS;          // u16 entries of one cell buffer
    float *s_prior = (float *)(s_cell + 2 * (size_t)cell_stride);
    TcFeat *s_feat = (TcFeat *)(s_prior + n_pad);
    uint64_t *bars = (uint64_t *)(((uintptr_t)(s_feat + nf) + 15) & ~(uintptr_t)15);
    uint64_t *fullA = bars, *fullB = bars + n_stages, *empty = bars + 2 * n_stages;
    uint64_t *tmem_full = bars + 3 * n_stages, *tmem_empty = tmem_full + 2;
    uint64_t *cells_full = tmem_empty + 2, *cells_empty = cells_full + 2;
    uint32_t *s_tmem = (uint32_t *)(cells_empty + 2);

    if (tid == 0) {
        for (int s = 0; s < n_stages; ++s) {
            mbar_init(smem_u32(&fullA[s]), 4);
            mbar_init(smem_u32(&fullB[s]), 1);
            mbar_init(smem_u32(&empty[s]), 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(smem_u32(&tmem_full[s]), 1); mbar_init(smem_u32(&tmem_empty[s]), 4);
            mbar_init(smem_u32(&cells_full[s]), 4); mbar_init(smem_u32(&cells_empty[s]), 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"((uint32_t)tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    {
        const float base = -logf((float)kd.sample_size + kd.alpha);
        const float sh_empty = logf((kd.alpha + kd.d * (float)(G - n_empty)) / (float)n_empty);
        for (int g = tid; g < n_pad; g += TC_THREADS)
            s_prior[g] = g < G ? (kd.counts[g] ? kd.shifted[g] : sh_empty) + base : 0.f;
        for (int j = tid; j < nf; j += TC_THREADS) s_feat[j] = tfeat[j];
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;

    const uint64_t n_tiles = (en - b + TC_ROWS - 1) / TC_ROWS;

    if (warp < 4) {
        // ================= one-hot producers + epilogue: thread = row =================
        const int row = tid;
        uint32_t q = 0;                      // ring use counter
        auto epilogue = [&](uint64_t tile, uint32_t it) {
            const uint32_t as = it & 1u;
            mbar_wait(smem_u32(&tmem_full[as]), (it >> 1) & 1u);
            tc_fence_after();
            const uint64_t r0 = b + tile * TC_ROWS;
            float *tr = s_tr + warp * 32 * 32;
            for (int cb = 0; cb < n_pad; cb += 32) {
                uint32_t v[32];
                tc_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + as * acc_stride + (uint32_t)cb, v);
#pragma unroll
                for (int i = 0; i < 32; ++i) tr[lane * 32 + (i ^ lane)] = __uint_as_float(v[i]) + (cb + i < n_pad ? s_prior[cb + i] : 0.f);
                __syncwarp();
                const int g = cb + lane;
                if (g < G) {
                    const uint64_t rw = r0 + warp * 32;                       // first row of this warp
                    const int n_valid = rw >= en ? 0 : (int)min((uint64_t)32, en - rw);
                    float *op = out + (rw - b) * (uint64_t)G + g;
#pragma unroll 8
                    for (int rr = 0; rr < n_valid; ++rr, op += G) *op = tr[rr * 32 + (lane ^ rr)];
                }
                __syncwarp();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&tmem_empty[as]));
        };
        uint32_t it = 0;
        uint64_t prev_tile = 0;
        for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const uint16_t *cell = s_cell + (size_t)(it & 1u) * cell_stride;     // filled by the loader warps
            mbar_wait(smem_u32(&cells_full[it & 1u]), (it >> 1) & 1u);
            for (int c = 0; c < n_chunks; ++c, ++q) {
                const uint32_t s = q % (uint32_t)n_stages, ph = (q / (uint32_t)n_stages) & 1u;
                mbar_wait(smem_u32(&empty[s]), ph ^ 1u);
                unsigned char *a = sA + (size_t)s * TC_A_BYTES + row * 16;
#pragma unroll
                for (int kc = 0; kc < 8; ++kc) *(uint4 *)(a + kc * 2048) = make_uint4(0u, 0u, 0u, 0u);
                const int jlo = chunk_lo[c], jhi = chunk_hi[c];
                for (int j = jlo; j <= jhi; ++j) {
                    const uint32_t k = cell[j * TC_ROWS + row];
                    if (k && (int)((k - 1u) >> 6) == c) {
                        const uint32_t kk = (k - 1u) & 63u;
                        *(uint16_t *)(a + (kk >> 3) * 2048 + (kk & 7u) * 2) = (uint16_t)0x3F80;      // bf16 1.0
                    }
                }
                fence_async_smem();                 // generic-proxy writes -> visible to the tensor core's async proxy
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&fullA[s]));
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&cells_empty[it & 1u]));      // the loaders may overwrite this buffer
            if (it > 0) epilogue(prev_tile, it - 1);
            prev_tile = tile;
        }
        if (it > 0) epilogue(prev_tile, it - 1);
    } else if (warp == 4) {
        // ================= MMA issuer =================
        if (lane == 0) {
            // instruction descriptor (cute::UMMA::InstrDescriptor): D = f32, A = B = bf16, both K-major, N, M = 128
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n_pad >> 3) << 17) | ((uint32_t)(TC_ROWS >> 4) << 24);
            uint32_t q = 0, it = 0;
            for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
                const uint32_t as = it & 1u;
                mbar_wait(smem_u32(&tmem_empty[as]), ((it >> 1) & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * acc_stride;
                for (int c = 0; c < n_chunks; ++c, ++q) {
                    const uint32_t s = q % (uint32_t)n_stages, ph = (q / (uint32_t)n_stages) & 1u;
                    mbar_wait(smem_u32(&fullA[s]), ph);
                    mbar_wait(smem_u32(&fullB[s]), ph);
                    tc_fence_after();
                    const uint32_t a0 = smem_u32(sA + (size_t)s * TC_A_BYTES), b0 = smem_u32(sB + (size_t)s * b_stage_bytes);
                    const int ksteps = min(TC_KC / 16, (K_total - c * TC_KC + 15) / 16);
                    for (int p = 0; p < planes; ++p)
                        for (int ks = 0; ks < ksteps; ++ks) {
                            const uint64_t ad = tc_desc(a0 + (uint32_t)ks * 4096u, 2048u, 128u);
                            const uint64_t bd = tc_desc(b0 + (uint32_t)p * (uint32_t)n_pad * 128u + (uint32_t)ks * (uint32_t)n_pad * 32u,
                                                        (uint32_t)n_pad * 16u, 128u);
                            tc_mma(d_tmem, ad, bd, idesc, (c | p | ks) ? 1u : 0u);
                        }
                    tc_commit(smem_u32(&empty[s]));          // the stage is free once these MMAs have read it
                }
                tc_commit(smem_u32(&tmem_full[as]));         // accumulator complete
            }
        }
    } else if (warp >= 6) {
        // ================= cell loaders: rows of the next tile -> operand column indices =================
        // lane q of loader warp lw takes rows 4q .. 4q+3 of the columns lw, lw + 4, ...: one 32-bit load is 4 rows of
        // an 8-bit column, one 128-bit load 4 rows of a 32-bit column; TC_LD_UNROLL columns are requested at once.
        const int lw = warp - 6;
        const bool aligned = (rows.R & 3ull) == 0;
        uint32_t it = 0;
        for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const uint32_t buf = it & 1u;
            mbar_wait(smem_u32(&cells_empty[buf]), ((it >> 1) & 1u) ^ 1u);
            uint16_t *cell = s_cell + (size_t)buf * cell_stride;
            const uint64_t r4 = b + tile * TC_ROWS + 4 * (uint64_t)lane;          // first of this lane's 4 rows
            // whole tile inside the block and 4-row groups aligned: branch-free vector loads, TC_LD_UNROLL columns in flight
            const bool fast = aligned && ((b & 3ull) == 0) && b + (tile + 1) * TC_ROWS <= en;
            auto store4 = [&](int j, uint32_t bits, uint32_t v0, uint32_t v1, uint32_t v2, uint32_t v3) {
                const uint32_t base = (uint32_t)s_feat[j].row_base;
                const uint32_t k0 = bits & 1u ? base + v0 : 0u, k1 = bits & 2u ? base + v1 : 0u;
                const uint32_t k2 = bits & 4u ? base + v2 : 0u, k3 = bits & 8u ? base + v3 : 0u;
                *(uint2 *)(cell + j * TC_ROWS + 4 * lane) = make_uint2(k0 | (k1 << 16), k2 | (k3 << 16));
            };
            if (fast) {
                const uint64_t mword = r4 >> 5;
                const uint32_t mshift = (uint32_t)(r4 & 31);
                for (int j0 = lw; j0 < n8; j0 += 4 * TC_LD_UNROLL) {          // 8-bit columns: one word = 4 rows
                    uint32_t w[TC_LD_UNROLL], m[TC_LD_UNROLL];
#pragma unroll
                    for (int u = 0; u < TC_LD_UNROLL; ++u) {
                        const TcFeat f = s_feat[min(j0 + 4 * u, n8 - 1)];
                        m[u] = rows.mask[(uint64_t)f.fid * rows.W + mword];
                        w[u] = *(const uint32_t *)(rows.cat8 + (uint64_t)f.col * rows.R + r4);
                    }
#pragma unroll
                    for (int u = 0; u < TC_LD_UNROLL; ++u)
                        if (j0 + 4 * u < n8)
                            store4(j0 + 4 * u, (m[u] >> mshift) & 15u, w[u] & 255u, (w[u] >> 8) & 255u, (w[u] >> 16) & 255u, w[u] >> 24);
                }
                for (int j0 = n8 + lw; j0 < nf; j0 += 4 * (TC_LD_UNROLL / 2)) {   // 32-bit columns: one uint4 = 4 rows
                    uint4 w[TC_LD_UNROLL / 2];
                    uint32_t m[TC_LD_UNROLL / 2];
#pragma unroll
                    for (int u = 0; u < TC_LD_UNROLL / 2; ++u) {
                        const TcFeat f = s_feat[min(j0 + 4 * u, nf - 1)];
                        m[u] = rows.mask[(uint64_t)f.fid * rows.W + mword];
                        w[u] = *(const uint4 *)(rows.wide + (uint64_t)f.col * rows.R + r4);
                    }
#pragma unroll
                    for (int u = 0; u < TC_LD_UNROLL / 2; ++u)
                        if (j0 + 4 * u < nf) store4(j0 + 4 * u, (m[u] >> mshift) & 15u, w[u].x, w[u].y, w[u].z, w[u].w);
                }
            } else {
                for (int j = lw; j < nf; j += 4) {
                    const TcFeat f = s_feat[j];
                    uint32_t bits = 0u, v[4] = {0u, 0u, 0u, 0u};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const uint64_t r = r4 + i;
                        if (r >= en) continue;
                        bits |= ((rows.mask[(uint64_t)f.fid * rows.W + (r >> 5)] >> (r & 31)) & 1u) << i;
                        v[i] = f.is_wide ? rows.wide[(uint64_t)f.col * rows.R + r] : (uint32_t)rows.cat8[(uint64_t)f.col * rows.R + r];
                    }
                    store4(j, bits, v[0], v[1], v[2], v[3]);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&cells_full[buf]));
        }
    } else {
        // ================= table loader: one bulk TMA copy per ring stage =================
        if (lane == 0) {
            uint32_t q = 0;
            for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                for (int c = 0; c < n_chunks; ++c, ++q) {
                    const uint32_t s = q % (uint32_t)n_stages, ph = (q / (uint32_t)n_stages) & 1u;
                    mbar_wait(smem_u32(&empty[s]), ph ^ 1u);
                    mbar_expect_tx(smem_u32(&fullB[s]), b_stage_bytes);
                    tma_bulk_load(smem_u32(sB + (size_t)s * b_stage_bytes), (const unsigned char *)btab + (size_t)c * b_stage_bytes,
                                  b_stage_bytes, smem_u32(&fullB[s]));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)tmem_cols) : "memory");
}

}  // namespace

// 0 = not applicable (mixed kind, too many table rows, too little shared memory): the caller uses the gather kernel.
// Otherwise the launch has been issued and *used = 1.
int lb_launch_score_rows_tc(Engine *e, int kid, uint64_t b, uint64_t en, const float *d_tab, const int32_t *row_base,
                            int K_total, float *d_out, int planes, int *used) {
    *used = 0;
    const KindHost &k = e->kinds[kid];
    const int nf = (int)k.fids.size();
    if (!nf || K_total < 1 || K_total > 65000 || k.G < 1 || k.G > 256 || en <= b) return 0;
    for (int f : k.fids) {
        const int t = e->specs[f].type;
        if (t != LB_BB && t != LB_DD && t != LB_DPD) return 0;         // GP / NICH stay on the CUDA cores
    }
    const int n_pad = std::max(16, (k.G + 15) / 16 * 16);
    const int n_chunks = (K_total + TC_KC - 1) / TC_KC;
    const size_t b_stage = (size_t)planes * n_pad * TC_KC * 2;
    const size_t fixed = 4 * 32 * 32 * 4 + 2 * (size_t)((nf + 1) & ~1) * TC_ROWS * 2 + (size_t)n_pad * 4 + sizeof(TcFeat) * (size_t)nf + 16 +
                         8 * (3 * 8 + 8) + 16 + 128;
    int n_stages = 0;
    for (int s = 6; s >= 2; --s)
        if (fixed + (size_t)s * (TC_A_BYTES + b_stage) <= 220 * 1024) { n_stages = s; break; }
    if (!n_stages) return 0;
    n_stages = std::min(n_stages, std::max(2, n_chunks));
    int tmem_cols = 32;
    while (tmem_cols < 2 * ((n_pad + 31) & ~31)) tmem_cols *= 2;
    if (tmem_cols > 512) return 0;
    const size_t smem = fixed + (size_t)n_stages * (TC_A_BYTES + b_stage);

    // per-feature producer records and the features each K chunk overlaps
    std::vector<TcFeat> tf(nf);
    int n8 = 0;                                  // the kind's 8-bit columns come first (feature ids ascend)
    for (int f : k.fids) n8 += f < e->F8;
    std::vector<uint16_t> lo(n_chunks, 0xFFFF), hi(n_chunks, 0);
    for (int j = 0; j < nf; ++j) {
        const int f = k.fids[j];
        const lb_feature_spec &s = e->specs[f];
        const bool wide = f >= e->F8;
        tf[j] = TcFeat{wide ? f - e->F8 : f, wide ? 1 : 0, f, row_base[j]};
        const int n = s.type == LB_BB ? 2 : s.dim;
        const int k0 = row_base[j] - 1, k1 = k0 + n - 1;
        for (int c = k0 / TC_KC; c <= k1 / TC_KC && c < n_chunks; ++c) {
            lo[c] = std::min<uint16_t>(lo[c], (uint16_t)j);
            hi[c] = std::max<uint16_t>(hi[c], (uint16_t)j);
        }
    }
    for (int c = 0; c < n_chunks; ++c) if (lo[c] == 0xFFFF) { lo[c] = 1; hi[c] = 0; }
    const size_t o_lo = (sizeof(TcFeat) * (size_t)nf + 255) / 256 * 256;
    const size_t o_hi = o_lo + (2 * (size_t)n_chunks + 255) / 256 * 256;
    const size_t o_tab = o_hi + (2 * (size_t)n_chunks + 255) / 256 * 256;
    const size_t need = o_tab + (size_t)n_chunks * b_stage + 256;
    if (need > e->tc_aux_bytes) {          // (the caller's output lives in d_scratch)
        if (e->d_tc_aux) cudaFree(e->d_tc_aux);
        e->d_tc_aux = nullptr;
        e->tc_aux_bytes = 0;
        LB_CUDA(cudaMalloc(&e->d_tc_aux, need + need / 4));
        e->tc_aux_bytes = need + need / 4;
    }
    char *sc = (char *)e->d_tc_aux;
    LB_CUDA(cudaMemcpyAsync(sc, tf.data(), sizeof(TcFeat) * (size_t)nf, cudaMemcpyHostToDevice, e->stream));
    LB_CUDA(cudaMemcpyAsync(sc + o_lo, lo.data(), 2 * (size_t)n_chunks, cudaMemcpyHostToDevice, e->stream));
    LB_CUDA(cudaMemcpyAsync(sc + o_hi, hi.data(), 2 * (size_t)n_chunks, cudaMemcpyHostToDevice, e->stream));
    __nv_bfloat16 *d_btab = (__nv_bfloat16 *)(sc + o_tab);
    {
        const uint64_t tot = (uint64_t)n_chunks * TC_KC * n_pad;
        const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((tot + 255) / 256, (uint64_t)e->n_sm * 8));
        k_score_table_tc<<<grid, 256, 0, e->stream>>>(d_tab, k.Gcap, k.G, K_total, n_pad, planes, n_chunks, d_btab);
        e->launches[LB_L_TOTAL]++; e->launches[LB_L_SCORE]++;
    }
    LB_CUDA(cudaFuncSetAttribute(k_score_rows_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint64_t n_tiles = (en - b + TC_ROWS - 1) / TC_ROWS;
    const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>(n_tiles, (uint64_t)e->n_sm));
    RowsDev rows{e->R, e->W, e->d_cat8, e->d_wide, e->d_mask};
    k_score_rows_tc<<<grid, TC_THREADS, smem, e->stream>>>(e->d_kinds, kid, (const TcFeat *)sc, (const uint16_t *)(sc + o_lo),
                                                         (const uint16_t *)(sc + o_hi), d_btab, rows, b, en, e->empty_group_count,
                                                         K_total, n_pad, planes, n_stages, tmem_cols, n8, d_out);
    LB_CUDA(cudaGetLastError());
    e->launches[LB_L_TOTAL]++; e->launches[LB_L_SCORE]++;
    *used = 1;
    return 0;
}
