// lb_struct.cu -- kind kernel and hyper kernel of libloomb200.so.
//
//   kind_proposer_score     bulk KindKernel::add_to_kind_proposer (src/kind_kernel.hpp:193-217):
//                           every assigned row, with ALL its features, into group g_k(row) of
//                           every kind's all-feature table  [k_prop_accumulate, K4]
//                           + KindProposer score phase (src/kind_proposer.cc:336-349):
//                           L[f][k] = sum_g log marginal     [k_prop_score, K5]
//   kind_run                block Pitman-Yor sampler (src/kind_proposer.cc:269-300, host, tiny)
//                           + move_features (src/kind_kernel.cc:83-116): moved features take the
//                           proposer's statistics (src/product_mixture.cc:886-915)
//   init_featureless_kinds  src/kind_kernel.cc:159-227
//   hyper_run               HyperKernel::run (src/hyper_kernel.cc:195-235), grid Gibbs
#include <algorithm>
#include <cmath>
#include <atomic>
#include <cstring>
#include <thread>

#include "lb_internal.h"
#include "lb_dpd_hyper.h"

#define FULL 0xffffffffu

template <class T> static int dalloc(T **p, size_t n) {
    *p = nullptr;
    LB_CUDA(cudaMalloc((void **)p, std::max<size_t>(n, 1) * sizeof(T)));
    return 0;
}

// Proposer tables are invalidated (prop_valid = false) whenever the kinds change, but the
// device buffers are kept and reused by the next proposal round.
void lb_free_proposer(Engine *e) {
    e->prop_valid = false;
    e->prop_accumulated = false;
    e->prop_scores.clear();
}
void lb_release_proposer(Engine *e) {
    if (e->d_prop_ints) { cudaFree(e->d_prop_ints); e->d_prop_ints = nullptr; }
    if (e->d_prop_flts) { cudaFree(e->d_prop_flts); e->d_prop_flts = nullptr; }
    if (e->d_prop_packed) { cudaFree(e->d_prop_packed); e->d_prop_packed = nullptr; }
    if (e->d_desc) { cudaFree(e->d_desc); e->d_desc = nullptr; }
    e->prop_cap_ints = e->prop_cap_flts = e->prop_cap_packed = e->d_desc_cap = 0;
    e->prop.clear();
    e->prop_valid = false;
    e->prop_scores.clear();
    if (e->d_prop_feats) { cudaFree(e->d_prop_feats); e->d_prop_feats = nullptr; }
    if (e->d_prop_scores) { cudaFree(e->d_prop_scores); e->d_prop_scores = nullptr; }
}

// ---------------------------------------------------------------------------
// K4: proposer accumulate = k_accum_feature (lb_accum.cuh) over ALL features into the
// all-feature tables [rows][pst] of every kind, with that kind's packed assignments.
// ---------------------------------------------------------------------------

// ---------------------------------------------------------------------------
// K5: L[f][k] = sum over groups of the log marginal likelihood of feature f's
// statistics in kind k.  One CTA per (f, k); DD/DPD spread (value, group)
// pairs over the threads, the others one group per thread.
// ---------------------------------------------------------------------------
#include "lb_marginal.cuh"

struct PropKindDev {       // device view of one kind's proposer tables
    uint32_t *ints;
    double *flts;
    int pst, G;
};

// sums -> moments, once per proposal round and on EVERY rank (move_features reads these tables,
// src/product_mixture.cc:886-915): NICH (sum x, sum x^2) -> (mean, count_times_variance); thread per (feature, kind, group)
__global__ void k_prop_moments(const FeatDev *pfeats, const int32_t *nich_fids, int n_nich, const PropKindDev *pk, int K) {
    const int kid = blockIdx.y;
    const PropKindDev p = pk[kid];
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n_nich * p.G; idx += gridDim.x * blockDim.x) {
        const FeatDev &f = pfeats[nich_fids[idx / p.G]];
        const int g = idx % p.G;
        const double n = (double)p.ints[(size_t)f.int_row * p.pst + g];
        double *fl = p.flts + (size_t)f.flt_row * p.pst + g;
        if (n == 0.0) { fl[0] = 0.0; fl[p.pst] = 0.0; }
        else {
            const double mean = fl[0] / n;
            fl[p.pst] = fmax(fl[p.pst] - n * mean * mean, 0.0);
            fl[0] = mean;
        }
    }
}

// lgamma of every Dirichlet concentration, once per proposal round: lga[aux_off + v] = lgamma(a_v); the totals
// lgamma(A) sit in lgA[f].  K5 then evaluates lgamma(a + n) - lgamma(a) with ONE Stirling / product form per count.
__global__ void k_prop_lgamma_table(const FeatDev *pfeats, int F, const float *aux, float *lga, float *lgA) {
    const int fid = blockIdx.x;
    const FeatDev &f = pfeats[fid];
    if (f.type != LB_DD && f.type != LB_DPD) return;
    for (int v = threadIdx.x; v < f.dim; v += blockDim.x)
        lga[f.aux_off + v] = lgammaf(f.a_scale * aux[f.aux_off + v]);
    if (threadIdx.x == 0) lgA[fid] = lgammaf(f.a_total);
}

// K5: L[f][k] = sum_g log marginal(statistics of feature f in kind k), src/product_mixture.cc:594-620.
// One CTA per (feature, kind).  DD / DPD read their [dim + 1][pst] count table with 128-bit loads (the padding
// columns hold zeros and add nothing); the table is read exactly once and nothing is written back, so the kernel
// is bounded by the table stream (stat bytes x G per feature-score).
__global__ void __launch_bounds__(256)
k_prop_score(const FeatDev *pfeats, const float *aux, const float *lga, const float *lgA, const PropKindDev *pk,
             int K, int f_begin, float *out) {
    __shared__ double s_red[8];
    const int fid = f_begin + blockIdx.x, kid = blockIdx.y;
    const FeatDev &f = pfeats[fid];
    const PropKindDev p = pk[kid];
    const int G = p.G, pst = p.pst;
    const uint32_t *in = p.ints + (size_t)f.int_row * pst;
    const double *fl = p.flts + (size_t)f.flt_row * pst;
    double acc = 0;
    if (f.type == LB_DD || f.type == LB_DPD) {
        const int q4 = pst >> 2;                         // uint4 chunks per row (pst is a multiple of 32)
        const int total = (f.dim + 1) * q4;
        float part = 0.f;
        for (int idx = threadIdx.x; idx < total; idx += 256) {
            const int v = idx / q4, c = idx - v * q4;
            const uint4 n = *(const uint4 *)(in + (size_t)v * pst + 4 * c);
            const bool tot = v == f.dim;
            const float a = tot ? f.a_total : f.a_scale * aux[f.aux_off + v];
            const float l = tot ? lgA[fid] : lga[f.aux_off + v];
            float t = 0.f;
            if (n.x) t += lb_lgamma_rise(a, l, n.x);
            if (n.y) t += lb_lgamma_rise(a, l, n.y);
            if (n.z) t += lb_lgamma_rise(a, l, n.z);
            if (n.w) t += lb_lgamma_rise(a, l, n.w);
            part += tot ? -t : t;
        }
        acc = (double)part;
    } else {
        for (int g = threadIdx.x; g < G; g += blockDim.x) acc += lb_marginal_group(f, aux, in + g, fl + g, pst);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(FULL, acc, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_red[w];
        out[(size_t)fid * K + kid] = (float)t;
    }
}

// all-feature FeatDev table: rows are laid out feature by feature over ALL F features
static int build_prop_feats(Engine *e) {
    const int F = e->F;
    e->prop_int_off.assign(F, 0);
    e->prop_flt_off.assign(F, 0);
    e->prop_int_rows = e->prop_flt_rows = 0;
    std::vector<FeatDev> pf(F);
    for (int f = 0; f < F; ++f) {
        const lb_feature_spec &s = e->specs[f];
        e->prop_int_off[f] = e->prop_int_rows; e->prop_int_rows += spec_int_rows(s);
        e->prop_flt_off[f] = e->prop_flt_rows; e->prop_flt_rows += spec_flt_rows(s);
        FeatDev &fd = pf[f];
        memset(&fd, 0, sizeof fd);
        fd.type = s.type; fd.dim = s.dim;
        for (int i = 0; i < 4; ++i) fd.hp[i] = s.hp[i];
        fd.aux_off = s.aux_off;
        fd.is_wide = f >= e->F8;
        fd.col = fd.is_wide ? f - e->F8 : f;
        fd.int_row = e->prop_int_off[f]; fd.flt_row = e->prop_flt_off[f]; fd.cache_row = 0;
        fd.a_total = 0.f; fd.a_scale = 1.f;
        if (s.type == LB_DD) {
            double t = 0;
            for (int v = 0; v < s.dim; ++v) t += e->aux[s.aux_off + v];
            fd.a_total = (float)t;
        } else if (s.type == LB_DPD) {
            fd.a_total = s.hp[1]; fd.a_scale = s.hp[1];
        }
    }
    if (e->d_prop_feats) cudaFree(e->d_prop_feats);
    LB_TRY(dalloc(&e->d_prop_feats, F));
    LB_CUDA(cudaMemcpyAsync(e->d_prop_feats, pf.data(), sizeof(FeatDev) * F, cudaMemcpyHostToDevice, e->stream));
    LB_CUDA(cudaStreamSynchronize(e->stream));
    return 0;
}

extern "C" int lb_proposer_accumulate(lb_handle h, uint64_t b, uint64_t en) {
    Engine *e = (Engine *)h;
    LB_CUDA(cudaSetDevice(e->device));
    if (!e->R) return lb_fail("proposer_accumulate before set_rows");
    if (b > en || en > e->R) return lb_fail("bad row range");
    LB_TRY(lb_sync_model(e));
    lb_free_proposer(e);
    LB_TRY(build_prop_feats(e));
    const int F = e->F, K = (int)e->kinds.size();
    if ((int)e->prop.size() < K) e->prop.resize(K);
    std::vector<int> all_features(F);
    for (int f = 0; f < F; ++f) all_features[f] = f;
    // every kind's tables live in one slab each (ints / flts / packed ids): two memsets, one pack
    // launch and one accumulate launch per shared-memory class cover the whole proposal round
    size_t need_i = 0, need_f = 0;
    for (int kid = 0; kid < K; ++kid) {
        ProposerKind &p = e->prop[kid];
        p.G = e->kinds[kid].G;
        p.pst = (p.G + 31) & ~31;
        need_i += (size_t)e->prop_int_rows * p.pst;
        need_f += (size_t)e->prop_flt_rows * p.pst;
    }
    const size_t need_p = 4 * (size_t)e->R * K;
    if (need_i > e->prop_cap_ints) {
        if (e->d_prop_ints) cudaFree(e->d_prop_ints);
        e->d_prop_ints = nullptr;
        LB_TRY(dalloc(&e->d_prop_ints, need_i + need_i / 4));
        e->prop_cap_ints = need_i + need_i / 4;
    }
    if (need_f > e->prop_cap_flts) {
        if (e->d_prop_flts) cudaFree(e->d_prop_flts);
        e->d_prop_flts = nullptr;
        LB_TRY(dalloc(&e->d_prop_flts, need_f + need_f / 4));
        e->prop_cap_flts = need_f + need_f / 4;
    }
    if (need_p > e->prop_cap_packed) {
        if (e->d_prop_packed) cudaFree(e->d_prop_packed);
        e->d_prop_packed = nullptr;
        LB_TRY(dalloc(&e->d_prop_packed, need_p + need_p / 4));
        e->prop_cap_packed = need_p + need_p / 4;
    }
    e->prop_used_ints = need_i; e->prop_used_flts = need_f;
    if (need_i) LB_CUDA(cudaMemsetAsync(e->d_prop_ints, 0, 4 * need_i, e->stream));
    if (need_f) LB_CUDA(cudaMemsetAsync(e->d_prop_flts, 0, 8 * need_f, e->stream));
    std::vector<PackSrc> src(K);
    std::vector<AccTarget> targets(K);
    size_t off_i = 0, off_f = 0;
    for (int kid = 0; kid < K; ++kid) {
        const KindHost &k = e->kinds[kid];
        ProposerKind &p = e->prop[kid];
        p.ints = e->d_prop_ints + off_i;
        p.flts = e->d_prop_flts + off_f;
        p.packed = e->d_prop_packed + 4 * (size_t)e->R * kid;
        off_i += (size_t)e->prop_int_rows * p.pst;
        off_f += (size_t)e->prop_flt_rows * p.pst;
        const int width = lb_packed_width(k.G);
        src[kid] = PackSrc{k.assign, k.g2p, p.packed, width};
        targets[kid] = AccTarget{p.packed, p.ints, p.flts, p.pst, p.G, width};
    }
    LB_TRY(lb_launch_pack_assign(e, src, b, en, LB_L_KIND));
    LB_TRY(lb_launch_accum_features(e, e->d_prop_feats, all_features, targets, b, en, +1, LB_L_KIND));
    LB_CUDA(cudaGetLastError());
    LB_CUDA(cudaStreamSynchronize(e->stream));
    e->prop_accumulated = true;
    return 0;
}

extern "C" int lb_proposer_tables(lb_handle h, int32_t kind, uint32_t **ints, uint64_t *n_ints, double **flts,
                                  uint64_t *n_flts) {
    Engine *e = (Engine *)h;
    if (!e->prop_accumulated || kind < 0 || kind >= (int)e->kinds.size() || kind >= (int)e->prop.size())
        return lb_fail("proposer_tables: call proposer_accumulate first");
    const ProposerKind &p = e->prop[kind];
    *ints = p.ints; *n_ints = (uint64_t)e->prop_int_rows * p.pst;
    *flts = p.flts; *n_flts = (uint64_t)e->prop_flt_rows * p.pst;
    return 0;
}

// Score phase over the features [f_lo, f_hi) of the (complete) proposer tables: sums -> moments for every NICH
// feature (all features, whatever the range: the tables must be whole on every rank), then K5 for the range.
// Scores stay on the device in d_prop_scores [F_pad][K]; nothing is synchronised.
int lb_proposer_score_range(Engine *e, int f_lo, int f_hi, int F_pad) {
    const int F = e->F, K = (int)e->kinds.size();
    if ((size_t)F_pad * K > e->prop_scores_cap) {
        if (e->d_prop_scores) cudaFree(e->d_prop_scores);
        LB_TRY(dalloc(&e->d_prop_scores, (size_t)F_pad * K));
        e->prop_scores_cap = (size_t)F_pad * K;
    }
    std::vector<PropKindDev> pk(K);
    for (int kid = 0; kid < K; ++kid) pk[kid] = PropKindDev{e->prop[kid].ints, e->prop[kid].flts, e->prop[kid].pst, e->prop[kid].G};
    std::vector<int32_t> nich;
    for (int f = 0; f < F; ++f) if (e->specs[f].type == LB_NICH) nich.push_back(f);
    // scratch: [PropKindDev K][nich fids][lga aux.size()][lgA F]
    const size_t o_nich = (sizeof(PropKindDev) * (size_t)K + 255) / 256 * 256;
    const size_t o_lga = o_nich + (4 * nich.size() + 255) / 256 * 256;
    const size_t o_lgA = o_lga + (4 * e->aux.size() + 255) / 256 * 256;
    LB_TRY(lb_ensure_scratch(e, o_lgA + 4 * (size_t)F + 256));
    char *sc = (char *)e->d_scratch;
    LB_CUDA(cudaMemcpyAsync(sc, pk.data(), sizeof(PropKindDev) * (size_t)K, cudaMemcpyHostToDevice, e->stream));
    if (!nich.empty()) {
        LB_CUDA(cudaMemcpyAsync(sc + o_nich, nich.data(), 4 * nich.size(), cudaMemcpyHostToDevice, e->stream));
        int maxG = 1;
        for (int kid = 0; kid < K; ++kid) maxG = std::max(maxG, e->prop[kid].G);
        const unsigned gx = (unsigned)std::max<size_t>(1, std::min<size_t>((nich.size() * (size_t)maxG + 255) / 256, 64));
        k_prop_moments<<<dim3(gx, (unsigned)K), 256, 0, e->stream>>>(e->d_prop_feats, (const int32_t *)(sc + o_nich), (int)nich.size(),
                                                                   (const PropKindDev *)sc, K);
        e->launches[LB_L_TOTAL]++; e->launches[LB_L_KIND]++;
    }
    k_prop_lgamma_table<<<F, 128, 0, e->stream>>>(e->d_prop_feats, F, e->d_aux, (float *)(sc + o_lga), (float *)(sc + o_lgA));
    e->launches[LB_L_TOTAL]++; e->launches[LB_L_KIND]++;
    if (f_hi > f_lo) {
        dim3 grid((unsigned)(f_hi - f_lo), (unsigned)K);
        LB_CUDA(cudaEventRecord(e->kev0, e->stream));
        k_prop_score<<<grid, 256, 0, e->stream>>>(e->d_prop_feats, e->d_aux, (const float *)(sc + o_lga), (const float *)(sc + o_lgA),
                                                 (const PropKindDev *)sc, K, f_lo, e->d_prop_scores);
        LB_CUDA(cudaEventRecord(e->kev1, e->stream));
        e->kev_pending = true;
        e->launches[LB_L_TOTAL]++; e->launches[LB_L_KIND]++;
    }
    LB_CUDA(cudaGetLastError());
    return 0;
}

// scores device -> host mirror; marks the proposer tables finished
int lb_proposer_fetch_scores(Engine *e, float *out) {
    const int F = e->F, K = (int)e->kinds.size();
    e->prop_scores.resize((size_t)F * K);
    LB_CUDA(cudaMemcpyAsync(e->prop_scores.data(), e->d_prop_scores, 4 * (size_t)F * K, cudaMemcpyDeviceToHost, e->stream));
    LB_CUDA(cudaStreamSynchronize(e->stream));
    if (e->kev_pending) {     // K5's own device time (family LB_L_KIND)
        float ms = 0.f;
        LB_CUDA(cudaEventElapsedTime(&ms, e->kev0, e->kev1));
        e->kernel_ms[LB_L_KIND] += ms; e->kernel_ms[LB_L_TOTAL] += ms;
        e->kev_pending = false;
    }
    if (out) memcpy(out, e->prop_scores.data(), 4 * (size_t)F * K);
    e->prop_accumulated = false;
    e->prop_valid = true;
    return 0;
}

extern "C" int lb_proposer_finish(lb_handle h, float *out) {
    Engine *e = (Engine *)h;
    LB_CUDA(cudaSetDevice(e->device));
    if (!e->prop_accumulated) return lb_fail("proposer_finish: call proposer_accumulate first");
    LB_TRY(lb_proposer_score_range(e, 0, e->F, e->F));
    return lb_proposer_fetch_scores(e, out);
}

extern "C" int lb_kind_proposer_score(lb_handle h, float *out) {
    Engine *e = (Engine *)h;
    LB_TRY(lb_proposer_accumulate(h, 0, e->R));
    return lb_proposer_finish(h, out);
}

// ---------------------------------------------------------------------------
// block Pitman-Yor sampler (src/kind_proposer.cc:131-300); F x K is tiny, host
// ---------------------------------------------------------------------------
static double likelihood_empty(double alpha, double d, int K, int n_empty) {
    return n_empty ? (alpha + d * (K - n_empty)) / n_empty : 0.0;
}
static int sample_from_likelihoods(double u, const double *p, int n, double total) {
    double t = u * total, c = 0;
    int last = 0;
    for (int i = 0; i < n; ++i) {
        if (p[i] > 0) last = i;
        c += p[i];
        if (c > t) return i;
    }
    return last;
}
static void block_py(double alpha, double d, int F, int K, const std::vector<double> &lik,
                     std::vector<uint32_t> &assign, int iterations, uint64_t seed) {
    std::vector<uint32_t> counts(K, 0);
    std::vector<double> prior(K), post(K);
    int n_empty = 0;
    for (int f = 0; f < F; ++f) counts[assign[f]]++;
    for (int k = 0; k < K; ++k) if (!counts[k]) ++n_empty;
    for (int k = 0; k < K; ++k) prior[k] = counts[k] ? counts[k] - d : likelihood_empty(alpha, d, K, n_empty);
    for (int it = 0; it < iterations; ++it)
        for (int f = 0; f < F; ++f) {
            int k = (int)assign[f];
            if (--counts[k] == 0) {
                ++n_empty;
                const double le = likelihood_empty(alpha, d, K, n_empty);
                for (int j = 0; j < K; ++j) if (!counts[j]) prior[j] = le;
            } else prior[k] = counts[k] - d;
            double total = 0;
            for (int j = 0; j < K; ++j) total += post[j] = prior[j] * lik[(size_t)f * K + j];
            const double u = (double)lb_uniform(seed, 1u, (uint64_t)it * F + f, 0u);
            k = sample_from_likelihoods(u, post.data(), K, total);
            assign[f] = (uint32_t)k;
            if (counts[k]++ == 0) {
                --n_empty;
                const double le = likelihood_empty(alpha, d, K, n_empty);
                for (int j = 0; j < K; ++j) if (!counts[j]) prior[j] = le;
            }
            prior[k] = counts[k] - d;
        }
}

// copy `rows` table rows [.][G] between tables of different strides
template <class T>
__global__ void k_copy_rows(const T *src, int src_row, int src_st, T *dst, int dst_row, int dst_st, int rows, int G) {
    const int total = rows * G;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int r = idx / G, g = idx % G;
        dst[(size_t)(dst_row + r) * dst_st + g] = src[(size_t)(src_row + r) * src_st + g];
    }
}

// move_features: rebuild kind `kid`'s tables for the new feature list
static int kind_relayout(Engine *e, int kid, const std::vector<int> &fids) {
    KindHost &k = e->kinds[kid];
    KindHost nk;
    nk.alpha = k.alpha; nk.d = k.d;
    LB_TRY(lb_kind_layout(e, nk, fids));
    LB_TRY(lb_kind_alloc(e, nk, k.G, k.Gcap));
    nk.sample_size = k.sample_size;
    // clustering state and the id tracker carry over unchanged
    LB_CUDA(cudaMemcpyAsync(nk.counts, k.counts, 4 * (size_t)k.Gcap, cudaMemcpyDeviceToDevice, e->stream));
    LB_CUDA(cudaMemcpyAsync(nk.shifted, k.shifted, 4 * (size_t)k.Gcap, cudaMemcpyDeviceToDevice, e->stream));
    LB_CUDA(cudaMemcpyAsync(nk.p2g, k.p2g, 4 * (size_t)k.Gcap, cudaMemcpyDeviceToDevice, e->stream));
    if (nk.g2p_cap != k.g2p_cap) {
        cudaFreeAsync(nk.g2p, e->stream);
        LB_CUDA(cudaMallocAsync((void **)&nk.g2p, 4 * (size_t)k.g2p_cap, e->stream));
        nk.g2p_cap = k.g2p_cap;
    }
    LB_CUDA(cudaMemcpyAsync(nk.g2p, k.g2p, 4 * (size_t)k.g2p_cap, cudaMemcpyDeviceToDevice, e->stream));
    nk.next_global = k.next_global;
    const ProposerKind &p = e->prop[kid];
    const int G = k.G;
    for (size_t j = 0; j < fids.size(); ++j) {
        const int f = fids[j];
        const lb_feature_spec &s = e->specs[f];
        int oj = -1;
        for (size_t q = 0; q < k.fids.size(); ++q) if (k.fids[q] == f) oj = (int)q;
        const int ir = spec_int_rows(s), fr = spec_flt_rows(s);
        const int grid = std::max(1, std::min((ir * G + 255) / 256, 256));
        if (oj >= 0) {
            k_copy_rows<uint32_t><<<grid, 256, 0, e->stream>>>(k.ints, k.int_off[oj], k.Gcap, nk.ints, nk.int_off[j], nk.Gcap, ir, G);
            if (fr) k_copy_rows<double><<<1, 256, 0, e->stream>>>(k.flts, k.flt_off[oj], k.Gcap, nk.flts, nk.flt_off[j], nk.Gcap, fr, G);
        } else {
            k_copy_rows<uint32_t><<<grid, 256, 0, e->stream>>>(p.ints, e->prop_int_off[f], p.pst, nk.ints, nk.int_off[j], nk.Gcap, ir, G);
            if (fr) k_copy_rows<double><<<1, 256, 0, e->stream>>>(p.flts, e->prop_flt_off[f], p.pst, nk.flts, nk.flt_off[j], nk.Gcap, fr, G);
        }
        e->launches[LB_L_TOTAL] += fr ? 2 : 1; e->launches[LB_L_KIND] += fr ? 2 : 1;
    }
    LB_CUDA(cudaGetLastError());
    LB_CUDA(cudaStreamSynchronize(e->stream));
    nk.assign = k.assign; nk.assign_pooled = k.assign_pooled;
    lb_kind_release(e, k);
    k = nk;
    return 0;
}

extern "C" int lb_kind_run(lb_handle h, int32_t iterations, uint64_t seed, uint32_t *new_f2k, int32_t *n_changed) {
    Engine *e = (Engine *)h;
    LB_CUDA(cudaSetDevice(e->device));
    if (iterations <= 0) return lb_fail("kind_run: iterations must be > 0");
    const int F = e->F, K = (int)e->kinds.size();
    if (!e->prop_valid || e->prop_scores.size() != (size_t)F * K || (int)e->prop.size() < K)
        return lb_fail("kind_run: call kind_proposer_score first");
    for (int k = 0; k < K; ++k)
        if (e->prop[k].G != e->kinds[k].G) return lb_fail("kind_run: stale proposer tables");
    std::vector<double> lik((size_t)F * K);
    for (int f = 0; f < F; ++f) {  // scores_to_likelihoods (kind_proposer.cc:347)
        float m = -INFINITY;
        for (int k = 0; k < K; ++k) m = std::max(m, e->prop_scores[(size_t)f * K + k]);
        for (int k = 0; k < K; ++k) lik[(size_t)f * K + k] = std::exp((double)e->prop_scores[(size_t)f * K + k] - (double)m);
    }
    std::vector<uint32_t> assign(e->f2k);
    block_py(e->topo[0], e->topo[1], F, K, lik, assign, iterations, seed);
    int changed = 0;
    for (int f = 0; f < F; ++f) if (assign[f] != e->f2k[f]) ++changed;
    for (int k = 0; k < K; ++k) {
        std::vector<int> fids;
        for (int f = 0; f < F; ++f) if (assign[f] == (uint32_t)k) fids.push_back(f);
        if (fids != e->kinds[k].fids) LB_TRY(kind_relayout(e, k, fids));
    }
    e->f2k = assign;
    e->model_dirty = true;
    LB_TRY(lb_sync_model(e));
    for (int k = 0; k < K; ++k) LB_TRY(lb_launch_refresh(e, k));   // init_cache (kind_kernel.cc:257-315)
    LB_CUDA(cudaStreamSynchronize(e->stream));
    if (new_f2k) memcpy(new_f2k, assign.data(), 4 * (size_t)F);
    if (n_changed) *n_changed = changed;
    lb_free_proposer(e);
    return 0;
}

// ---------------------------------------------------------------------------
// Pitman-Yor seating of the ephemeral kinds (src/kind_kernel.cc:159-189): one sequential pass over the assigned rows
// per kind.  The kinds are independent and every draw is keyed by (seed, row, kind id), so they are seated side by
// side on host threads; the arithmetic (and therefore every pick) is the single-threaded loop's.
// ---------------------------------------------------------------------------
struct Seating {
    int count = 0, first_kid = 0;
    uint64_t seed = 0, R = 0;
    std::vector<uint64_t> assigned;                     // bit per row: assigned in kind 0 when the seating was requested
    std::vector<std::pair<float, float>> py;            // (alpha, d) of every new kind
    std::vector<std::vector<uint32_t>> picks, counts;   // results
    bool same_request(const Seating &o) const {
        return count == o.count && first_kid == o.first_kid && seed == o.seed && R == o.R && py == o.py && assigned == o.assigned;
    }
};
struct PendingSeating {
    Seating s;
    std::vector<uint32_t> a0;
    std::thread th;
};

static void seating_describe(Engine *e, Seating &s, int count, uint64_t seed, int first_kid, const std::vector<uint32_t> &a0) {
    s.count = count; s.seed = seed; s.first_kid = first_kid; s.R = count > 0 ? e->R : 0;
    s.assigned.assign(count > 0 ? (e->R + 63) / 64 : 0, 0ull);
    if (count > 0)
        for (uint64_t r = 0; r < e->R; ++r) if (a0[r] != LB_UNASSIGNED) s.assigned[r >> 6] |= 1ull << (r & 63);
    const std::vector<float> &grid = e->hp[1];   // clustering grid [n][2]
    s.py.assign(std::max(count, 0), {e->kinds[0].alpha, e->kinds[0].d});
    for (int c = 0; c < count; ++c)
        if (!grid.empty()) {   // sample_clustering_prior (infer_grid.hpp:127-139)
            const int n = (int)grid.size() / 2;
            const double u = (double)lb_uniform(seed, 3u, 0xFFFFFFFFFFFFFFFFull, (uint32_t)(first_kid + c));
            int i = (int)(u * n);
            if (i >= n) i = n - 1;
            s.py[c] = {grid[2 * i], grid[2 * i + 1]};
        }
}

static void seat_kinds(Seating &s, const std::vector<uint32_t> &a0) {
    const int count = s.count;
    s.picks.assign(std::max(count, 0), {});
    s.counts.assign(std::max(count, 0), {});
    if (count <= 0) return;
    auto seat = [&](int c) {
        const int kid = s.first_kid + c;
        const double alpha = s.py[c].first, d = s.py[c].second;
        std::vector<uint32_t> &pick = s.picks[c], &counts = s.counts[c];
        pick.assign(s.R, LB_UNASSIGNED);
        int m = 0;
        uint64_t seated = 0;
        for (uint64_t r = 0; r < s.R; ++r) {
            if (a0[r] == LB_UNASSIGNED) continue;
            // weights counts[g] - d for the occupied tables, alpha + d m for a new one: they sum to seated + alpha in
            // closed form, so the scan stops at the pick (the oracle does the same arithmetic)
            const double total = (double)seated + alpha;
            const double u = (double)lb_uniform(s.seed, 3u, r, (uint32_t)kid);
            const double t = u * total;
            double cum = 0;
            int g = m;
            for (int i = 0; i < m; ++i) {
                cum += (double)counts[i] - d;
                if (cum > t) { g = i; break; }
            }
            if (g == m) { counts.push_back(0); ++m; }
            counts[g]++;
            ++seated;
            pick[r] = (uint32_t)g;
        }
    };
    const int n_thr = std::max(1, std::min<int>(count, (int)std::thread::hardware_concurrency()));
    std::vector<std::thread> pool;
    std::atomic<int> next(0);
    for (int t = 0; t < n_thr; ++t)
        pool.emplace_back([&] { for (int c = next++; c < count; c = next++) seat(c); });
    for (auto &t : pool) t.join();
}

void lb_release_seating(Engine *e) {
    PendingSeating *p = (PendingSeating *)e->seating;
    if (!p) return;
    if (p->th.joinable()) p->th.join();
    delete p;
    e->seating = nullptr;
}

// Draw the NEXT lb_init_featureless_kinds(count, seed) ahead of time on background host threads: the seating depends
// only on the seed, the kind ids and WHICH rows are assigned -- not on where they sit -- so a driver calls this before
// it launches a pass that keeps the assigned set (a REMOVE + ADD pass), and the host work runs under the device sweep.
extern "C" int lb_prepare_featureless_kinds(lb_handle h, int32_t count, uint64_t seed) {
    Engine *e = (Engine *)h;
    LB_CUDA(cudaSetDevice(e->device));
    lb_release_seating(e);
    if (count <= 0 || !e->R || e->kinds.empty() || e->kinds[0].fids.empty()) return 0;
    PendingSeating *p = new PendingSeating;
    p->a0.resize(e->R);
    LB_CUDA(cudaMemcpyAsync(p->a0.data(), e->kinds[0].assign, 4 * (size_t)e->R, cudaMemcpyDeviceToHost, e->stream));
    LB_CUDA(cudaStreamSynchronize(e->stream));
    int first_kid = 0;
    for (const KindHost &k : e->kinds) first_kid += !k.fids.empty();     // the ids the new kinds get once the featureless ones are gone
    seating_describe(e, p->s, count, seed, first_kid, p->a0);
    e->seating = p;
    p->th = std::thread([p] { seat_kinds(p->s, p->a0); });
    return 0;
}

// ---------------------------------------------------------------------------
// init_featureless_kinds (src/kind_kernel.cc:159-227).  The Pitman-Yor prior
// partition of the assigned rows is a sequential seating process
// (Clustering::PitmanYor::sample_assignments); it is drawn on the host in the
// same float64 arithmetic as the oracle and uploaded.
// ---------------------------------------------------------------------------
extern "C" int lb_init_featureless_kinds(lb_handle h, int32_t count, uint64_t seed) {
    Engine *e = (Engine *)h;
    LB_CUDA(cudaSetDevice(e->device));
    lb_free_proposer(e);
    for (int i = (int)e->kinds.size() - 1; i >= 0; --i) {
        if (!e->kinds[i].fids.empty()) continue;
        if (e->kinds.size() == 1) return lb_fail("cannot remove the only kind");
        lb_kind_release(e, e->kinds[i]);
        if (e->kinds[i].assign) {
            if (e->kinds[i].assign_pooled) cudaFreeAsync(e->kinds[i].assign, e->stream);
            else cudaFree(e->kinds[i].assign);
        }
        const int last = (int)e->kinds.size() - 1;
        if (i != last) {   // remove_featureless_kind: packed_remove + patch featureid_to_kindid
            e->kinds[i] = e->kinds[last];
            for (int f : e->kinds[i].fids) e->f2k[f] = (uint32_t)i;
        }
        e->kinds.pop_back();
    }
    std::vector<uint32_t> a0;
    if (count > 0 && e->R) {
        a0.resize(e->R);
        LB_CUDA(cudaMemcpyAsync(a0.data(), e->kinds[0].assign, 4 * (size_t)e->R, cudaMemcpyDeviceToHost, e->stream));
        LB_CUDA(cudaStreamSynchronize(e->stream));
    }
    const int first_kid = (int)e->kinds.size();
    // the seating may have been drawn ahead of time, under the sweep (lb_prepare_featureless_kinds); it is used when it
    // was drawn for exactly this situation (same seed, count, kind ids, clustering, set of assigned rows)
    Seating local, *seating = &local;
    PendingSeating *pend = (PendingSeating *)e->seating;
    if (pend && pend->th.joinable()) pend->th.join();
    {
        Seating want;
        seating_describe(e, want, count, seed, first_kid, a0);
        if (pend && pend->s.same_request(want)) seating = &pend->s;
        else { local = std::move(want); seat_kinds(local, a0); }
    }
    const std::vector<std::vector<uint32_t>> &picks = seating->picks, &all_counts = seating->counts;
    const std::vector<std::pair<float, float>> &py = seating->py;
    for (int c = 0; c < count; ++c) {
        e->kinds.emplace_back();
        KindHost &k = e->kinds.back();
        k.alpha = py[c].first; k.d = py[c].second;
        const std::vector<uint32_t> &pick = picks[c], &counts = all_counts[c];
        const int m = (int)counts.size();
        const int G = m + e->empty_group_count;
        LB_TRY(lb_kind_layout(e, k, std::vector<int>()));
        int cap = 256;
        while (cap < 2 * G) cap *= 2;
        LB_TRY(lb_kind_alloc(e, k, G, cap));
        uint64_t total = 0;
        for (int g = 0; g < m; ++g) total += counts[g];
        k.sample_size = total;
        if (m) LB_CUDA(cudaMemcpyAsync(k.counts, counts.data(), 4 * (size_t)m, cudaMemcpyHostToDevice, e->stream));
        LB_CUDA(cudaMallocAsync((void **)&k.assign, 4 * std::max<size_t>(e->R, 1), e->stream));
        k.assign_pooled = true;
        if (e->R) LB_CUDA(cudaMemcpyAsync(k.assign, pick.data(), 4 * (size_t)e->R, cudaMemcpyHostToDevice, e->stream));
        LB_CUDA(cudaStreamSynchronize(e->stream));
    }
    lb_release_seating(e);                 // the picks are on the device now
    e->model_dirty = true;
    LB_TRY(lb_sync_model(e));
    for (int kid = 0; kid < (int)e->kinds.size(); ++kid)
        if (e->kinds[kid].fids.empty()) LB_TRY(lb_launch_refresh(e, kid));
    LB_CUDA(cudaStreamSynchronize(e->stream));
    return 0;
}

// ---------------------------------------------------------------------------
// hyper kernel
// ---------------------------------------------------------------------------
extern "C" int lb_set_hyper_prior(lb_handle h, const lb_hyper_prior *pr) {
    Engine *e = (Engine *)h;
    for (auto &v : e->hp) v.clear();
    if (!pr) return 0;
    {   // the grid kernels hold one coordinate's scores in 256 slots (loom's default grids have at most 201 points)
        const int32_t ns[] = {pr->n_bb_alpha, pr->n_bb_beta, pr->n_dd_alpha, pr->n_dpd_gamma, pr->n_dpd_alpha, pr->n_gp_alpha,
                              pr->n_gp_inv_beta, pr->n_nich_mu, pr->n_nich_kappa, pr->n_nich_sigmasq, pr->n_nich_nu};
        for (int32_t n : ns)
            if (n > 256) return lb_fail("set_hyper_prior: a feature grid has %d points (limit 256)", (int)n);
    }
    auto put = [&](int i, const float *p, int n, int mult) { if (n > 0) e->hp[i].assign(p, p + (size_t)n * mult); };
    put(0, pr->topology, pr->n_topology, 2); put(1, pr->clustering, pr->n_clustering, 2);
    put(2, pr->bb_alpha, pr->n_bb_alpha, 1); put(3, pr->bb_beta, pr->n_bb_beta, 1);
    put(4, pr->dd_alpha, pr->n_dd_alpha, 1);
    put(5, pr->dpd_gamma, pr->n_dpd_gamma, 1); put(6, pr->dpd_alpha, pr->n_dpd_alpha, 1);
    put(7, pr->gp_alpha, pr->n_gp_alpha, 1); put(8, pr->gp_inv_beta, pr->n_gp_inv_beta, 1);
    put(9, pr->nich_mu, pr->n_nich_mu, 1); put(10, pr->nich_kappa, pr->n_nich_kappa, 1);
    put(11, pr->nich_sigmasq, pr->n_nich_sigmasq, 1); put(12, pr->nich_nu, pr->n_nich_nu, 1);
    return 0;
}

// Pitman-Yor score_counts in float64, host: the count vectors are tiny
static double py_score_counts(double alpha, double d, const uint32_t *counts, int n) {
    double score = 0, total = 0;
    int m = 0;
    for (int i = 0; i < n; ++i)
        if (counts[i]) {
            score += std::log(alpha + d * m);
            score += std::lgamma(counts[i] - d) - std::lgamma(1.0 - d);
            total += counts[i];
            ++m;
        }
    return score + std::lgamma(alpha) - std::lgamma(alpha + total);
}
static int sample_from_scores(double u, std::vector<double> &s) {
    double m = -INFINITY, total = 0;
    for (double v : s) m = std::max(m, v);
    for (double &v : s) { v = std::exp(v - m); total += v; }
    return sample_from_likelihoods(u, s.data(), (int)s.size(), total);
}
static void sample_py_posterior(const std::vector<float> &grid, const uint32_t *counts, int nc, uint64_t seed,
                                uint64_t task, float *alpha, float *d) {
    const int n = (int)grid.size() / 2;
    if (n <= 0) return;
    int i = 0;
    if (n > 1) {
        std::vector<double> scores(n);
        for (int j = 0; j < n; ++j) scores[j] = py_score_counts(grid[2 * j], grid[2 * j + 1], counts, nc);
        i = sample_from_scores((double)lb_uniform(seed, 2u, task, 0u), scores);
    }
    *alpha = grid[2 * i]; *d = grid[2 * i + 1];
}

// Feature-hyper grid Gibbs, one CTA per feature (task 1 + K + f of HyperKernel::run).
// For every coordinate in the order of src/hyper_prior.hpp:38-146 the block scores each
// grid value j by sum_g log marginal(group g | shared with coord := grid[j])
// (InferShared::done, src/infer_grid.hpp:73-89), draws j and commits it.  DD evaluates only
// the terms that depend on the coordinate (the rest is constant in j and cancels in the draw).
struct HyperGrids {
    const float *g[13];
    int n[13];
};

__device__ double hyper_block_sum(double v, double *s_red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_red[w];
    return t;
}

__global__ void __launch_bounds__(256)
k_hyper_features(const KindDev *kinds, FeatDev *feats, const uint32_t *f2k, float *aux, HyperGrids hg,
                 uint64_t seed, int K, const uint8_t *task_mask) {
    __shared__ double s_red[8];
    __shared__ double s_scores[256];
    __shared__ int s_pick;
    const int fid = blockIdx.x;
    if (task_mask && !task_mask[1 + K + fid]) return;      // another rank runs this feature's task
    FeatDev f = feats[fid];
    const KindDev &kd = kinds[f2k[fid]];
    const int G = kd.G, st = kd.Gcap;
    const uint32_t *in = kd.ints + (size_t)f.int_row * st;
    const double *fl = kd.flts + (size_t)f.flt_row * st;
    const uint64_t task = 1ull + (uint64_t)K + (uint64_t)fid;
    uint32_t draw = 0;

    auto choose = [&](int n) -> int {   // softmax draw over s_scores[0..n), float64 like the oracle
        if (threadIdx.x == 0) {
            double m = -INFINITY, total = 0;
            for (int j = 0; j < n; ++j) m = fmax(m, s_scores[j]);
            for (int j = 0; j < n; ++j) { s_scores[j] = exp(s_scores[j] - m); total += s_scores[j]; }
            const double t = (double)lb_uniform(seed, 2u, task, draw) * total;
            double c = 0;
            int pick = 0, last = 0;
            bool found = false;
            for (int j = 0; j < n && !found; ++j) {
                if (s_scores[j] > 0) last = j;
                c += s_scores[j];
                if (c > t) { pick = j; found = true; }
            }
            s_pick = found ? pick : last;
        }
        __syncthreads();
        return s_pick;
    };

    if (f.type == LB_BB || f.type == LB_GP || f.type == LB_NICH) {
        const int first = f.type == LB_BB ? 2 : (f.type == LB_GP ? 7 : 9);
        const int ncoord = f.type == LB_NICH ? 4 : 2;
        for (int c = 0; c < ncoord; ++c) {
            const int n = hg.n[first + c];
            const float *grid = hg.g[first + c];
            if (n <= 0) continue;
            if (n == 1) { f.hp[c] = grid[0]; continue; }
            for (int j0 = 0; j0 < n; j0 += 256) {
                const int nj = min(256, n - j0);
                for (int j = 0; j < nj; ++j) {
                    FeatDev t = f;
                    t.hp[c] = grid[j0 + j];
                    double acc = 0;
                    for (int g = threadIdx.x; g < G; g += blockDim.x) acc += lb_marginal_group(t, aux, in + g, fl + g, st);
                    const double tot = hyper_block_sum(acc, s_red);
                    if (threadIdx.x == 0) s_scores[j] = tot;
                }
                __syncthreads();
                if (n > 256) break;   // grids larger than 256 points are not supported (default max is 201)
            }
            const int pick = choose(min(n, 256));
            ++draw;
            f.hp[c] = grid[pick];
            __syncthreads();
        }
    } else if (f.type == LB_DD) {
        const int n = hg.n[4];
        const float *grid = hg.g[4];
        if (n == 1) {
            for (int v = threadIdx.x; v < f.dim; v += blockDim.x) aux[f.aux_off + v] = grid[0];
        } else if (n > 1) {
            // running total of the alphas in float64, in index order like the oracle's sum
            for (int v = 0; v < f.dim; ++v) {
                double a_rest = 0;
                for (int w = 0; w < f.dim; ++w) if (w != v) a_rest += (double)aux[f.aux_off + w];
                for (int j = 0; j < n && j < 256; ++j) {
                    const float a_v = grid[j];
                    const float a_tot = (float)(a_rest + (double)a_v);
                    double acc = 0;
                    for (int g = threadIdx.x; g < G; g += blockDim.x) {
                        const uint32_t nv = in[(size_t)v * st + g], nt = in[(size_t)f.dim * st + g];
                        if (nv) acc += (double)lb_lgamma_diff(a_v, (float)nv);
                        if (nt) acc -= (double)lb_lgamma_diff(a_tot, (float)nt);
                    }
                    const double tot = hyper_block_sum(acc, s_red);
                    if (threadIdx.x == 0) s_scores[j] = tot;
                }
                __syncthreads();
                const int pick = choose(min(n, 256));
                ++draw;
                if (threadIdx.x == 0) aux[f.aux_off + v] = grid[pick];
                __syncthreads();
            }
        }
    }
    if (threadIdx.x == 0) {
        for (int c = 0; c < 4; ++c) feats[fid].hp[c] = f.hp[c];
    }
}

extern "C" int lb_hyper_run(lb_handle h, uint64_t seed) { return lb_hyper_run_tasks(h, seed, nullptr); }

extern "C" int lb_hyper_run_tasks(lb_handle h, uint64_t seed, const uint8_t *task_mask) {
    Engine *e = (Engine *)h;
    LB_CUDA(cudaSetDevice(e->device));
    LB_TRY(lb_sync_model(e));
    const int K = (int)e->kinds.size(), F = e->F;
    // task 0: topology; tasks 1..K: per-kind clustering (host, float64: tiny count vectors)
    if (!e->hp[0].empty() && (!task_mask || task_mask[0])) {
        std::vector<uint32_t> fc(K);
        for (int k = 0; k < K; ++k) fc[k] = (uint32_t)e->kinds[k].fids.size();
        sample_py_posterior(e->hp[0], fc.data(), K, seed, 0, &e->topo[0], &e->topo[1]);
    }
    if (!e->hp[1].empty()) {
        for (int k = 0; k < K; ++k) {
            if (task_mask && !task_mask[1 + k]) continue;
            KindHost &kh = e->kinds[k];
            std::vector<uint32_t> counts(kh.G);
            LB_CUDA(cudaMemcpyAsync(counts.data(), kh.counts, 4 * (size_t)kh.G, cudaMemcpyDeviceToHost, e->stream));
            LB_CUDA(cudaStreamSynchronize(e->stream));
            sample_py_posterior(e->hp[1], counts.data(), kh.G, seed, (uint64_t)(1 + k), &kh.alpha, &kh.d);
        }
    }
    // tasks 1+K..: per-feature grids on the device
    bool any = false;
    for (int i = 2; i < 13; ++i) if (!e->hp[i].empty()) any = true;
    if (any) {
        HyperGrids hg;
        size_t total = 0;
        for (int i = 0; i < 13; ++i) total += e->hp[i].size();
        std::vector<float> flat(total + 1);
        float *d_flat = nullptr;
        LB_TRY(dalloc(&d_flat, total + 1));
        size_t off = 0;
        for (int i = 0; i < 13; ++i) {
            hg.g[i] = d_flat + off;
            hg.n[i] = (int)e->hp[i].size() / (i < 2 ? 2 : 1);
            std::copy(e->hp[i].begin(), e->hp[i].end(), flat.begin() + off);
            off += e->hp[i].size();
        }
        LB_CUDA(cudaMemcpyAsync(d_flat, flat.data(), 4 * (total + 1), cudaMemcpyHostToDevice, e->stream));
        uint32_t *d_f2k = nullptr;
        LB_TRY(dalloc(&d_f2k, F));
        LB_CUDA(cudaMemcpyAsync(d_f2k, e->f2k.data(), 4 * (size_t)F, cudaMemcpyHostToDevice, e->stream));
        uint8_t *d_mask = nullptr;
        if (task_mask) {
            LB_TRY(dalloc(&d_mask, (size_t)(1 + K + F)));
            LB_CUDA(cudaMemcpyAsync(d_mask, task_mask, (size_t)(1 + K + F), cudaMemcpyHostToDevice, e->stream));
        }
        k_hyper_features<<<F, 256, 0, e->stream>>>(e->d_kinds, e->d_feats, d_f2k, e->d_aux, hg, seed, K, d_mask);
        LB_CUDA(cudaGetLastError());
        e->launches[LB_L_TOTAL]++; e->launches[LB_L_HYPER]++;
        // pull the chosen hypers back into the host mirror
        std::vector<FeatDev> feats(F);
        LB_CUDA(cudaMemcpyAsync(feats.data(), e->d_feats, sizeof(FeatDev) * F, cudaMemcpyDeviceToHost, e->stream));
        if (!e->aux.empty())
            LB_CUDA(cudaMemcpyAsync(e->aux.data(), e->d_aux, 4 * e->aux.size(), cudaMemcpyDeviceToHost, e->stream));
        LB_CUDA(cudaStreamSynchronize(e->stream));
        for (int f = 0; f < F; ++f)
            for (int c = 0; c < 4; ++c) e->specs[f].hp[c] = feats[f].hp[c];
        cudaFree(d_flat);
        cudaFree(d_f2k);
        if (d_mask) cudaFree(d_mask);
    }
    // DPD features (src/hyper_kernel.cc:90-181): auxiliary-variable sampler on the host over the
    // feature's count table (lb_dpd_hyper.h); task id 1 + K + f like every other feature
    if (!e->hp[5].empty() || !e->hp[6].empty()) {
        for (int f = 0; f < F; ++f) {
            lb_feature_spec &s = e->specs[f];
            if (s.type != LB_DPD || s.dim <= 0) continue;
            if (task_mask && !task_mask[1 + K + f]) continue;
            KindHost &kh = e->kinds[e->f2k[f]];
            int j = -1;
            for (size_t q = 0; q < kh.fids.size(); ++q) if (kh.fids[q] == f) j = (int)q;
            if (j < 0 || kh.G <= 0) continue;
            const size_t G = (size_t)kh.G;
            std::vector<uint32_t> table((size_t)(s.dim + 1) * G);
            LB_CUDA(cudaMemcpy2DAsync(table.data(), 4 * G, kh.ints + (size_t)kh.int_off[j] * kh.Gcap, 4 * (size_t)kh.Gcap,
                                      4 * G, (size_t)s.dim + 1, cudaMemcpyDeviceToHost, e->stream));
            LB_CUDA(cudaStreamSynchronize(e->stream));
            const uint64_t task = (uint64_t)(1 + K + f);
            lb_dpd::hyper(table.data(), s.dim, (int)G, G, s.hp, e->aux.data() + s.aux_off, e->hp[5].data(),
                          (int)e->hp[5].size(), e->hp[6].data(), (int)e->hp[6].size(),
                          [&](uint32_t i) { return (double)lb_uniform(seed, 2u, task, i); });
        }
    }
    e->model_dirty = true;   // derived fields (a_total, ...) and cached scorers depend on the hypers
    LB_TRY(lb_sync_model(e));
    for (int k = 0; k < K; ++k) LB_TRY(lb_launch_refresh(e, k));   // mixture.init(shared)
    LB_CUDA(cudaStreamSynchronize(e->stream));
    return 0;
}

// the gather step of the sharded hyper kernel: install hypers chosen elsewhere
extern "C" int lb_set_hypers(lb_handle h, const lb_feature_spec *specs, const float *aux, const float *kind_py,
                             const float *topology_py) {
    Engine *e = (Engine *)h;
    LB_CUDA(cudaSetDevice(e->device));
    if (!e->F) return lb_fail("set_hypers before set_model");
    LB_TRY(lb_sync_model(e));
    if (specs)
        for (int f = 0; f < e->F; ++f) {
            if (specs[f].type != e->specs[f].type || specs[f].dim != e->specs[f].dim || specs[f].aux_off != e->specs[f].aux_off)
                return lb_fail("set_hypers: feature %d does not match the model", f);
            for (int c = 0; c < 4; ++c) e->specs[f].hp[c] = specs[f].hp[c];
        }
    if (aux && !e->aux.empty()) std::copy(aux, aux + e->aux.size(), e->aux.begin());
    if (kind_py)
        for (size_t k = 0; k < e->kinds.size(); ++k) {
            e->kinds[k].alpha = kind_py[2 * k];
            e->kinds[k].d = kind_py[2 * k + 1];
        }
    if (topology_py) { e->topo[0] = topology_py[0]; e->topo[1] = topology_py[1]; }
    e->model_dirty = true;
    LB_TRY(lb_sync_model(e));
    for (int k = 0; k < (int)e->kinds.size(); ++k) LB_TRY(lb_launch_refresh(e, k));
    LB_CUDA(cudaStreamSynchronize(e->stream));
    return 0;
}
