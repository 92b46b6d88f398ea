// lb_query.cu -- the sampling half of loom's query server on the device (sm_100a).
//
// QueryServer::call(Sample) (src/query_server.cc:100-176) scores the conditioning row against
// every kind (K1, lb_launch_score_rows), turns the scores into probabilities, and then per
// sample draws a group per kind and a value per requested feature from that group
// (ProductMixture_::sample_value, src/product_mixture.cc:622-649 -> Group::sample_value of
// distributions 2.0.28).  Here: one small kernel builds the cumulative probabilities of each
// kind, one thread per (sample, kind) inverts them, one thread per (sample, feature) draws the
// value.  Every draw is a pure function of (seed, sample, kind | feature) through Philox, so
// the result does not depend on the launch shape; the value samplers work in float64 (this is
// not a bandwidth-bound path: a request is a few thousand draws).
#include <algorithm>
#include <vector>

#include "lb_internal.h"

struct QSlot {          // one kind that holds requested features
    int32_t kind, G;
    int32_t off;        // first entry of this kind in the score / cdf arrays
};

// cdf[g] = sum_{i <= g} exp(score_i - max): distributions::scores_to_probs without the final
// division (the pick scales its uniform by the total instead).  One block per slot; the sum is
// taken serially in group order, the order sample_from_probs walks.
__global__ void k_query_cdf(const QSlot *slots, const float *scores, double *cdf) {
    const QSlot s = slots[blockIdx.x];
    const float *sc = scores + s.off;
    double *out = cdf + s.off;
    __shared__ float s_max[32];
    float m = -INFINITY;
    for (int g = threadIdx.x; g < s.G; g += blockDim.x) m = fmaxf(m, sc[g]);
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = m;
    __syncthreads();
    m = s_max[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmaxf(m, s_max[w]);
    for (int g = threadIdx.x; g < s.G; g += blockDim.x) out[g] = exp((double)sc[g] - (double)m);
    __syncthreads();
    if (threadIdx.x == 0) {
        double c = 0;
        for (int g = 0; g < s.G; ++g) { c += out[g]; out[g] = c; }
    }
}

// distributions::sample_from_probs on the cumulative form: first g with cdf[g] > u * total
__global__ void k_query_pick(const QSlot *slots, int n_slots, const double *cdf, uint32_t n_samples, int K,
                             uint64_t seed, uint64_t draw_offset, uint32_t *groups) {
    const uint64_t idx = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (idx >= (uint64_t)n_samples * n_slots) return;
    const uint32_t smp = (uint32_t)(idx / n_slots);
    const QSlot s = slots[idx % n_slots];
    const double *c = cdf + s.off;
    const double u = (double)(lb_philox_u32(seed, 5, draw_offset + smp, (uint32_t)s.kind) >> 8) * (1.0 / 16777216.0);
    const double t = u * c[s.G - 1];
    int lo = 0, hi = s.G - 1;           // c is non-decreasing and c[G-1] > t
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (c[mid] > t) hi = mid; else lo = mid + 1;
    }
    groups[(uint64_t)smp * K + s.kind] = (uint32_t)lo;
}

struct Draws {
    uint64_t seed, a;
    uint32_t base, j;
    uint32_t stream = 6u;
    __device__ double u() { return (double)(lb_philox_u32(seed, stream, a, base | j++) >> 8) * (1.0 / 16777216.0); }   // [0,1)
    __device__ double uo() { return u() + 0.5 / 16777216.0; }                                                     // (0,1)
};

// Marsaglia & Tsang (2000), shape boosted above 1 for small shapes
static __device__ double lb_sample_gamma(Draws &d, double shape) {
    double boost = 1.0;
    if (shape < 1.0) { boost = pow(d.uo(), 1.0 / shape); shape += 1.0; }
    const double dd = shape - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * dd);
    for (int it = 0; it < 100; ++it) {
        const double u1 = d.uo(), u2 = d.u(), u3 = d.uo();
        const double z = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
        double v = 1.0 + c * z;
        if (v <= 0) continue;
        v = v * v * v;
        if (log(u3) < 0.5 * z * z + dd - dd * v + dd * log(v)) return boost * dd * v;
    }
    return boost * dd;
}

// inversion below 30, Hoermann's transformed rejection (PTRS, 1993) above
static __device__ double lb_sample_poisson(Draws &d, double lam) {
    if (lam < 30.0) {
        const double t = d.u();
        double p = exp(-lam), c = p, x = 0;
        while (t >= c && x < 1000) { x += 1; p *= lam / x; c += p; }
        return x;
    }
    const double slam = sqrt(lam), loglam = log(lam), b = 0.931 + 2.53 * slam, a = -0.059 + 0.02483 * b;
    const double inv_alpha = 1.1239 + 1.1328 / (b - 3.4), vr = 0.9277 - 3.6224 / (b - 2.0);
    for (int it = 0; it < 100; ++it) {
        const double u = d.u() - 0.5, v = d.uo(), us = 0.5 - fabs(u);
        const double k = floor((2.0 * a / us + b) * u + lam + 0.43);
        if (us >= 0.07 && v <= vr) return k;
        if (k < 0 || (us < 0.013 && v > us)) continue;
        if (log(v) + log(inv_alpha) - log(a / (us * us) + b) <= -lam + k * loglam - lgamma(k + 1.0)) return k;
    }
    return floor(lam);
}

// Group::sample_value of one feature for the group whose statistics start at in / fl (column stride st)
static __device__ uint32_t lb_sample_value(const FeatDev &f, const float *aux, const uint32_t *in, const double *fl,
                                           size_t st, Draws &d) {
    switch (f.type) {
    case LB_BB: {
        const double al = f.hp[0], be = f.hp[1], h = in[0], t = in[st];
        return d.u() < (al + h) / (al + be + h + t);
    }
    case LB_DD:
    case LB_DPD: {
        const float *w0 = aux + f.aux_off;
        const double scale = f.type == LB_DPD ? (double)f.hp[1] : 1.0;
        const int n = f.dim + (f.type == LB_DPD);
        double total = 0;
        for (int v = 0; v < n; ++v) total += v < f.dim ? scale * (double)w0[v] + (double)in[(size_t)v * st] : scale * (double)f.hp[2];
        const double t = d.u() * total;
        double c = 0;
        int last = 0;
        for (int v = 0; v < n; ++v) {
            const double w = v < f.dim ? scale * (double)w0[v] + (double)in[(size_t)v * st] : scale * (double)f.hp[2];
            if (w > 0) last = v;
            c += w;
            if (c > t) { last = v; break; }
        }
        return last == f.dim ? LB_DPD_OTHER : (uint32_t)last;
    }
    case LB_GP: {
        const double shape = (double)f.hp[0] + (double)in[st], rate = (double)f.hp[1] + (double)in[0];
        const double lam = lb_sample_gamma(d, shape) / rate;
        const double x = lb_sample_poisson(d, lam);
        return x >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)x;
    }
    case LB_NICH: {
        const double mu = f.hp[0], kappa = f.hp[1], sigmasq = f.hp[2], nu = f.hp[3];
        const double n = in[0], mean = fl[0], ctv = fl[st];
        const double kn = kappa + n, mun = (kappa * mu + n * mean) / kn, nun = nu + n, dm = mu - mean;
        const double sn = (nu * sigmasq + ctv + n * kappa * dm * dm / kn) / nun;
        const double u = d.uo(), v = d.u();
        const double t = sqrt(nun * (pow(u, -2.0 / nun) - 1.0)) * cospi(2.0 * v);
        return __float_as_uint((float)(mun + sqrt(sn * (kn + 1.0) / kn) * t));
    }
    }
    return 0u;
}

// Group::sample_value: one thread per (sample, requested feature)
__global__ void k_query_values(const KindDev *kinds, const FeatDev *feats, const float *aux, const uint32_t *f2k,
                               const uint32_t *to_sample, int n_to, uint32_t n_samples, int K, uint64_t seed,
                               uint64_t draw_offset, const uint32_t *groups, uint32_t *values) {
    const uint64_t idx = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (idx >= (uint64_t)n_samples * n_to) return;
    const uint32_t smp = (uint32_t)(idx / n_to);
    const uint32_t fid = to_sample[idx % n_to];
    const FeatDev f = feats[fid];
    const uint32_t kid = f2k[fid];
    const KindDev &kd = kinds[kid];
    const uint32_t g = groups[(uint64_t)smp * K + kid];
    const size_t st = (size_t)kd.Gcap;
    Draws d{seed, draw_offset + smp, fid << 10, 0u};
    values[idx] = lb_sample_value(f, aux, kd.ints + (size_t)f.int_row * st + g, kd.flts + (size_t)f.flt_row * st + g, st, d);
}

// generate_rows (src/generate.hpp:36-93): one thread per (feature of the kind, group) walks the group's rows in order --
// observed with probability `density`, value from the group's posterior predictive, added to the group -- and writes
// the cells of its column.  Chains of different (group, feature) pairs touch disjoint statistics and disjoint cells;
// only the mask words are shared between the groups of a column (atomicOr).
__global__ void k_generate(KindDev *kinds, int kid, const FeatDev *feats, const int32_t *kind_feats, const float *aux,
                           uint64_t R, uint64_t W, uint8_t *cat8, uint32_t *wide, uint32_t *mask,
                           const uint32_t *group_ptr, const uint32_t *group_rows, int n_groups, float density, uint64_t seed) {
    KindDev &kd = kinds[kid];
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= kd.nf * n_groups) return;
    const int j = idx / n_groups, g = idx % n_groups;
    const int fid = kind_feats[kd.feat_begin + j];
    const FeatDev f = feats[fid];
    const int st = kd.Gcap;
    uint32_t *in = kd.ints + (size_t)f.int_row * st + g;
    double *fl = kd.flts + (size_t)f.flt_row * st + g;
    float *c = kd.cache + (size_t)f.cache_row * st + g;
    for (uint32_t i = group_ptr[g]; i < group_ptr[g + 1]; ++i) {
        const uint64_t r = group_rows[i];
        Draws d{seed, r, (uint32_t)fid << 10, 0u, 8u};
        if (!(d.u() < (double)density)) continue;
        const uint32_t raw = lb_sample_value(f, aux, in, fl, (size_t)st, d);
        if (f.type == LB_DPD && raw == LB_DPD_OTHER) continue;
        lb_update_cell(f, aux, in, fl, c, st, raw, +1);
        if (f.is_wide) wide[(uint64_t)f.col * R + r] = raw;
        else cat8[(uint64_t)f.col * R + r] = (uint8_t)raw;
        atomicOr(&mask[(uint64_t)fid * W + (r >> 5)], 1u << (r & 31));
    }
}

extern "C" int lb_query_sample(lb_handle h, uint64_t row, const uint32_t *to_sample, int32_t n_to, uint32_t n_samples,
                               uint64_t seed, uint64_t draw_offset, uint32_t *out_values, uint32_t *out_groups) {
    Engine *e = (Engine *)h;
    LB_CUDA(cudaSetDevice(e->device));
    if (row >= e->R) return lb_fail("query_sample: bad row");
    if (n_to < 0 || (n_to && !to_sample)) return lb_fail("query_sample: bad to_sample");
    for (int j = 0; j < n_to; ++j) {
        if (to_sample[j] >= (uint32_t)e->F) return lb_fail("query_sample: bad feature id %u", to_sample[j]);
        if (j && to_sample[j] <= to_sample[j - 1]) return lb_fail("query_sample: to_sample must be ascending");
    }
    const int K = (int)e->kinds.size();
    if (out_groups) std::fill(out_groups, out_groups + (size_t)n_samples * K, LB_UNASSIGNED);
    if (!n_to || !n_samples) return 0;
    LB_TRY(lb_sync_model(e));

    std::vector<QSlot> slots;
    int total_g = 0;
    for (int kid = 0; kid < K; ++kid) {
        bool wanted = false;
        for (int j = 0; j < n_to; ++j) wanted = wanted || (int)e->f2k[to_sample[j]] == kid;
        if (!wanted) continue;                                  // query_server.cc:157-159
        slots.push_back(QSlot{kid, e->kinds[kid].G, total_g});
        total_g += (e->kinds[kid].G + 1) & ~1;                  // keep the doubles 8-byte aligned per slot
    }
    const int n_slots = (int)slots.size();
    auto up = [](size_t x) { return (x + 255) / 256 * 256; };
    const size_t o_cdf = 0, o_scores = o_cdf + up(8 * (size_t)total_g), o_slots = o_scores + up(4 * (size_t)total_g),
                 o_to = o_slots + up(sizeof(QSlot) * slots.size()), o_f2k = o_to + up(4 * (size_t)n_to),
                 o_groups = o_f2k + up(4 * (size_t)e->F), o_values = o_groups + up(4 * (size_t)n_samples * K),
                 bytes = o_values + up(4 * (size_t)n_samples * n_to);
    LB_TRY(lb_ensure_scratch(e, bytes));
    char *base = (char *)e->d_scratch;
    double *d_cdf = (double *)(base + o_cdf);
    float *d_scores = (float *)(base + o_scores);
    QSlot *d_slots = (QSlot *)(base + o_slots);
    uint32_t *d_to = (uint32_t *)(base + o_to), *d_f2k = (uint32_t *)(base + o_f2k);
    uint32_t *d_groups = (uint32_t *)(base + o_groups), *d_values = (uint32_t *)(base + o_values);
    LB_CUDA(cudaMemcpyAsync(d_slots, slots.data(), sizeof(QSlot) * slots.size(), cudaMemcpyHostToDevice, e->stream));
    LB_CUDA(cudaMemcpyAsync(d_to, to_sample, 4 * (size_t)n_to, cudaMemcpyHostToDevice, e->stream));
    LB_CUDA(cudaMemcpyAsync(d_f2k, e->f2k.data(), 4 * (size_t)e->F, cudaMemcpyHostToDevice, e->stream));
    LB_CUDA(cudaMemsetAsync(d_groups, 0xFF, 4 * (size_t)n_samples * K, e->stream));
    for (const QSlot &s : slots) LB_TRY(lb_launch_score_rows(e, s.kind, row, row + 1, d_scores + s.off));
    k_query_cdf<<<n_slots, 128, 0, e->stream>>>(d_slots, d_scores, d_cdf);
    LB_CUDA(cudaGetLastError());
    const uint64_t n_pick = (uint64_t)n_samples * n_slots, n_val = (uint64_t)n_samples * n_to;
    k_query_pick<<<(unsigned)((n_pick + 127) / 128), 128, 0, e->stream>>>(d_slots, n_slots, d_cdf, n_samples, K, seed,
                                                                          draw_offset, d_groups);
    LB_CUDA(cudaGetLastError());
    k_query_values<<<(unsigned)((n_val + 127) / 128), 128, 0, e->stream>>>(e->d_kinds, e->d_feats, e->d_aux, d_f2k, d_to, n_to,
                                                                           n_samples, K, seed, draw_offset, d_groups, d_values);
    LB_CUDA(cudaGetLastError());
    e->launches[LB_L_TOTAL] += 3; e->launches[LB_L_MISC] += 3;
    LB_CUDA(cudaMemcpyAsync(out_values, d_values, 4 * (size_t)n_samples * n_to, cudaMemcpyDeviceToHost, e->stream));
    if (out_groups)
        LB_CUDA(cudaMemcpyAsync(out_groups, d_groups, 4 * (size_t)n_samples * K, cudaMemcpyDeviceToHost, e->stream));
    LB_CUDA(cudaStreamSynchronize(e->stream));
    return 0;
}


extern "C" int lb_generate_rows(lb_handle h, uint64_t R, float density, uint64_t seed) {
    Engine *e = (Engine *)h;
    e->diff_valid = false;
    LB_CUDA(cudaSetDevice(e->device));
    if (!e->F) return lb_fail("generate_rows before set_model");
    if (!(density >= 0.f && density <= 1.f)) return lb_fail("generate_rows: density must be in [0, 1]");
    if (R == 0 || R >= (1ull << 31)) return lb_fail("generate_rows: bad row count");
    const uint64_t W = (R + 31) / 32;
    {   // an all-unobserved block of R rows
        std::vector<uint8_t> cat8((size_t)e->F8 * R + 1, 0);
        std::vector<uint32_t> wide((size_t)e->FW * R + 1, 0), mask((size_t)e->F * W, 0);
        LB_TRY(lb_set_rows(h, R, cat8.data(), wide.data(), mask.data()));
    }
    const int K = (int)e->kinds.size();
    std::vector<uint32_t> pick(R), counts, group_ptr, group_rows(R), cursor;
    std::vector<double> p;
    for (int kid = 0; kid < K; ++kid) {
        // the kind's Pitman-Yor process (Clustering::PitmanYor seating), float64 on the host like the oracle
        const double alpha = e->kinds[kid].alpha, d = e->kinds[kid].d;
        counts.clear();
        int m = 0;
        for (uint64_t r = 0; r < R; ++r) {
            p.resize((size_t)m + 1);
            double total = 0;
            for (int g = 0; g < m; ++g) total += p[g] = counts[g] - d;
            total += p[m] = alpha + d * m;
            const double t = (double)lb_uniform(seed, 7u, r, (uint32_t)kid) * total;
            double c = 0;
            int g = m, last = 0;
            for (int i = 0; i <= m; ++i) {          // distributions::sample_from_likelihoods
                if (p[i] > 0) last = i;
                c += p[i];
                if (c > t) { last = i; break; }
            }
            g = last;
            if (g == m) { counts.push_back(0); ++m; }
            counts[g]++;
            pick[r] = (uint32_t)g;
        }
        LB_TRY(lb_set_groups(h, kid, m, counts.data(), nullptr, nullptr));
        LB_TRY(lb_set_assignments(h, kid, pick.data()));
        const KindHost &k = e->kinds[kid];
        if (k.fids.empty()) continue;
        // rows of every group in order (counting sort of the seating)
        group_ptr.assign((size_t)m + 1, 0);
        for (int g = 0; g < m; ++g) group_ptr[g + 1] = group_ptr[g] + counts[g];
        cursor.assign(group_ptr.begin(), group_ptr.end() - 1);
        for (uint64_t r = 0; r < R; ++r) group_rows[cursor[pick[r]]++] = (uint32_t)r;
        const size_t ptr_bytes = (4 * ((size_t)m + 1) + 255) / 256 * 256;
        LB_TRY(lb_ensure_scratch(e, ptr_bytes + 4 * (size_t)R + 256));
        uint32_t *d_ptr = (uint32_t *)e->d_scratch, *d_rows = (uint32_t *)((char *)e->d_scratch + ptr_bytes);
        LB_CUDA(cudaMemcpyAsync(d_ptr, group_ptr.data(), 4 * ((size_t)m + 1), cudaMemcpyHostToDevice, e->stream));
        LB_CUDA(cudaMemcpyAsync(d_rows, group_rows.data(), 4 * (size_t)R, cudaMemcpyHostToDevice, e->stream));
        LB_TRY(lb_sync_model(e));
        const int n_threads = (int)k.fids.size() * m;
        k_generate<<<(n_threads + 63) / 64, 64, 0, e->stream>>>(e->d_kinds, kid, e->d_feats, e->d_kind_feats, e->d_aux, R, W,
                                                                e->d_cat8, e->d_wide, e->d_mask, d_ptr, d_rows, m, density, seed);
        LB_CUDA(cudaGetLastError());
        e->launches[LB_L_TOTAL]++; e->launches[LB_L_MISC]++;
        LB_TRY(lb_launch_refresh(e, kid));
        LB_CUDA(cudaStreamSynchronize(e->stream));       // the staging vectors are reused by the next kind
    }
    return 0;
}

extern "C" int lb_get_rows(lb_handle h, uint8_t *cat8, uint32_t *wide, uint32_t *mask) {
    Engine *e = (Engine *)h;
    LB_CUDA(cudaSetDevice(e->device));
    if (!e->R) return lb_fail("get_rows: no rows loaded");
    if (e->F8) LB_CUDA(cudaMemcpyAsync(cat8, e->d_cat8, (size_t)e->F8 * e->R, cudaMemcpyDeviceToHost, e->stream));
    if (e->FW) LB_CUDA(cudaMemcpyAsync(wide, e->d_wide, 4 * (size_t)e->FW * e->R, cudaMemcpyDeviceToHost, e->stream));
    LB_CUDA(cudaMemcpyAsync(mask, e->d_mask, 4 * (size_t)e->F * e->W, cudaMemcpyDeviceToHost, e->stream));
    LB_CUDA(cudaStreamSynchronize(e->stream));
    return 0;
}
