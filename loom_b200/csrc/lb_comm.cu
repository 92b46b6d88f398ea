// lb_comm.cu -- NCCL collectives of libloomb200.so (include/loom_b200_comm.h): the row-sharded kind proposer and
// the task-sharded hyper kernel.  One communicator per engine, every collective on the engine's stream.
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include <vector>

#include <dlfcn.h>
#include <nccl.h>

#include "../../include/loom_b200_comm.h"
#include "lb_internal.h"

struct Comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    double *d_tmp = nullptr;     // small device staging (barrier word, timing maxima, hyper exchange)
    size_t tmp_cap = 0;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};

// NCCL is bound at run time, privately (dlopen, RTLD_LOCAL), the first time a communicator is asked for: a process
// that never shards (one GPU, the tests, loom_infer) does not load it at all, and a host program that brings its own
// NCCL build (PyTorch bundles a newer one than the system's) keeps its symbols -- linking libnccl into this library
// made `import torch` fail with an undefined ncclDevCommCreate once libloomb200 had been loaded first.
namespace {
struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    void *handle = nullptr;
    bool ok = false;
};
NcclApi g_nccl;
}  // namespace

static int nccl_load() {
    if (g_nccl.ok) return 0;
    const char *names[] = {getenv("LOOM_B200_NCCL"), "libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        if (!n || !*n) continue;
        g_nccl.handle = dlopen(n, RTLD_NOW | RTLD_LOCAL);
        if (g_nccl.handle) break;
    }
    if (!g_nccl.handle) return lb_fail("NCCL not found (libnccl.so.2; LOOM_B200_NCCL overrides the path): %s", dlerror());
    bool all = true;
    auto bind = [&](auto &fn, const char *sym) {
        fn = (std::remove_reference_t<decltype(fn)>)dlsym(g_nccl.handle, sym);
        all = all && fn != nullptr;
    };
    bind(g_nccl.GetUniqueId, "ncclGetUniqueId"); bind(g_nccl.CommInitRank, "ncclCommInitRank");
    bind(g_nccl.CommDestroy, "ncclCommDestroy"); bind(g_nccl.AllReduce, "ncclAllReduce");
    bind(g_nccl.AllGather, "ncclAllGather"); bind(g_nccl.Broadcast, "ncclBroadcast");
    bind(g_nccl.GroupStart, "ncclGroupStart"); bind(g_nccl.GroupEnd, "ncclGroupEnd");
    bind(g_nccl.GetErrorString, "ncclGetErrorString");
    if (!all) return lb_fail("the NCCL library lacks an entry point this layer needs");
    g_nccl.ok = true;
    return 0;
}
#define ncclGetUniqueId g_nccl.GetUniqueId
#define ncclCommInitRank g_nccl.CommInitRank
#define ncclCommDestroy g_nccl.CommDestroy
#define ncclAllReduce g_nccl.AllReduce
#define ncclAllGather g_nccl.AllGather
#define ncclBroadcast g_nccl.Broadcast
#define ncclGroupStart g_nccl.GroupStart
#define ncclGroupEnd g_nccl.GroupEnd
#define ncclGetErrorString g_nccl.GetErrorString

#define LB_NCCL(expr)                                                                   \
    do {                                                                                \
        ncclResult_t r__ = (expr);                                                      \
        if (r__ != ncclSuccess)                                                         \
            return lb_fail("%s failed: %s (%s:%d)", #expr, ncclGetErrorString(r__),     \
                           __FILE__, __LINE__);                                         \
    } while (0)

static Comm *comm_of(Engine *e) { return (Comm *)e->comm; }

static int comm_tmp(Engine *e, size_t bytes) {
    Comm *c = comm_of(e);
    if (bytes > c->tmp_cap) {
        if (c->d_tmp) cudaFree(c->d_tmp);
        c->d_tmp = nullptr;
        LB_CUDA(cudaMalloc((void **)&c->d_tmp, bytes + 256));
        c->tmp_cap = bytes + 256;
    }
    return 0;
}

void lb_release_comm(Engine *e) {
    Comm *c = comm_of(e);
    if (!c) return;
    if (c->comm) ncclCommDestroy(c->comm);
    if (c->d_tmp) cudaFree(c->d_tmp);
    for (cudaEvent_t ev : c->ev) if (ev) cudaEventDestroy(ev);
    delete c;
    e->comm = nullptr;
}

extern "C" int lb_comm_unique_id(uint8_t *id_out) {
    static_assert(sizeof(ncclUniqueId) == LB_COMM_ID_BYTES, "ncclUniqueId size");
    LB_TRY(nccl_load());
    ncclUniqueId id;
    LB_NCCL(ncclGetUniqueId(&id));
    memcpy(id_out, &id, sizeof id);
    return 0;
}

extern "C" int lb_comm_init(lb_handle h, const uint8_t *id_bytes, int32_t rank, int32_t world) {
    Engine *e = (Engine *)h;
    if (!e) return lb_fail("comm_init: null handle");
    if (world < 1 || rank < 0 || rank >= world) return lb_fail("comm_init: bad rank %d of %d", rank, world);
    LB_CUDA(cudaSetDevice(e->device));
    LB_TRY(nccl_load());
    lb_release_comm(e);
    Comm *c = new Comm;
    c->rank = rank; c->world = world;
    e->comm = c;
    ncclUniqueId id;
    memcpy(&id, id_bytes, sizeof id);
    LB_NCCL(ncclCommInitRank(&c->comm, world, id, rank));
    for (cudaEvent_t &ev : c->ev) LB_CUDA(cudaEventCreate(&ev));
    return comm_tmp(e, 4096);
}

extern "C" int lb_comm_destroy(lb_handle h) {
    Engine *e = (Engine *)h;
    if (e) { cudaSetDevice(e->device); cudaStreamSynchronize(e->stream); lb_release_comm(e); }
    return 0;
}

extern "C" int lb_comm_rank(lb_handle h, int32_t *rank, int32_t *world) {
    Engine *e = (Engine *)h;
    Comm *c = e ? comm_of(e) : nullptr;
    if (rank) *rank = c ? c->rank : 0;
    if (world) *world = c ? c->world : 1;
    return 0;
}

extern "C" int lb_comm_barrier(lb_handle h) {
    Engine *e = (Engine *)h;
    LB_CUDA(cudaSetDevice(e->device));
    Comm *c = comm_of(e);
    if (c && c->world > 1) {
        LB_CUDA(cudaMemsetAsync(c->d_tmp, 0, 8, e->stream));
        LB_NCCL(ncclAllReduce(c->d_tmp, c->d_tmp, 1, ncclFloat64, ncclSum, c->comm, e->stream));
    }
    LB_CUDA(cudaStreamSynchronize(e->stream));
    return 0;
}

extern "C" int lb_comm_max_f64(lb_handle h, double *inout, int32_t n) {
    Engine *e = (Engine *)h;
    LB_CUDA(cudaSetDevice(e->device));
    Comm *c = comm_of(e);
    if (!c || c->world == 1 || n <= 0) return 0;
    LB_TRY(comm_tmp(e, 8 * (size_t)n));
    LB_CUDA(cudaMemcpyAsync(c->d_tmp, inout, 8 * (size_t)n, cudaMemcpyHostToDevice, e->stream));
    LB_NCCL(ncclAllReduce(c->d_tmp, c->d_tmp, (size_t)n, ncclFloat64, ncclMax, c->comm, e->stream));
    LB_CUDA(cudaMemcpyAsync(inout, c->d_tmp, 8 * (size_t)n, cudaMemcpyDeviceToHost, e->stream));
    LB_CUDA(cudaStreamSynchronize(e->stream));
    return 0;
}

extern "C" int lb_comm_kind_proposer_score(lb_handle h, float *out_scores, lb_comm_round_stats *stats) {
    Engine *e = (Engine *)h;
    LB_CUDA(cudaSetDevice(e->device));
    Comm *c = comm_of(e);
    const int world = c ? c->world : 1, rank = c ? c->rank : 0;
    const int F = e->F, K = (int)e->kinds.size();
    const uint64_t lo = e->R * (uint64_t)rank / world, hi = e->R * (uint64_t)(rank + 1) / world;
    const int fb = (F + world - 1) / world;                         // feature block of every rank
    const int f_lo = std::min(F, rank * fb), f_hi = std::min(F, (rank + 1) * fb);
    cudaEvent_t local[6];
    cudaEvent_t *ev = c ? c->ev : local;
    if (!c) for (cudaEvent_t &x : local) LB_CUDA(cudaEventCreate(&x));
    LB_CUDA(cudaEventRecord(ev[0], e->stream));
    LB_TRY(lb_proposer_accumulate(h, lo, hi));                      // K4 over this rank's rows, zeroed slabs first
    LB_CUDA(cudaEventRecord(ev[1], e->stream));
    if (world > 1) {
        // counts are uint32 sums (exact in any order); float64 slabs hold sum x, sum x^2, sum log x!
        if (e->prop_used_ints)
            LB_NCCL(ncclAllReduce(e->d_prop_ints, e->d_prop_ints, e->prop_used_ints, ncclUint32, ncclSum, c->comm, e->stream));
        if (e->prop_used_flts)
            LB_NCCL(ncclAllReduce(e->d_prop_flts, e->d_prop_flts, e->prop_used_flts, ncclFloat64, ncclSum, c->comm, e->stream));
    }
    LB_CUDA(cudaEventRecord(ev[2], e->stream));
    LB_TRY(lb_proposer_score_range(e, f_lo, f_hi, fb * world));     // K5 over this rank's features
    LB_CUDA(cudaEventRecord(ev[3], e->stream));
    if (world > 1)
        LB_NCCL(ncclAllGather(e->d_prop_scores + (size_t)rank * fb * K, e->d_prop_scores, (size_t)fb * K, ncclFloat32,
                              c->comm, e->stream));
    LB_CUDA(cudaEventRecord(ev[4], e->stream));
    LB_TRY(lb_proposer_fetch_scores(e, out_scores));                // synchronises the stream
    if (stats) {
        float a = 0, b = 0, s = 0, g = 0;
        LB_CUDA(cudaEventElapsedTime(&a, ev[0], ev[1]));
        LB_CUDA(cudaEventElapsedTime(&b, ev[1], ev[2]));
        LB_CUDA(cudaEventElapsedTime(&s, ev[2], ev[3]));
        LB_CUDA(cudaEventElapsedTime(&g, ev[3], ev[4]));
        stats->accumulate_ms = a; stats->allreduce_ms = b; stats->score_ms = s; stats->allgather_ms = g;
        stats->allreduce_bytes = world > 1 ? 4.0 * (double)e->prop_used_ints + 8.0 * (double)e->prop_used_flts : 0.0;
        stats->allgather_bytes = world > 1 ? 4.0 * (double)fb * K * world : 0.0;
        stats->row_begin = (double)lo; stats->row_end = (double)hi;
        stats->feature_begin = f_lo; stats->feature_end = f_hi;
    }
    if (!c) for (cudaEvent_t x : local) cudaEventDestroy(x);
    return 0;
}

extern "C" int lb_comm_hyper_run(lb_handle h, uint64_t seed, double *allreduce_ms) {
    Engine *e = (Engine *)h;
    LB_CUDA(cudaSetDevice(e->device));
    Comm *c = comm_of(e);
    if (allreduce_ms) *allreduce_ms = 0.0;
    if (!c || c->world == 1) return lb_hyper_run(h, seed);
    const int F = e->F, K = (int)e->kinds.size(), T = 1 + K + F;
    std::vector<uint8_t> mask(T);
    for (int t = 0; t < T; ++t) mask[t] = (uint8_t)(t % c->world == c->rank);
    LB_TRY(lb_hyper_run_tasks(h, seed, mask.data()));
    // exchange: [topology 2][kind alpha, d 2K][hp 4F][aux]; owners contribute their choice, everyone else zero
    const size_t A = e->aux.size(), n = 2 + 2 * (size_t)K + 4 * (size_t)F + A;
    std::vector<float> buf(n, 0.f);
    if (mask[0]) { buf[0] = e->topo[0]; buf[1] = e->topo[1]; }
    for (int k = 0; k < K; ++k)
        if (mask[1 + k]) { buf[2 + 2 * k] = e->kinds[k].alpha; buf[3 + 2 * k] = e->kinds[k].d; }
    float *hp = buf.data() + 2 + 2 * (size_t)K, *aux = hp + 4 * (size_t)F;
    for (int f = 0; f < F; ++f) {
        if (!mask[1 + K + f]) continue;
        const lb_feature_spec &s = e->specs[f];
        for (int q = 0; q < 4; ++q) hp[4 * (size_t)f + q] = s.hp[q];
        if (s.type == LB_DD || s.type == LB_DPD)
            for (int v = 0; v < s.dim; ++v) aux[s.aux_off + v] = e->aux[s.aux_off + v];
    }
    LB_TRY(comm_tmp(e, 4 * n));
    float *d = (float *)c->d_tmp;
    LB_CUDA(cudaMemcpyAsync(d, buf.data(), 4 * n, cudaMemcpyHostToDevice, e->stream));
    LB_CUDA(cudaEventRecord(c->ev[0], e->stream));
    LB_NCCL(ncclAllReduce(d, d, n, ncclFloat32, ncclSum, c->comm, e->stream));
    LB_CUDA(cudaEventRecord(c->ev[1], e->stream));
    LB_CUDA(cudaMemcpyAsync(buf.data(), d, 4 * n, cudaMemcpyDeviceToHost, e->stream));
    LB_CUDA(cudaStreamSynchronize(e->stream));
    if (allreduce_ms) { float ms = 0; LB_CUDA(cudaEventElapsedTime(&ms, c->ev[0], c->ev[1])); *allreduce_ms = ms; }
    std::vector<lb_feature_spec> specs(e->specs);
    for (int f = 0; f < F; ++f)
        for (int q = 0; q < 4; ++q) specs[f].hp[q] = hp[4 * (size_t)f + q];
    // features without a grid task (no grid given for their type) contributed zeros from every rank: keep theirs
    // -- a task that ran always owns its entries, so "all zero" can only mean "nobody ran it"
    for (int f = 0; f < F; ++f) {
        bool any = false;
        for (int q = 0; q < 4; ++q) any = any || hp[4 * (size_t)f + q] != 0.f;
        if (!any) for (int q = 0; q < 4; ++q) specs[f].hp[q] = e->specs[f].hp[q];
    }
    std::vector<float> new_aux(aux, aux + A);
    std::vector<float> kind_py(2 * (size_t)K);
    for (int k = 0; k < K; ++k) { kind_py[2 * k] = buf[2 + 2 * k]; kind_py[2 * k + 1] = buf[3 + 2 * k]; }
    float topo[2] = {buf[0], buf[1]};
    return lb_set_hypers(h, specs.data(), A ? new_aux.data() : nullptr, kind_py.data(), topo);
}


// ---------------------------------------------------------------------------------------------------------------
// Kind-sharded cat sweep: "tasks in different kinds are independent" (doc/adapting.md:144-147),
// so rank owner[k] sweeps kind k (lb_cat_sweep with a kind mask) and nobody talks during the sweep.  Before a batch
// kernel that needs every kind's partition (the row-sharded proposer, the hyper kernel, a dump) the owners publish
// their kinds: one all-reduce of the kinds' headers (G, capacity, id-tracker extent, sample size), then ONE NCCL group
// of broadcasts -- per kind: clustering counts, id maps, the three group tables and the assignment column.
extern "C" int lb_comm_sync_kinds(lb_handle h, const int32_t *owner, double *bytes_out, double *ms_out) {
    Engine *e = (Engine *)h;
    LB_CUDA(cudaSetDevice(e->device));
    Comm *c = comm_of(e);
    if (bytes_out) *bytes_out = 0.0;
    if (ms_out) *ms_out = 0.0;
    if (!c || c->world == 1) return 0;
    const int K = (int)e->kinds.size();
    for (int k = 0; k < K; ++k)
        if (owner[k] < 0 || owner[k] >= c->world) return lb_fail("comm_sync_kinds: kind %d has owner %d of %d", k, owner[k], c->world);
    LB_TRY(lb_sync_model(e));
    LB_TRY(lb_pull_kinds(e));
    std::vector<unsigned long long> hdr(5 * (size_t)K, 0ull);
    for (int k = 0; k < K; ++k) {
        if (owner[k] != c->rank) continue;
        const KindHost &kh = e->kinds[k];
        unsigned long long *x = &hdr[5 * (size_t)k];
        x[0] = (unsigned long long)kh.G; x[1] = (unsigned long long)kh.Gcap; x[2] = kh.g2p_cap;
        x[3] = kh.next_global; x[4] = kh.sample_size;
    }
    LB_TRY(comm_tmp(e, 8 * hdr.size()));
    LB_CUDA(cudaMemcpyAsync(c->d_tmp, hdr.data(), 8 * hdr.size(), cudaMemcpyHostToDevice, e->stream));
    LB_NCCL(ncclAllReduce(c->d_tmp, c->d_tmp, hdr.size(), ncclUint64, ncclSum, c->comm, e->stream));
    LB_CUDA(cudaMemcpyAsync(hdr.data(), c->d_tmp, 8 * hdr.size(), cudaMemcpyDeviceToHost, e->stream));
    LB_CUDA(cudaStreamSynchronize(e->stream));
    for (int k = 0; k < K; ++k) {               // receivers take the owner's table shape
        KindHost &kh = e->kinds[k];
        const unsigned long long *x = &hdr[5 * (size_t)k];
        if (owner[k] != c->rank) {
            if (kh.Gcap != (int)x[1]) LB_TRY(lb_kind_alloc(e, kh, (int)x[0], (int)x[1]));
            if (kh.g2p_cap != (uint32_t)x[2]) {
                if (kh.g2p) LB_CUDA(cudaFreeAsync(kh.g2p, e->stream));
                kh.g2p = nullptr;
                LB_CUDA(cudaMallocAsync((void **)&kh.g2p, 4 * (size_t)x[2], e->stream));
                kh.g2p_cap = (uint32_t)x[2];
            }
            kh.G = (int)x[0]; kh.next_global = (uint32_t)x[3]; kh.sample_size = x[4];
        }
    }
    double bytes = 0.0;
    LB_CUDA(cudaEventRecord(c->ev[0], e->stream));
    LB_NCCL(ncclGroupStart());
    for (int k = 0; k < K; ++k) {
        KindHost &kh = e->kinds[k];
        const size_t cap = (size_t)kh.Gcap;
        struct { void *p; size_t n; } bufs[] = {
            {kh.counts, 4 * cap}, {kh.shifted, 4 * cap}, {kh.p2g, 4 * cap}, {kh.g2p, 4 * (size_t)kh.g2p_cap},
            {kh.ints, 4 * (size_t)kh.n_int_rows * cap}, {kh.flts, 8 * (size_t)kh.n_flt_rows * cap},
            {kh.cache, 4 * (size_t)(kh.n_cache_rows + 1) * cap}, {kh.assign, 4 * (size_t)e->R}};
        for (auto &b : bufs) {
            if (!b.p || !b.n) continue;
            LB_NCCL(ncclBroadcast(b.p, b.p, b.n, ncclUint8, owner[k], c->comm, e->stream));
            bytes += (double)b.n;
        }
    }
    LB_NCCL(ncclGroupEnd());
    LB_CUDA(cudaEventRecord(c->ev[1], e->stream));
    LB_TRY(lb_sync_model(e));                   // the KindDev mirrors follow the host fields (synchronises)
    lb_free_proposer(e);
    if (bytes_out) *bytes_out = bytes;
    if (ms_out) { float ms = 0; LB_CUDA(cudaEventElapsedTime(&ms, c->ev[0], c->ev[1])); *ms_out = ms; }
    return 0;
}


// ---------------------------------------------------------------------------------------------------------------
// Row-sharded bulk accumulate (K2; ProductMixture::add_value over all rows with given assignments,
// src/product_mixture.cc:165-184).  Every rank holds the same groups and the assignments of ALL rows; it zeroes the
// statistics, adds rows [R r / N, R (r + 1) / N), and the tables are summed over the ranks: clustering counts and the
// integer statistics as uint32 (exact in any order), the float statistics as float64 SUMS (GP sum log x!, NICH sum x and
// sum x^2 -- the moments are formed after the reduce, as in the proposer).  One NCCL group per kind.  Every rank ends
// with the statistics a single-device lb_accumulate_rows over all rows leaves.
extern "C" int lb_comm_accumulate_rows(lb_handle h, double *allreduce_ms, double *allreduce_bytes) {
    Engine *e = (Engine *)h;
    LB_CUDA(cudaSetDevice(e->device));
    Comm *c = comm_of(e);
    const int world = c ? c->world : 1, rank = c ? c->rank : 0;
    if (allreduce_ms) *allreduce_ms = 0.0;
    if (allreduce_bytes) *allreduce_bytes = 0.0;
    LB_TRY(lb_sync_model(e));
    LB_TRY(lb_pull_kinds(e));
    const int K = (int)e->kinds.size();
    const uint64_t lo = e->R * (uint64_t)rank / world, hi = e->R * (uint64_t)(rank + 1) / world;
    for (int kid = 0; kid < K; ++kid) {
        KindHost &k = e->kinds[kid];
        const size_t cap = (size_t)k.Gcap;
        LB_CUDA(cudaMemsetAsync(k.counts, 0, 4 * cap, e->stream));
        if (k.n_int_rows) LB_CUDA(cudaMemsetAsync(k.ints, 0, 4 * (size_t)k.n_int_rows * cap, e->stream));
        if (k.n_flt_rows) LB_CUDA(cudaMemsetAsync(k.flts, 0, 8 * (size_t)k.n_flt_rows * cap, e->stream));
        k.sample_size = 0;
    }
    LB_TRY(lb_sync_model(e));
    double bytes = 0.0;
    for (int kid = 0; kid < K; ++kid) LB_TRY(lb_launch_accumulate(e, kid, lo, hi, +1, /*sums_only=*/true));
    if (world > 1) {
        LB_CUDA(cudaEventRecord(c->ev[0], e->stream));
        LB_NCCL(ncclGroupStart());
        for (int kid = 0; kid < K; ++kid) {
            KindHost &k = e->kinds[kid];
            const size_t cap = (size_t)k.Gcap;
            LB_NCCL(ncclAllReduce(k.counts, k.counts, cap, ncclUint32, ncclSum, c->comm, e->stream));
            bytes += 4.0 * cap;
            if (k.n_int_rows) { LB_NCCL(ncclAllReduce(k.ints, k.ints, (size_t)k.n_int_rows * cap, ncclUint32, ncclSum, c->comm, e->stream)); bytes += 4.0 * k.n_int_rows * cap; }
            if (k.n_flt_rows) { LB_NCCL(ncclAllReduce(k.flts, k.flts, (size_t)k.n_flt_rows * cap, ncclFloat64, ncclSum, c->comm, e->stream)); bytes += 8.0 * k.n_flt_rows * cap; }
        }
        LB_NCCL(ncclGroupEnd());
        LB_CUDA(cudaEventRecord(c->ev[1], e->stream));
    }
    // sample sizes follow the summed counts
    for (int kid = 0; kid < K; ++kid) {
        KindHost &k = e->kinds[kid];
        std::vector<uint32_t> counts((size_t)std::max(k.G, 1));
        LB_CUDA(cudaMemcpyAsync(counts.data(), k.counts, 4 * (size_t)k.G, cudaMemcpyDeviceToHost, e->stream));
        LB_CUDA(cudaStreamSynchronize(e->stream));
        uint64_t n = 0;
        for (int g = 0; g < k.G; ++g) n += counts[g];
        k.sample_size = n;
    }
    LB_TRY(lb_sync_model(e));
    for (int kid = 0; kid < K; ++kid) LB_TRY(lb_launch_sums_to_moments(e, kid));
    LB_CUDA(cudaStreamSynchronize(e->stream));
    lb_free_proposer(e);
    if (world > 1 && allreduce_ms) { float ms = 0; LB_CUDA(cudaEventElapsedTime(&ms, c->ev[0], c->ev[1])); *allreduce_ms = ms; }
    if (allreduce_bytes) *allreduce_bytes = bytes;
    return 0;
}
