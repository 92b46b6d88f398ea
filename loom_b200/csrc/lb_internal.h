// lb_internal.h -- host-side engine state of libloomb200.so.
#pragma once
#include <string>
#include <vector>

#include "lb_common.cuh"

struct KindHost {
    std::vector<int> fids;  // global feature ids, ascending
    float alpha = 2.f, d = 0.1f;
    int G = 0, Gcap = 0;
    uint64_t sample_size = 0;
    uint32_t next_global = 0, g2p_cap = 0;
    int n_int_rows = 0, n_flt_rows = 0, n_cache_rows = 0;
    std::vector<int> int_off, flt_off, cache_off;  // per kind-local feature
    uint32_t *counts = nullptr, *p2g = nullptr, *g2p = nullptr, *ints = nullptr, *assign = nullptr;
    float *shifted = nullptr, *cache = nullptr;
    double *flts = nullptr;
    int32_t *cache_row_feat = nullptr;
    bool assign_pooled = false;   // assign came from the stream-ordered pool (ephemeral kinds: no device-wide sync to make or drop one)
};

struct ProposerKind {      // all-feature tables of one kind (KindProposer::Kind): views into the slabs
    int G = 0, pst = 0;        // groups, column stride
    uint32_t *ints = nullptr;  // [prop_int_rows][pst]
    double *flts = nullptr;    // [prop_flt_rows][pst]
    void *packed = nullptr;    // [R] packed group id of every row in this kind (1, 2 or 4 bytes each)
};

// launch descriptors of the bulk accumulate kernels (lb_accum.cuh)
struct AccTarget {             // packed ids of every row + the [rows][st] tables the statistics go to
    const void *packed;
    uint32_t *ints;
    double *flts;
    int st, G, width;
};
struct KindDev;
struct SweepChain {            // one chain's part of a sweep launch (k_cat_sweep, blockIdx.y)
    KindDev *kinds;
    const int32_t *kind_ids;
    int n_ids;
    const FeatDev *feats;
    const int32_t *kind_feats;
    const float *aux;
    RowsDev rows;
    const uint32_t *actions;
    uint64_t n_actions, seed, offset;
    int K, n_empty;
    uint32_t *dbg_picks;
    float *dbg_scores;
    int dbg_stride;
    unsigned long long *prof;
};
struct PackSrc {               // out[r] = g2p[assign[r]] narrowed to `width` bytes
    const uint32_t *assign, *g2p;
    void *out;
    int width;
};

struct Engine {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr, stream3 = nullptr, stream4 = nullptr;   // K3s / K3f: the classes of a sweep run side by side
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_join3 = nullptr, ev_join4 = nullptr;
    cudaEvent_t cls_ev[4][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};   // LB_SWEEP_TIMING=1: per-class kernel times
    int cls_pairs[4] = {0, 0, 0, 0};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int n_sm = 148;

    int F = 0, F8 = 0, FW = 0;
    std::vector<lb_feature_spec> specs;
    std::vector<float> aux;
    std::vector<uint32_t> f2k;
    float topo[2] = {2.f, 0.1f};
    int empty_group_count = 1;
    std::vector<KindHost> kinds;

    // device mirrors of the model
    FeatDev *d_feats = nullptr;    // [F] global feature order
    int32_t *d_kind_feats = nullptr;  // kind-major list of global feature ids
    float *d_aux = nullptr;
    KindDev *d_kinds = nullptr;
    int d_kinds_cap = 0;
    std::vector<KindDev> h_kinds;
    KindDev *pinned_kinds = nullptr;   // pinned staging for the device->host status pull
    int pinned_kinds_cap = 0;
    bool model_dirty = true;

    // rows
    uint64_t R = 0, W = 0;
    bool rows_borrowed = false;        // row block belongs to another engine (lb_share_rows)
    Engine *rows_owner = nullptr;      // ... this one; it lists us in `borrowers` until we let go
    std::vector<Engine *> borrowers;   // engines reading OUR row block: detached before the block is freed or replaced
    uint8_t *d_cat8 = nullptr;
    uint32_t *d_wide = nullptr, *d_mask = nullptr;

    // scratch
    void *d_scratch = nullptr;
    size_t scratch_bytes = 0;
    uint32_t *d_actions = nullptr;
    size_t actions_cap = 0;
    int32_t *d_flag = nullptr;
    int32_t *d_kind_ids = nullptr;
    int32_t *d_fid_list = nullptr;     // K2/K4: feature id lists per shared-memory class
    int d_fid_list_cap = 0;
    void *d_score_aux = nullptr;       // K1: GP feature index + predictive tables
    size_t score_aux_bytes = 0;
    // the tare-compressed form of the loaded block, kept on the device by lb_set_rows_diff for the sparse scorer (K1d)
    void *d_diff = nullptr;            // [tare_obs F][tare_val F][row_has_tare R?][pos_ptr R+1][pos_feature][pos_value][neg_ptr R+1][neg_feature]
    size_t diff_bytes = 0, diff_off[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    bool diff_valid = false, diff_has_row_flags = false;
    // the pos cells of TARED columns only (the untared ones are scored densely): built on first use by K1d
    uint64_t *d_tpos_ptr = nullptr;
    uint32_t *d_tpos_feature = nullptr, *d_tpos_value = nullptr;
    size_t tpos_cap_rows = 0, tpos_cap = 0;
    bool tpos_valid = false;
    void *d_tc_aux = nullptr;          // K1t: producer records, chunk lists and the bf16 operand planes
    size_t tc_aux_bytes = 0;
    int d_kind_ids_cap = 0;

    // hyper prior
    std::vector<float> hp[13];

    // kind proposer
    std::vector<ProposerKind> prop;
    uint32_t *d_prop_ints = nullptr;   // slab holding every kind's integer tables
    double *d_prop_flts = nullptr;
    unsigned char *d_prop_packed = nullptr;   // [K][R] packed ids, 4 bytes per row reserved
    size_t prop_cap_ints = 0, prop_cap_flts = 0, prop_cap_packed = 0;
    size_t prop_used_ints = 0, prop_used_flts = 0;   // elements of the slabs the current round uses (all-reduce extent)
    void *d_desc = nullptr;            // launch descriptor staging
    size_t d_desc_cap = 0;
    std::vector<int> prop_int_off, prop_flt_off;
    int prop_int_rows = 0, prop_flt_rows = 0;
    std::vector<float> prop_scores;  // [F][K]
    float *d_prop_scores = nullptr;
    size_t prop_scores_cap = 0;
    bool prop_valid = false;
    bool prop_accumulated = false;     // tables hold (possibly partial) sums, not yet finished
    FeatDev *d_prop_feats = nullptr;   // all-feature row layout of the proposer tables

    uint64_t launches[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    double kernel_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    cudaEvent_t kev0 = nullptr, kev1 = nullptr;
    bool kev_pending = false;
    unsigned long long *d_prof = nullptr;  // LB_PROFILE=1: per-kind, per-warp phase cycles of the sweep
    void *differ = nullptr;                // tare / sparsify state (lb_differ.cu)
    void *comm = nullptr;                  // NCCL communicator + rank (lb_comm.cu)
    void *seating = nullptr;               // ephemeral-kind seating drawn ahead of time (lb_struct.cu)
};

// launch-family indices for launch_count
enum { LB_L_TOTAL = 0, LB_L_SWEEP = 1, LB_L_SCORE = 2, LB_L_ACCUM = 3, LB_L_REFRESH = 4,
       LB_L_KIND = 5, LB_L_HYPER = 6, LB_L_MISC = 7 };

int lb_fail(const char *fmt, ...);
#define LB_CUDA(expr)                                                                   \
    do {                                                                                \
        cudaError_t err__ = (expr);                                                     \
        if (err__ != cudaSuccess)                                                       \
            return lb_fail("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__),   \
                           __FILE__, __LINE__);                                         \
    } while (0)
#define LB_TRY(expr) do { int rc__ = (expr); if (rc__) return rc__; } while (0)

inline int spec_int_rows(const lb_feature_spec &s) {
    switch (s.type) {
    case LB_BB: return 2;
    case LB_DD: case LB_DPD: return s.dim + 1;
    case LB_GP: return 2;
    case LB_NICH: return 1;
    }
    return 0;
}
inline int spec_flt_rows(const lb_feature_spec &s) {
    return s.type == LB_GP ? 1 : (s.type == LB_NICH ? 2 : 0);
}
inline int spec_cache_rows(const lb_feature_spec &s) {
    switch (s.type) {
    case LB_BB: return 2;
    case LB_DD: case LB_DPD: return s.dim + 1;
    case LB_GP: return 2;
    case LB_NICH: return 5;
    }
    return 0;
}

// engine helpers (lb_engine.cu)
int lb_sync_model(Engine *e);                 // upload FeatDev / KindDev if dirty
int lb_pull_kinds(Engine *e);                 // KindDev -> KindHost dynamic fields
int lb_kind_alloc(Engine *e, KindHost &k, int G, int Gcap);
void lb_kind_release(Engine *e, KindHost &k);
int lb_kind_layout(Engine *e, KindHost &k, const std::vector<int> &fids);
int lb_kind_grow(Engine *e, int kid, int new_cap);
int lb_refresh_kind(Engine *e, int kid);      // rebuild all cached scorers of a kind
int lb_ensure_scratch(Engine *e, size_t bytes);

// kernels (lb_cat.cu)
int lb_sweep_chain(Engine *e, const std::vector<int> &kind_ids, const uint32_t *d_actions, uint64_t n_actions,
                   uint64_t seed, uint64_t offset, uint32_t *d_picks, float *d_scores, int stride, SweepChain *out);
int lb_launch_sweep_chains(Engine *leader, const SweepChain *chains, int n, const std::vector<Engine *> &engines,
                           const std::vector<std::vector<int>> &kind_ids);
int lb_launch_sweep(Engine *e, const std::vector<int> &kind_ids, uint64_t n_actions, uint64_t seed,
                    uint64_t offset, uint32_t *d_picks, float *d_scores, int stride);
int lb_launch_row_lse_add(Engine *e, const float *d_scores, uint64_t n_rows, int G, float *d_acc, int first);
int lb_launch_score_rows(Engine *e, int kid, uint64_t b, uint64_t en, float *d_out, int lse = 0);
int lb_launch_score_rows_tc(Engine *e, int kid, uint64_t b, uint64_t en, const float *d_tab, const int32_t *row_base,
                            int K_total, float *d_out, int planes, int *used);   // lb_score_tc.cu
int lb_launch_score_rows_diff(Engine *e, int kid, uint64_t b, uint64_t en, float *d_out);
int lb_launch_accumulate(Engine *e, int kid, uint64_t b, uint64_t en, int sign, bool sums_only = false);
int lb_launch_sums_to_moments(Engine *e, int kid);
int lb_launch_accum_features(Engine *e, const FeatDev *d_feats, const std::vector<int> &fids,
                             const std::vector<AccTarget> &targets, uint64_t b, uint64_t en, int sign,
                             int launch_family);
int lb_packed_width(int G);
int lb_launch_pack_assign(Engine *e, const std::vector<PackSrc> &src, uint64_t b, uint64_t en, int launch_family);
int lb_launch_refresh(Engine *e, int kid);
int lb_launch_restride(Engine *e, const KindHost &src, KindHost &dst);
int lb_launch_renumber(Engine *e, int kid);
int lb_launch_validate_rows(Engine *e, int *bad);
int lb_launch_score_data(Engine *e, double *out);

// tare / sparsify (lb_differ.cu)
void lb_release_differ(Engine *e);

// structure kernels (lb_struct.cu)
int lb_proposer_score_range(Engine *e, int f_lo, int f_hi, int F_pad);
int lb_proposer_fetch_scores(Engine *e, float *out);
// collectives (lb_comm.cu)
void lb_release_comm(Engine *e);
void lb_release_seating(Engine *e);
void lb_free_proposer(Engine *e);      // invalidate (buffers kept)
void lb_release_proposer(Engine *e);   // free the buffers
