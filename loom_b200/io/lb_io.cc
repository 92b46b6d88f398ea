// lb_io.cc -- loom's protobuf files <-> the arrays of include/loom_b200.h.
// C ABI in include/loom_b200_io.h; host code only (g++ + zlib), no protobuf runtime.
//
// Field numbers of the protobuf.loom messages are those of src/schema.proto (cited at each
// codec).  The distributions messages (namespace dist below) are a stand-in, see the header.
#include "../../include/loom_b200_io.h"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <future>
#include <string>
#include <sys/time.h>
#include <sys/resource.h>
#include <unordered_map>
#include <utility>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "pbwire.hpp"

using pbw::Reader;
using pbw::Writer;

// ---------------------------------------------------------------------------------------------
// distributions/io/schema.proto, tag 2.0.28 (requirements.txt:18) -- STAND-IN: the package is not
// vendored in the reference tree, so these numbers restate its published schema from memory and
// cannot be checked offline.  Message and field NAMES are confirmed by loom's own uses
// (src/kind_proposer.cc:135-136, src/product_model.cc:60, src/query_server.cc:294-299,
// loom/format.py:348-351, loom/test/test_query.py:69-72).
namespace dist {
enum { PY_ALPHA = 1, PY_D = 2 };                                     // Clustering.PitmanYor
enum { BB_ALPHA = 1, BB_BETA = 2, BB_HEADS = 1, BB_TAILS = 2 };      // BetaBernoulli.Shared / .Group
enum { DD_ALPHAS = 1, DD_COUNTS = 1 };                               // DirichletDiscrete
enum { DPD_GAMMA = 1, DPD_ALPHA = 2, DPD_VALUES = 3, DPD_BETAS = 4, DPD_COUNTS = 5,   // DirichletProcessDiscrete.Shared
       DPD_KEYS = 1, DPD_GVALUES = 2 };                              // .Group (sparse counter)
enum { GP_ALPHA = 1, GP_INV_BETA = 2, GP_COUNT = 1, GP_SUM = 2, GP_LOG_PROD = 3 };   // GammaPoisson
enum { NICH_MU = 1, NICH_KAPPA = 2, NICH_SIGMASQ = 3, NICH_NU = 4,   // NormalInverseChiSq.Shared
       NICH_COUNT = 1, NICH_MEAN = 2, NICH_CTV = 3 };                // .Group
}  // namespace dist

// ---------------------------------------------------------------------------------------------
static thread_local std::string g_error;

static int fail(const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_error = buf;
    return 1;
}

extern "C" const char *lbio_last_error(void) { return g_error.c_str(); }

static bool read_whole(const char *path, std::vector<uint8_t> &out) {
    pbw::InFile f(path);
    if (!f.is_open()) { fail("failed to open input file %s", path); return false; }
    if (!f.read_all(out)) { fail("failed to read %s", path); return false; }
    return true;
}

static int write_whole(const char *path, const Writer &w) {
    pbw::OutFile f(path);
    if (!f.is_open()) return fail("failed to open output file %s", path);
    if (!f.write(w.buf) || !f.close()) return fail("failed to write %s", path);
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Config (src/schema.proto:167-229)
extern "C" void lbio_config_defaults(lbio_config *c) {   // loom/config.py:34-74
    std::memset(c, 0, sizeof(*c));
    c->seed = 0;
    c->target_mem_bytes = 4e9f;
    c->extra_passes = 500.0f;
    c->small_data_size = 4e3f;
    c->big_data_size = 1e9f;
    c->max_reject_iters = 100;
    c->checkpoint_period_sec = 1e9f;
    c->cat_empty_group_count = 1;
    c->cat_row_queue_capacity = 255;
    c->cat_parser_threads = 6;
    c->hyper_run = 1;
    c->hyper_parallel = 1;
    c->kind_iterations = 32;
    c->kind_empty_kind_count = 32;
    c->kind_row_queue_capacity = 255;
    c->kind_parser_threads = 6;
    c->kind_score_parallel = 1;
    c->posterior_enum_sample_count = 100;
    c->posterior_enum_sample_skip = 10;
    c->generate_row_count = 100;
    c->generate_density = 0.5f;
    c->generate_sample_skip = 10;
    c->has_query = 1;
    c->query_parallel = 1;
}

extern "C" int lbio_config_read(const char *path, lbio_config *c) {
    std::vector<uint8_t> bytes;
    if (!read_whole(path, bytes)) return 1;
    std::memset(c, 0, sizeof(*c));
    Reader r(bytes);
    uint32_t f, wt;
    unsigned seen = 0;
    while (r.next(f, wt)) {
        if (f == 1 && wt == pbw::VARINT) { c->seed = r.varint(); seen |= 1; }
        else if (f == 2 && wt == pbw::BYTES) {               // Schedule :169-176
            seen |= 2;
            Reader s = r.sub();
            while (s.next(f, wt)) {
                if (f == 1 && wt == pbw::FIXED32) c->extra_passes = s.f32();
                else if (f == 2 && wt == pbw::FIXED32) c->small_data_size = s.f32();
                else if (f == 3 && wt == pbw::FIXED32) c->big_data_size = s.f32();
                else if (f == 4 && wt == pbw::VARINT) c->max_reject_iters = (uint32_t)s.varint();
                else if (f == 5 && wt == pbw::FIXED32) c->checkpoint_period_sec = s.f32();
                else s.skip(wt);
            }
            r.ok = r.ok && s.ok;
        } else if (f == 3 && wt == pbw::BYTES) {             // Kernels :177-201
            seen |= 4;
            Reader k = r.sub();
            while (k.next(f, wt)) {
                if (wt != pbw::BYTES) { k.skip(wt); continue; }
                const uint32_t which = f;
                Reader s = k.sub();
                while (s.next(f, wt)) {
                    if (wt != pbw::VARINT) { s.skip(wt); continue; }
                    const uint64_t v = s.varint();
                    if (which == 1) {          // Cat
                        if (f == 1) c->cat_empty_group_count = (uint32_t)v;
                        else if (f == 2) c->cat_row_queue_capacity = (uint32_t)v;
                        else if (f == 3) c->cat_parser_threads = (uint32_t)v;
                    } else if (which == 2) {   // Hyper
                        if (f == 1) c->hyper_run = v != 0;
                        else if (f == 2) c->hyper_parallel = v != 0;
                    } else if (which == 3) {   // Kind
                        if (f == 1) c->kind_iterations = (uint32_t)v;
                        else if (f == 2) c->kind_empty_kind_count = (uint32_t)v;
                        else if (f == 3) c->kind_row_queue_capacity = (uint32_t)v;
                        else if (f == 4) c->kind_parser_threads = (uint32_t)v;
                        else if (f == 5) c->kind_score_parallel = v != 0;
                    }
                }
                k.ok = k.ok && s.ok;
            }
            r.ok = r.ok && k.ok;
        } else if (f == 4 && wt == pbw::BYTES) {             // PosteriorEnum :207-211
            Reader s = r.sub();
            while (s.next(f, wt)) {
                if (f == 1 && wt == pbw::VARINT) c->posterior_enum_sample_count = (uint32_t)s.varint();
                else if (f == 2 && wt == pbw::VARINT) c->posterior_enum_sample_skip = (uint32_t)s.varint();
                else s.skip(wt);
            }
            r.ok = r.ok && s.ok;
        } else if (f == 5 && wt == pbw::BYTES) {             // Generate :212-217
            Reader s = r.sub();
            while (s.next(f, wt)) {
                if (f == 1 && wt == pbw::VARINT) c->generate_row_count = s.varint();
                else if (f == 2 && wt == pbw::FIXED32) c->generate_density = s.f32();
                else if (f == 3 && wt == pbw::VARINT) c->generate_sample_skip = (uint32_t)s.varint();
                else s.skip(wt);
            }
            r.ok = r.ok && s.ok;
        } else if (f == 6 && wt == pbw::FIXED32) c->target_mem_bytes = r.f32();
        else if (f == 7 && wt == pbw::BYTES) {               // Query :218-221
            c->has_query = 1;
            Reader s = r.sub();
            while (s.next(f, wt)) {
                if (f == 1 && wt == pbw::VARINT) c->query_parallel = s.varint() != 0;
                else s.skip(wt);
            }
            r.ok = r.ok && s.ok;
        } else r.skip(wt);
    }
    if (!r.ok) return fail("failed to parse message from %s", path);
    if (seen != 7) return fail("%s: Config lacks seed / schedule / kernels", path);
    return 0;
}

extern "C" int lbio_config_write(const char *path, const lbio_config *c) {
    Writer w, s, k, t;
    w.u64(1, c->seed);
    s.f32(1, c->extra_passes); s.f32(2, c->small_data_size); s.f32(3, c->big_data_size);
    s.u64(4, c->max_reject_iters); s.f32(5, c->checkpoint_period_sec);
    w.msg(2, s);
    t.u64(1, c->cat_empty_group_count); t.u64(2, c->cat_row_queue_capacity); t.u64(3, c->cat_parser_threads);
    k.msg(1, t); t.clear();
    t.boolean(1, c->hyper_run); t.boolean(2, c->hyper_parallel);
    k.msg(2, t); t.clear();
    t.u64(1, c->kind_iterations); t.u64(2, c->kind_empty_kind_count); t.u64(3, c->kind_row_queue_capacity);
    t.u64(4, c->kind_parser_threads); t.boolean(5, c->kind_score_parallel);
    k.msg(3, t); t.clear();
    w.msg(3, k);
    t.u64(1, c->posterior_enum_sample_count); t.u64(2, c->posterior_enum_sample_skip);
    w.msg(4, t); t.clear();
    t.u64(1, c->generate_row_count); t.f32(2, c->generate_density); t.u64(3, c->generate_sample_skip);
    w.msg(5, t); t.clear();
    w.f32(6, c->target_mem_bytes);
    if (c->has_query) { t.boolean(1, c->query_parallel); w.msg(7, t); }
    return write_whole(path, w);
}

// ---------------------------------------------------------------------------------------------
// Checkpoint (src/schema.proto:234-253)
extern "C" int lbio_checkpoint_read(const char *path, lbio_checkpoint *c) {
    std::vector<uint8_t> bytes;
    if (!read_whole(path, bytes)) return 1;
    std::memset(c, 0, sizeof(*c));
    Reader r(bytes);
    uint32_t f, wt;
    while (r.next(f, wt)) {
        if (f == 1 && wt == pbw::VARINT) c->finished = r.varint() != 0;
        else if (f == 2 && wt == pbw::VARINT) c->seed = r.varint();
        else if (f == 3 && wt == pbw::VARINT) c->tardis_iter = r.varint();
        else if (f == 4 && wt == pbw::BYTES) {
            Reader s = r.sub();
            while (s.next(f, wt)) {
                if (f == 1 && wt == pbw::FIXED64) c->annealing_state = s.f64();
                else if (f == 2 && wt == pbw::VARINT) c->schedule_row_count = s.varint();
                else if (f == 3 && wt == pbw::VARINT) c->reject_iters = s.varint();
                else s.skip(wt);
            }
            r.ok = r.ok && s.ok;
        } else if (f == 5 && wt == pbw::VARINT) c->row_count = r.varint();
        else if (f == 6 && wt == pbw::BYTES) {
            Reader s = r.sub();
            while (s.next(f, wt)) {
                if (f == 1 && wt == pbw::VARINT) c->unassigned_pos = s.varint();
                else if (f == 2 && wt == pbw::VARINT) c->assigned_pos = s.varint();
                else s.skip(wt);
            }
            r.ok = r.ok && s.ok;
        } else r.skip(wt);
    }
    if (!r.ok) return fail("failed to parse message from %s", path);
    return 0;
}

extern "C" int lbio_checkpoint_write(const char *path, const lbio_checkpoint *c) {
    Writer w, s;
    w.boolean(1, c->finished);
    w.u64(2, c->seed);
    w.u64(3, c->tardis_iter);
    s.f64(1, c->annealing_state); s.u64(2, c->schedule_row_count); s.u64(3, c->reject_iters);
    w.msg(4, s); s.clear();
    w.u64(5, c->row_count);
    s.u64(1, c->unassigned_pos); s.u64(2, c->assigned_pos);
    w.msg(6, s);
    return write_whole(path, w);
}

// ---------------------------------------------------------------------------------------------
// Model
struct lbio_model {
    std::vector<lb_feature_spec> specs;
    std::vector<float> aux;
    std::vector<uint32_t> f2k;
    std::vector<float> kind_py;
    float topology[2] = {0, 0};
    bool has_prior = false;
    std::vector<float> grid[13];           // in the order of lb_hyper_prior's members
    std::vector<int32_t> dpd_offsets;
    std::vector<uint32_t> dpd_values;
    std::vector<uint64_t> dpd_counts;      // Shared.counts as read (written back unchanged)
    // derived
    int nb = 0, nc = 0, nr = 0;            // ValueSchema booleans / counts / reals sizes (product_value.hpp:112-142)
    std::vector<int32_t> dpd_ordinal;      // feature -> index among DPD features, or -1
    std::vector<std::vector<std::pair<uint32_t, uint32_t>>> dpd_lookup;   // per DPD feature: sorted (raw value, index)
};

static int type_rank(int t) { return t; }   // LB_BB < LB_DD < LB_DPD < LB_GP < LB_NICH is loom's order

static int model_finish(lbio_model *m) {
    const int F = (int)m->specs.size();
    m->nb = m->nc = m->nr = 0;
    m->dpd_ordinal.assign(F, -1);
    int n_dpd = 0;
    for (int f = 0; f < F; ++f) {
        const lb_feature_spec &s = m->specs[f];
        if (s.type < LB_BB || s.type > LB_NICH) return fail("feature %d has unknown type %d", f, s.type);
        if (f) {
            const lb_feature_spec &p = m->specs[f - 1];
            if (type_rank(p.type) > type_rank(s.type) || (p.type == LB_DD && s.type == LB_DD && p.dim > s.dim))
                return fail("features are not in loom's order (bb < dd by dim < dpd < gp < nich) at feature %d", f);
        }
        if (s.type == LB_DD && (s.dim < 1 || s.dim > 256)) return fail("dim is too large: %d", s.dim);   // product_model.cc:58-68
        if (s.type == LB_BB) ++m->nb;
        else if (s.type == LB_NICH) ++m->nr;
        else ++m->nc;
        if (s.type == LB_DPD) m->dpd_ordinal[f] = n_dpd++;
    }
    if (m->dpd_offsets.empty()) {           // identity dictionaries
        m->dpd_offsets.push_back(0);
        for (int f = 0; f < F; ++f)
            if (m->specs[f].type == LB_DPD) {
                for (int v = 0; v < m->specs[f].dim; ++v) m->dpd_values.push_back((uint32_t)v);
                m->dpd_offsets.push_back((int32_t)m->dpd_values.size());
            }
    }
    if ((int)m->dpd_offsets.size() != n_dpd + 1) return fail("dpd_offsets must have n_dpd + 1 entries");
    m->dpd_counts.resize(m->dpd_values.size(), 0);
    m->dpd_lookup.assign(n_dpd, {});
    for (int f = 0; f < F; ++f) {
        const int o = m->dpd_ordinal[f];
        if (o < 0) continue;
        const int n = m->dpd_offsets[o + 1] - m->dpd_offsets[o];
        if (n != m->specs[f].dim) return fail("DPD feature %d: %d betas but %d dictionary values", f, m->specs[f].dim, n);
        auto &lk = m->dpd_lookup[o];
        for (int i = 0; i < n; ++i) lk.emplace_back(m->dpd_values[m->dpd_offsets[o] + i], (uint32_t)i);
        std::sort(lk.begin(), lk.end());
        for (int i = 1; i < n; ++i)
            if (lk[i].first == lk[i - 1].first) return fail("DPD feature %d: duplicate dictionary value %u", f, lk[i].first);
    }
    const int K = (int)m->kind_py.size() / 2;
    for (int f = 0; f < F; ++f)
        if (m->f2k[f] >= (uint32_t)K) return fail("feature %d appears in no kind", f);
    return 0;
}

static void read_py(Reader s, float out[2]) {
    uint32_t f, wt;
    while (s.next(f, wt)) {
        if (f == dist::PY_ALPHA && wt == pbw::FIXED32) out[0] = s.f32();
        else if (f == dist::PY_D && wt == pbw::FIXED32) out[1] = s.f32();
        else s.skip(wt);
    }
}

static void write_py(Writer &w, uint32_t field, float alpha, float d) {
    Writer s;
    s.f32(dist::PY_ALPHA, alpha);
    s.f32(dist::PY_D, d);
    w.msg(field, s);
}

struct SharedMsg {       // one feature's Shared as read from a kind
    int type = 0;
    float hp[4] = {0, 0, 0, 0};
    std::vector<float> vec;            // DD alphas / DPD betas
    std::vector<uint32_t> values;      // DPD values
    std::vector<uint64_t> counts;      // DPD counts
};

static bool read_shared(Reader s, int type, SharedMsg &out) {
    out.type = type;
    uint32_t f, wt;
    while (s.next(f, wt)) {
        switch (type) {
        case LB_BB:
            if (f == dist::BB_ALPHA && wt == pbw::FIXED32) out.hp[0] = s.f32();
            else if (f == dist::BB_BETA && wt == pbw::FIXED32) out.hp[1] = s.f32();
            else s.skip(wt);
            break;
        case LB_DD:
            if (f == dist::DD_ALPHAS) s.rep_f32(wt, out.vec);
            else s.skip(wt);
            break;
        case LB_DPD:
            if (f == dist::DPD_GAMMA && wt == pbw::FIXED32) out.hp[0] = s.f32();
            else if (f == dist::DPD_ALPHA && wt == pbw::FIXED32) out.hp[1] = s.f32();
            else if (f == dist::DPD_VALUES) s.rep_varint(wt, out.values);
            else if (f == dist::DPD_BETAS) s.rep_f32(wt, out.vec);
            else if (f == dist::DPD_COUNTS) s.rep_varint(wt, out.counts);
            else s.skip(wt);
            break;
        case LB_GP:
            if (f == dist::GP_ALPHA && wt == pbw::FIXED32) out.hp[0] = s.f32();
            else if (f == dist::GP_INV_BETA && wt == pbw::FIXED32) out.hp[1] = s.f32();
            else s.skip(wt);
            break;
        case LB_NICH:
            if (f >= 1 && f <= 4 && wt == pbw::FIXED32) out.hp[f - 1] = s.f32();   // mu, kappa, sigmasq, nu
            else s.skip(wt);
            break;
        }
    }
    return s.ok;
}

extern "C" int lbio_model_read(const char *path, lbio_model **out) {
    std::vector<uint8_t> bytes;
    if (!read_whole(path, bytes)) return 1;
    struct KindMsg { float py[2] = {0, 0}; std::vector<SharedMsg> shareds; std::vector<uint32_t> featureids; };
    std::vector<KindMsg> kinds;
    auto *m = new lbio_model;
    Reader r(bytes);
    uint32_t f, wt;
    bool bad = false;
    while (r.next(f, wt)) {
        if (f == 1 && wt == pbw::BYTES) {                     // CrossCat.kinds :133
            kinds.emplace_back();
            KindMsg &kind = kinds.back();
            Reader k = r.sub();
            while (k.next(f, wt)) {
                if (f == 1 && wt == pbw::BYTES) {             // Kind.product_model :128, ProductModel.Shared :103-111
                    Reader pm = k.sub();
                    std::vector<SharedMsg> by_type[5];
                    while (pm.next(f, wt)) {
                        if (f == 1 && wt == pbw::BYTES) read_py(pm.sub(), kind.py);
                        else if (f >= 2 && f <= 7 && wt == pbw::BYTES) {
                            if (f == 6) { delete m; return fail("%s: BetaNegativeBinomial features are not supported", path); }
                            const int type = f == 7 ? LB_NICH : (int)f - 2;    // bb 2, dd 3, dpd 4, gp 5, nich 7
                            by_type[type].emplace_back();
                            bad |= !read_shared(pm.sub(), type, by_type[type].back());
                        } else pm.skip(wt);
                    }
                    bad |= !pm.ok;
                    for (auto &list : by_type)
                        for (auto &s : list) kind.shareds.push_back(std::move(s));
                } else if (f == 2) k.rep_varint(wt, kind.featureids);   // Kind.featureids :129
                else k.skip(wt);
            }
            bad |= !k.ok;
        } else if (f == 2 && wt == pbw::BYTES) read_py(r.sub(), m->topology);   // CrossCat.topology :134
        else if (f == 3 && wt == pbw::BYTES) {                // CrossCat.hyper_prior :135, HyperPrior :34-71
            m->has_prior = true;
            Reader h = r.sub();
            while (h.next(f, wt)) {
                if (wt != pbw::BYTES) { h.skip(wt); continue; }
                const uint32_t which = f;
                Reader s = h.sub();
                if (which == 1 || which == 2) {               // topology / clustering grids of PitmanYor
                    float py[2] = {0, 0};
                    read_py(s, py);
                    m->grid[which - 1].push_back(py[0]);
                    m->grid[which - 1].push_back(py[1]);
                    continue;
                }
                while (s.next(f, wt)) {
                    int g = -1;
                    if (which == 3) g = f == 1 ? 2 : f == 2 ? 3 : -1;                 // bb alpha, beta
                    else if (which == 4) g = f == 1 ? 4 : -1;                           // dd alpha
                    else if (which == 5) g = f == 1 ? 5 : f == 2 ? 6 : -1;              // dpd gamma, alpha
                    else if (which == 6) g = f == 1 ? 7 : f == 2 ? 8 : -1;              // gp alpha, inv_beta
                    else if (which == 8) g = (f >= 1 && f <= 4) ? 8 + (int)f : -1;      // nich mu, kappa, sigmasq, nu
                    if (g >= 0) s.rep_f32(wt, m->grid[g]);
                    else s.skip(wt);
                }
                h.ok = h.ok && s.ok;
            }
            bad |= !h.ok;
        } else r.skip(wt);
    }
    if (bad || !r.ok) { delete m; return fail("failed to parse message from %s", path); }

    // CrossCat::model_load (src/cross_cat.cc:50-110)
    size_t F = 0;
    for (size_t k = 0; k < kinds.size(); ++k) {
        if (kinds[k].featureids.empty()) { delete m; return fail("kind %zu has no features", k); }
        if (kinds[k].featureids.size() != kinds[k].shareds.size()) {
            delete m;
            return fail("kind has %zu features, but featureids has %zu entries", kinds[k].shareds.size(), kinds[k].featureids.size());
        }
        F += kinds[k].featureids.size();
    }
    if (kinds.empty()) { delete m; return fail("no kinds, loom is empty"); }
    std::vector<const SharedMsg *> by_feature(F, nullptr);
    m->f2k.assign(F, 0xFFFFFFFFu);
    for (size_t k = 0; k < kinds.size(); ++k) {
        m->kind_py.push_back(kinds[k].py[0]);
        m->kind_py.push_back(kinds[k].py[1]);
        for (size_t i = 0; i < kinds[k].featureids.size(); ++i) {
            const uint32_t fid = kinds[k].featureids[i];
            if (fid >= F) { delete m; return fail("featureid out of bounds: %u", fid); }
            if (by_feature[fid]) { delete m; return fail("kind %zu has duplicate feature %u", k, fid); }
            by_feature[fid] = &kinds[k].shareds[i];
            m->f2k[fid] = (uint32_t)k;
        }
    }
    m->dpd_offsets.push_back(0);
    for (size_t fid = 0; fid < F; ++fid) {
        const SharedMsg &s = *by_feature[fid];
        lb_feature_spec spec;
        std::memset(&spec, 0, sizeof(spec));
        spec.type = s.type;
        std::memcpy(spec.hp, s.hp, sizeof(spec.hp));
        if (s.type == LB_DD || s.type == LB_DPD) {
            spec.dim = (int32_t)s.vec.size();
            spec.aux_off = (int32_t)m->aux.size();
            m->aux.insert(m->aux.end(), s.vec.begin(), s.vec.end());
        }
        if (s.type == LB_DPD) {
            if (s.values.size() != s.vec.size()) { delete m; return fail("DPD feature %zu: values and betas differ in length", fid); }
            double sum = 0;
            for (float b : s.vec) sum += b;
            spec.hp[2] = (float)std::max(0.0, 1.0 - sum);     // beta0 = 1 - sum betas
            m->dpd_values.insert(m->dpd_values.end(), s.values.begin(), s.values.end());
            for (size_t i = 0; i < s.values.size(); ++i) m->dpd_counts.push_back(i < s.counts.size() ? s.counts[i] : 0);
            m->dpd_offsets.push_back((int32_t)m->dpd_values.size());
        }
        m->specs.push_back(spec);
    }
    if (model_finish(m)) { delete m; return 1; }
    *out = m;
    return 0;
}

extern "C" int lbio_model_create(int32_t n_features, const lb_feature_spec *specs, const float *aux, int32_t n_aux,
                                 int32_t n_kinds, const uint32_t *f2k, const float *kind_py, const float *topology_py,
                                 const lb_hyper_prior *prior, const int32_t *dpd_offsets, const uint32_t *dpd_values,
                                 lbio_model **out) {
    auto *m = new lbio_model;
    m->specs.assign(specs, specs + n_features);
    if (n_aux > 0) m->aux.assign(aux, aux + n_aux);
    m->f2k.assign(f2k, f2k + n_features);
    m->kind_py.assign(kind_py, kind_py + 2 * (size_t)n_kinds);
    m->topology[0] = topology_py[0];
    m->topology[1] = topology_py[1];
    if (prior) {
        m->has_prior = true;
        const float *const ptrs[13] = {prior->topology, prior->clustering, prior->bb_alpha, prior->bb_beta, prior->dd_alpha,
                                       prior->dpd_gamma, prior->dpd_alpha, prior->gp_alpha, prior->gp_inv_beta, prior->nich_mu,
                                       prior->nich_kappa, prior->nich_sigmasq, prior->nich_nu};
        const int32_t lens[13] = {prior->n_topology * 2, prior->n_clustering * 2, prior->n_bb_alpha, prior->n_bb_beta,
                                  prior->n_dd_alpha, prior->n_dpd_gamma, prior->n_dpd_alpha, prior->n_gp_alpha,
                                  prior->n_gp_inv_beta, prior->n_nich_mu, prior->n_nich_kappa, prior->n_nich_sigmasq,
                                  prior->n_nich_nu};
        for (int i = 0; i < 13; ++i)
            if (ptrs[i] && lens[i] > 0) m->grid[i].assign(ptrs[i], ptrs[i] + lens[i]);
    }
    if (dpd_offsets) {
        int n_dpd = 0;
        for (int f = 0; f < n_features; ++f) n_dpd += specs[f].type == LB_DPD;
        m->dpd_offsets.assign(dpd_offsets, dpd_offsets + n_dpd + 1);
        m->dpd_values.assign(dpd_values, dpd_values + dpd_offsets[n_dpd]);
    }
    for (int f = 0; f < n_features; ++f) {
        const lb_feature_spec &s = m->specs[f];
        if ((s.type == LB_DD || s.type == LB_DPD) && (s.aux_off < 0 || s.aux_off + s.dim > n_aux)) {
            delete m;
            return fail("feature %d: aux range out of bounds", f);
        }
    }
    if (model_finish(m)) { delete m; return 1; }
    *out = m;
    return 0;
}

// ValueSchema::load(ProductValue) (src/product_value.hpp:125-130): a schema row only carries the SIZES of the three
// value classes.  The tare / sparsify passes need nothing more, so the model built from it types every count column
// as GammaPoisson (a 32-bit cell that takes any value) and puts everything in one kind.
extern "C" int lbio_model_from_schema_row(const char *path, lbio_model **out) {
    pbw::InFile file(path);
    if (!file.is_open()) return fail("failed to open input file %s", path);
    std::vector<uint8_t> msg;
    if (!file.read_all(msg)) return fail("failed to parse message from %s", path);
    Reader r(msg);
    std::vector<uint8_t> booleans;
    std::vector<uint32_t> counts;
    std::vector<float> reals;
    uint32_t f, wt;
    while (r.next(f, wt)) {
        if (f == 2) r.rep_varint(wt, booleans);
        else if (f == 3) r.rep_varint(wt, counts);
        else if (f == 4) r.rep_f32(wt, reals);
        else r.skip(wt);
    }
    if (!r.ok) return fail("failed to parse message from %s", path);
    const size_t F = booleans.size() + counts.size() + reals.size();
    if (!F) return fail("schema row %s has no fields", path);
    std::vector<lb_feature_spec> specs(F);
    for (size_t i = 0; i < F; ++i) {
        lb_feature_spec &s = specs[i];
        std::memset(&s, 0, sizeof(s));
        s.type = i < booleans.size() ? LB_BB : i < booleans.size() + counts.size() ? LB_GP : LB_NICH;
        s.hp[0] = s.type == LB_NICH ? 0.f : 1.f;
        s.hp[1] = s.hp[2] = s.hp[3] = 1.f;
    }
    const std::vector<uint32_t> f2k(F, 0);
    const float py[2] = {1.f, 0.f};
    return lbio_model_create((int32_t)F, specs.data(), nullptr, 0, 1, f2k.data(), py, py, nullptr, nullptr, nullptr, out);
}

extern "C" int lbio_model_get(const lbio_model *m, lbio_model_view *v) {
    std::memset(v, 0, sizeof(*v));
    v->n_features = (int32_t)m->specs.size();
    v->specs = m->specs.data();
    v->aux = m->aux.data();
    v->n_aux = (int32_t)m->aux.size();
    v->n_kinds = (int32_t)m->kind_py.size() / 2;
    v->featureid_to_kindid = m->f2k.data();
    v->kind_py = m->kind_py.data();
    v->topology_py[0] = m->topology[0];
    v->topology_py[1] = m->topology[1];
    v->has_hyper_prior = m->has_prior;
    const float **ptrs[13] = {&v->hyper_prior.topology, &v->hyper_prior.clustering, &v->hyper_prior.bb_alpha,
                              &v->hyper_prior.bb_beta, &v->hyper_prior.dd_alpha, &v->hyper_prior.dpd_gamma,
                              &v->hyper_prior.dpd_alpha, &v->hyper_prior.gp_alpha, &v->hyper_prior.gp_inv_beta,
                              &v->hyper_prior.nich_mu, &v->hyper_prior.nich_kappa, &v->hyper_prior.nich_sigmasq,
                              &v->hyper_prior.nich_nu};
    int32_t *lens[13] = {&v->hyper_prior.n_topology, &v->hyper_prior.n_clustering, &v->hyper_prior.n_bb_alpha,
                         &v->hyper_prior.n_bb_beta, &v->hyper_prior.n_dd_alpha, &v->hyper_prior.n_dpd_gamma,
                         &v->hyper_prior.n_dpd_alpha, &v->hyper_prior.n_gp_alpha, &v->hyper_prior.n_gp_inv_beta,
                         &v->hyper_prior.n_nich_mu, &v->hyper_prior.n_nich_kappa, &v->hyper_prior.n_nich_sigmasq,
                         &v->hyper_prior.n_nich_nu};
    for (int i = 0; i < 13; ++i) {
        *ptrs[i] = m->grid[i].empty() ? nullptr : m->grid[i].data();
        *lens[i] = (int32_t)(i < 2 ? m->grid[i].size() / 2 : m->grid[i].size());
    }
    v->n_dpd = (int32_t)m->dpd_offsets.size() - 1;
    v->dpd_offsets = m->dpd_offsets.data();
    v->dpd_values = m->dpd_values.data();
    return 0;
}

static void write_shared(Writer &pm, const lbio_model *m, int f) {
    const lb_feature_spec &s = m->specs[f];
    Writer w;
    switch (s.type) {
    case LB_BB:
        w.f32(dist::BB_ALPHA, s.hp[0]); w.f32(dist::BB_BETA, s.hp[1]);
        pm.msg(2, w);
        break;
    case LB_DD:
        for (int v = 0; v < s.dim; ++v) w.f32(dist::DD_ALPHAS, m->aux[s.aux_off + v]);
        pm.msg(3, w);
        break;
    case LB_DPD: {
        const int o = m->dpd_ordinal[f], at = m->dpd_offsets[o];
        w.f32(dist::DPD_GAMMA, s.hp[0]); w.f32(dist::DPD_ALPHA, s.hp[1]);
        for (int v = 0; v < s.dim; ++v) w.u64(dist::DPD_VALUES, m->dpd_values[at + v]);
        for (int v = 0; v < s.dim; ++v) w.f32(dist::DPD_BETAS, m->aux[s.aux_off + v]);
        for (int v = 0; v < s.dim; ++v) w.u64(dist::DPD_COUNTS, m->dpd_counts[at + v]);
        pm.msg(4, w);
    } break;
    case LB_GP:
        w.f32(dist::GP_ALPHA, s.hp[0]); w.f32(dist::GP_INV_BETA, s.hp[1]);
        pm.msg(5, w);
        break;
    case LB_NICH:
        for (int i = 0; i < 4; ++i) w.f32(dist::NICH_MU + i, s.hp[i]);
        pm.msg(7, w);
        break;
    }
}

// CrossCat::model_dump (src/cross_cat.cc:112-135)
extern "C" int lbio_model_write(const lbio_model *m, const char *path) {
    const int F = (int)m->specs.size(), K = (int)m->kind_py.size() / 2;
    Writer w;
    for (int k = 0; k < K; ++k) {
        Writer kind, pm;
        write_py(pm, 1, m->kind_py[2 * k], m->kind_py[2 * k + 1]);
        int n = 0;
        for (int f = 0; f < F; ++f)          // ascending feature id = bb, dd, dpd, gp, nich order
            if ((int)m->f2k[f] == k) { write_shared(pm, m, f); ++n; }
        if (!n) return fail("kind %d has no features", k);
        kind.msg(1, pm);
        for (int f = 0; f < F; ++f)
            if ((int)m->f2k[f] == k) kind.u64(2, (uint64_t)f);
        w.msg(1, kind);
    }
    write_py(w, 2, m->topology[0], m->topology[1]);
    if (m->has_prior) {
        Writer h, s;
        for (int which = 0; which < 2; ++which)
            for (size_t i = 0; i + 1 < m->grid[which].size(); i += 2)
                write_py(h, which + 1, m->grid[which][i], m->grid[which][i + 1]);
        auto rep = [&](Writer &dst, uint32_t field, const std::vector<float> &v) { for (float x : v) dst.f32(field, x); };
        rep(s, 1, m->grid[2]); rep(s, 2, m->grid[3]); h.msg(3, s); s.clear();
        rep(s, 1, m->grid[4]); h.msg(4, s); s.clear();
        rep(s, 1, m->grid[5]); rep(s, 2, m->grid[6]); h.msg(5, s); s.clear();
        rep(s, 1, m->grid[7]); rep(s, 2, m->grid[8]); h.msg(6, s); s.clear();
        for (int i = 0; i < 4; ++i) rep(s, 1 + i, m->grid[9 + i]);
        h.msg(8, s);
        w.msg(3, h);
    }
    return write_whole(path, w);
}

extern "C" void lbio_model_free(lbio_model *m) { delete m; }

// ---------------------------------------------------------------------------------------------
// ProductValue -> cells (src/product_value.hpp:586-767)
struct ValueScratch {
    std::vector<uint8_t> dense, booleans;
    std::vector<uint32_t> sparse, counts;
    std::vector<float> reals;
};

enum { SP_NONE = 0, SP_SPARSE = 1, SP_DENSE = 2, SP_ALL = 3 };   // ProductValue.Observed.Sparsity :76-81

// Decodes one ProductValue; appends (feature, raw value) in ascending feature order.
// `values` may be null (only the observed positions are wanted).  Returns an error text or null.
// With `misses` set, a DPD value missing from the dictionary is not an error: its RAW value is stored and
// the index of the cell within `values` is recorded for the caller to resolve (open dictionaries).
static const char *decode_value(Reader r, const lbio_model &m, ValueScratch &t, std::vector<uint32_t> &features,
                                std::vector<uint32_t> *values, std::vector<uint64_t> *misses = nullptr) {
    t.dense.clear(); t.booleans.clear(); t.sparse.clear(); t.counts.clear(); t.reals.clear();
    int sparsity = -1;
    uint32_t f, wt;
    while (r.next(f, wt)) {
        if (f == 1 && wt == pbw::BYTES) {                     // observed :93
            Reader o = r.sub();
            while (o.next(f, wt)) {
                if (f == 1 && wt == pbw::VARINT) sparsity = (int)o.varint();
                else if (f == 2) o.rep_varint(wt, t.dense);
                else if (f == 3) o.rep_varint(wt, t.sparse);
                else o.skip(wt);
            }
            if (!o.ok) return "failed to parse ProductValue.Observed";
        } else if (f == 2) r.rep_varint(wt, t.booleans);      // booleans :94
        else if (f == 3) r.rep_varint(wt, t.counts);          // counts :95
        else if (f == 4) r.rep_f32(wt, t.reals);              // reals :96
        else r.skip(wt);
    }
    if (!r.ok) return "failed to parse ProductValue";
    const size_t F = m.specs.size(), nb = (size_t)m.nb, nc = (size_t)m.nc;
    size_t ib = 0, ic = 0, ir = 0;
    const char *err = nullptr;
    auto take = [&](size_t p) {
        const lb_feature_spec &s = m.specs[p];
        uint32_t raw = 0;
        if (p < nb) {
            if (ib >= t.booleans.size()) { err = "fewer booleans than observed bb features"; return; }
            raw = t.booleans[ib++] ? 1u : 0u;
        } else if (p < nb + nc) {
            if (ic >= t.counts.size()) { err = "fewer counts than observed dd/dpd/gp features"; return; }
            raw = t.counts[ic++];
            if (s.type == LB_DD) {
                if (raw >= (uint32_t)s.dim) { err = "DirichletDiscrete value out of range"; return; }
            } else if (s.type == LB_DPD) {
                const auto &lk = m.dpd_lookup[m.dpd_ordinal[p]];
                auto it = std::lower_bound(lk.begin(), lk.end(), std::make_pair(raw, 0u));
                if (it != lk.end() && it->first == raw) raw = it->second;
                else if (!values) {}                          // neg part of a diff: only the positions are used
                else if (misses) misses->push_back(values->size());
                else { err = "DirichletProcessDiscrete value is not in the model's dictionary"; return; }
            }
        } else {
            if (ir >= t.reals.size()) { err = "fewer reals than observed nich features"; return; }
            std::memcpy(&raw, &t.reals[ir++], 4);
        }
        features.push_back((uint32_t)p);
        if (values) values->push_back(raw);
    };
    switch (sparsity) {
    case SP_NONE: break;
    case SP_ALL:
        for (size_t p = 0; p < F && !err; ++p) take(p);
        break;
    case SP_DENSE:
        if (t.dense.size() != F) return "observed.dense has the wrong size";
        for (size_t p = 0; p < F && !err; ++p)
            if (t.dense[p]) take(p);
        break;
    case SP_SPARSE: {
        uint32_t prev = 0;
        bool first = true;
        for (uint32_t p : t.sparse) {
            if (p >= F) return "observed.sparse position out of range";
            if (!first && p <= prev) return "observed.sparse is not ascending";
            prev = p;
            first = false;
            take(p);
            if (err) break;
        }
    } break;
    default: return "ProductValue lacks observed.sparsity";
    }
    if (err) return err;
    if (ib != t.booleans.size() || ic != t.counts.size() || ir != t.reals.size()) return "more packed values than observed features";
    return nullptr;
}

struct RowsOwner {
    std::vector<uint64_t> rowids, pos_ptr, neg_ptr;
    std::vector<uint8_t> has_tare;
    std::vector<uint32_t> pos_feature, pos_value, neg_feature;
};

struct RowChunk {           // one thread's share of a batch
    std::vector<uint64_t> rowids;
    std::vector<uint8_t> has_tare;
    std::vector<uint32_t> pos_count, neg_count, pos_feature, pos_value, neg_feature;
    std::vector<uint64_t> misses;     // open dictionaries: cells of pos_value still holding a raw DPD value
    std::string error;
    void clear() { rowids.clear(); has_tare.clear(); pos_count.clear(); neg_count.clear(); pos_feature.clear(); pos_value.clear(); neg_feature.clear(); misses.clear(); error.clear(); }
};

// Row :153-156, ProductValue.Diff :87-91
// One ProductValue.Diff (:87-91) appended to `out` as a row with the given id.  Returns an error text or null
// (on error the partial cells are rolled back).
static const char *const BAD_TARE = "row references a tare other than tare 0 (one tare row is supported)";
static const char *decode_diff(Reader d, const lbio_model &m, ValueScratch &scratch, std::vector<uint32_t> &tares,
                               uint64_t id, RowChunk &out, bool open_dictionaries) {
    const size_t pos0 = out.pos_feature.size(), neg0 = out.neg_feature.size(), miss0 = out.misses.size();
    const char *err = nullptr;
    uint8_t has_tare = 0;
    uint32_t f, wt;
    bool has_pos = false, has_neg = false;
    tares.clear();
    while (!err && d.next(f, wt)) {
        if (f == 1 && wt == pbw::BYTES) {
            has_pos = true;
            err = decode_value(d.sub(), m, scratch, out.pos_feature, &out.pos_value, open_dictionaries ? &out.misses : nullptr);
        } else if (f == 2 && wt == pbw::BYTES) {
            has_neg = true;
            err = decode_value(d.sub(), m, scratch, out.neg_feature, nullptr);
        } else if (f == 3) d.rep_varint(wt, tares);
        else d.skip(wt);
    }
    if (!err && (!d.ok || !has_pos || !has_neg)) err = "failed to parse ProductValue.Diff (pos and neg are required)";
    for (uint32_t t : tares) {
        if (t != 0 && !err) err = BAD_TARE;
        has_tare = 1;
    }
    if (err) {
        out.pos_feature.resize(pos0); out.pos_value.resize(pos0); out.neg_feature.resize(neg0); out.misses.resize(miss0);
        return err;
    }
    out.rowids.push_back(id);
    out.has_tare.push_back(has_tare);
    out.pos_count.push_back((uint32_t)(out.pos_feature.size() - pos0));
    out.neg_count.push_back((uint32_t)(out.neg_feature.size() - neg0));
    return nullptr;
}

// Row :153-156
static void decode_rows(const lbio_model &m, const uint8_t *pool, const uint64_t *offsets, size_t begin, size_t end,
                        RowChunk &out, bool open_dictionaries) {
    ValueScratch scratch;
    std::vector<uint32_t> tares;
    for (size_t i = begin; i < end; ++i) {
        Reader r(pool + offsets[i], (size_t)(offsets[i + 1] - offsets[i]));
        uint64_t id = 0;
        bool has_id = false, has_diff = false;
        uint32_t f, wt;
        const char *err = nullptr;
        Reader diff(nullptr, 0);
        while (r.next(f, wt)) {
            if (f == 1 && wt == pbw::VARINT) { id = r.varint(); has_id = true; }
            else if (f == 2 && wt == pbw::BYTES) { diff = r.sub(); has_diff = true; }
            else r.skip(wt);
        }
        if (!r.ok || !has_id || !has_diff) err = "failed to parse Row (id and diff are required)";
        else err = decode_diff(diff, m, scratch, tares, id, out, open_dictionaries);
        if (err) {
            char buf[256];
            snprintf(buf, sizeof(buf), "row %zu of the batch: %s", i, err);
            out.error = buf;
            return;
        }
    }
}

static void publish_rows(RowsOwner *o, lbio_row_block *out) {
    out->n_rows = o->rowids.size();
    out->rowids = o->rowids.data();
    out->row_has_tare = o->has_tare.data();
    out->pos_ptr = o->pos_ptr.data();
    out->pos_feature = o->pos_feature.data();
    out->pos_value = o->pos_value.data();
    out->neg_ptr = o->neg_ptr.data();
    out->neg_feature = o->neg_feature.data();
    out->owner = o;
}

// `fresh` (open dictionaries): per DPD feature, the values met that the dictionary lacks, in stream order
static int rows_read_impl(const char *path, const lbio_model *m, uint64_t first_row, uint64_t max_rows,
                          int32_t n_threads, lbio_row_block *out, std::vector<std::vector<uint32_t>> *fresh) {
    std::memset(out, 0, sizeof(*out));
    std::vector<std::unordered_map<uint32_t, uint32_t>> fresh_index(fresh ? m->dpd_lookup.size() : 0);
    if (fresh) fresh->assign(m->dpd_lookup.size(), {});
    pbw::InFile file(path);
    if (!file.is_open()) return fail("failed to open input file %s", path);
    bool bad = false;
    uint32_t size = 0;
    while (file.position() < first_row) {
        if (!file.skip_stream(size, bad)) return fail("failed to set position of %s", path);   // InFile::set_position
    }
    int T = n_threads;
#ifdef _OPENMP
    if (T <= 0) T = omp_get_max_threads();
#else
    T = 1;
#endif
    T = std::max(1, std::min(T, 64));
    // gunzip + framing are serial (one gzip stream); they run one batch ahead of the decoders
    const size_t BATCH = 1 << 14;
    auto *o = new RowsOwner;
    o->pos_ptr.push_back(0);
    o->neg_ptr.push_back(0);
    struct Batch { std::vector<uint8_t> pool; std::vector<uint64_t> offsets; bool bad = false; };
    Batch batches[2];
    std::vector<RowChunk> chunks((size_t)T);
    bool at_end = false;
    uint64_t taken = 0;
    auto read_batch = [&](Batch &b) {
        b.pool.clear();
        b.offsets.assign(1, 0);
        b.bad = false;
        uint32_t sz = 0;
        bool broken = false;
        while (!at_end && b.offsets.size() <= BATCH && (max_rows == 0 || taken < max_rows)) {
            if (!file.try_read_stream_into(b.pool, sz, broken)) {
                b.bad = broken;
                at_end = true;
                break;
            }
            b.offsets.push_back(b.pool.size());
            ++taken;
        }
    };
    read_batch(batches[0]);
    for (int cur = 0;; cur ^= 1) {
        Batch &b = batches[cur];
        if (b.bad) { delete o; return fail("failed to parse message from %s (truncated stream)", path); }
        const size_t n = b.offsets.size() - 1;
        if (!n) break;
        std::future<void> ahead = std::async(std::launch::async, read_batch, std::ref(batches[cur ^ 1]));
        const int used = (int)std::min<size_t>((size_t)T, (n + 255) / 256);
#pragma omp parallel for num_threads(used) schedule(static, 1)
        for (int t = 0; t < used; ++t) {
            chunks[t].clear();
            decode_rows(*m, b.pool.data(), b.offsets.data(), n * t / used, n * (t + 1) / used, chunks[t], fresh != nullptr);
        }
        std::string error;
        for (int t = 0; t < used && error.empty(); ++t) {
            const RowChunk &c = chunks[t];
            if (!c.error.empty()) { error = c.error; break; }
            o->rowids.insert(o->rowids.end(), c.rowids.begin(), c.rowids.end());
            o->has_tare.insert(o->has_tare.end(), c.has_tare.begin(), c.has_tare.end());
            for (uint32_t k : c.pos_count) o->pos_ptr.push_back(o->pos_ptr.back() + k);
            for (uint32_t k : c.neg_count) o->neg_ptr.push_back(o->neg_ptr.back() + k);
            const size_t at = o->pos_value.size();
            o->pos_feature.insert(o->pos_feature.end(), c.pos_feature.begin(), c.pos_feature.end());
            o->pos_value.insert(o->pos_value.end(), c.pos_value.begin(), c.pos_value.end());
            for (uint64_t idx : c.misses) {      // serial and in row order: first appearance decides the index
                const int ord = m->dpd_ordinal[c.pos_feature[idx]];
                const uint32_t raw = c.pos_value[idx];
                auto found = fresh_index[ord].find(raw);
                if (found == fresh_index[ord].end()) {
                    const uint32_t index = (uint32_t)(m->dpd_lookup[ord].size() + (*fresh)[ord].size());
                    (*fresh)[ord].push_back(raw);
                    found = fresh_index[ord].emplace(raw, index).first;
                }
                o->pos_value[at + idx] = found->second;
            }
            o->neg_feature.insert(o->neg_feature.end(), c.neg_feature.begin(), c.neg_feature.end());
        }
        ahead.get();
        if (!error.empty()) { delete o; return fail("%s: %s", path, error.c_str()); }
    }
    publish_rows(o, out);
    out->stream_rows = at_end ? file.position() : 0;
    return 0;
}

extern "C" int lbio_rows_read(const char *path, const lbio_model *m, uint64_t first_row, uint64_t max_rows,
                              int32_t n_threads, lbio_row_block *out) {
    return rows_read_impl(path, m, first_row, max_rows, n_threads, out, nullptr);
}

// Philox4x32-10 (Salmon et al. 2011), the generator of include/loom_b200.h; u = (x0 >> 8) * 2^-24
static double philox_uniform(uint64_t seed, uint32_t stream, uint64_t a, uint32_t b) {
    uint32_t c0 = (uint32_t)a, c1 = (uint32_t)(a >> 32), c2 = b, c3 = stream;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int round = 0; round < 10; ++round) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return (double)(c0 >> 8) * (1.0 / 16777216.0);
}

#define LBIO_DPD_MIN_BETA 1e-6   // distributions' DirichletProcessDiscrete::MIN_BETA() is not vendored: stand-in

// Append values to DPD dictionaries; each new value takes beta_v = beta0 * Beta(1, gamma) off the stick
// (distributions DPD Shared.add_value).  Beta(1, gamma) by inverse CDF: 1 - (1 - u)^(1 / gamma).
static int model_extend(lbio_model *m, const std::vector<std::vector<uint32_t>> &fresh, uint64_t seed) {
    const int F = (int)m->specs.size();
    std::vector<float> aux;
    std::vector<uint32_t> values;
    std::vector<uint64_t> counts;
    std::vector<int32_t> offsets(1, 0);
    for (int f = 0; f < F; ++f) {
        lb_feature_spec &s = m->specs[f];
        if (s.type != LB_DD && s.type != LB_DPD) continue;
        const int old_off = s.aux_off;
        s.aux_off = (int32_t)aux.size();
        aux.insert(aux.end(), m->aux.begin() + old_off, m->aux.begin() + old_off + s.dim);
        if (s.type != LB_DPD) continue;
        const int ord = m->dpd_ordinal[f], at = m->dpd_offsets[ord];
        values.insert(values.end(), m->dpd_values.begin() + at, m->dpd_values.begin() + at + s.dim);
        counts.insert(counts.end(), m->dpd_counts.begin() + at, m->dpd_counts.begin() + at + s.dim);
        double beta0 = s.hp[2];
        const double gamma = s.hp[0] > 0 ? s.hp[0] : 1.0;
        uint32_t draw = (uint32_t)s.dim;        // draw index = position in the dictionary, so a later extension continues the stream
        for (uint32_t v : fresh[ord]) {
            const double u = philox_uniform(seed, 4u, (uint64_t)f, draw++);
            const double beta = std::max(beta0 * (1.0 - std::pow(1.0 - u, 1.0 / gamma)), (double)LBIO_DPD_MIN_BETA);
            beta0 = std::max(beta0 - beta, (double)LBIO_DPD_MIN_BETA);
            aux.push_back((float)beta);
            values.push_back(v);
            counts.push_back(0);
        }
        s.dim += (int32_t)fresh[ord].size();
        s.hp[2] = (float)beta0;
        offsets.push_back((int32_t)values.size());
    }
    m->aux.swap(aux);
    m->dpd_values.swap(values);
    m->dpd_counts.swap(counts);
    m->dpd_offsets.swap(offsets);
    return model_finish(m);
}

extern "C" int lbio_rows_read_open(const char *path, lbio_model *m, uint64_t seed, int32_t n_threads, lbio_row_block *out,
                                   int32_t *n_new) {
    std::vector<std::vector<uint32_t>> fresh;
    if (rows_read_impl(path, m, 0, 0, n_threads, out, &fresh)) return 1;
    int32_t appended = 0;
    for (size_t f = 0; f < m->specs.size(); ++f) {      // no feature keeps an empty table
        const int ord = m->dpd_ordinal[f];
        if (ord >= 0 && m->specs[f].dim == 0 && fresh[ord].empty()) fresh[ord].push_back(0);
    }
    for (const auto &list : fresh) appended += (int32_t)list.size();
    if (appended && model_extend(m, fresh, seed)) { lbio_rows_free(out); return 1; }
    if (n_new) *n_new = appended;
    return 0;
}

extern "C" int lbio_model_adopt_dictionaries(lbio_model *dst, const lbio_model *src, uint64_t seed) {
    if (dst->specs.size() != src->specs.size() || dst->dpd_lookup.size() != src->dpd_lookup.size())
        return fail("adopt_dictionaries: the models have different schemas");
    std::vector<std::vector<uint32_t>> fresh(src->dpd_lookup.size());
    for (size_t ord = 0; ord < fresh.size(); ++ord) {
        const int a = src->dpd_offsets[ord], b = src->dpd_offsets[ord + 1];
        const auto &have = dst->dpd_lookup[ord];
        for (int i = a; i < b; ++i) {
            const uint32_t v = src->dpd_values[i];
            auto it = std::lower_bound(have.begin(), have.end(), std::make_pair(v, 0u));
            if (it == have.end() || it->first != v) fresh[ord].push_back(v);
        }
    }
    if (model_extend(dst, fresh, seed)) return 1;
    if (dst->dpd_values != src->dpd_values || dst->dpd_offsets != src->dpd_offsets)
        return fail("adopt_dictionaries: the dictionaries differ in more than missing values");
    return 0;
}

extern "C" void lbio_rows_free(lbio_row_block *b) {
    if (b && b->owner) delete (RowsOwner *)b->owner;
    if (b) std::memset(b, 0, sizeof(*b));
}

extern "C" int lbio_stream_stats(const char *path, uint64_t *message_count, uint32_t *max_message_size) {
    pbw::InFile file(path);
    if (!file.is_open()) return fail("failed to open input file %s", path);
    bool bad = false;
    uint32_t size = 0, biggest = 0;
    while (file.skip_stream(size, bad)) biggest = std::max(biggest, size);
    if (bad) return fail("failed to count %s", path);
    if (message_count) *message_count = file.position();
    if (max_message_size) *max_message_size = biggest;
    return 0;
}

// shuffle_stream (src/shuffle.hpp:38-86): message j of the input goes to position index[j] of the output, index a
// seeded permutation; the output is produced in chunks of at most target_mem_bytes, one pass over the input each,
// so the result does not depend on the chunk size.  (The permutation comes from SplitMix64 + Fisher-Yates, not from
// the reference's std::shuffle(rng_t): same law, different draws.)
extern "C" int lbio_sorted_groupids(const uint32_t *counts, int32_t n_groups, uint32_t *order_out, int32_t *n_out) {
    std::vector<uint32_t> order;
    for (int32_t g = 0; g < n_groups; ++g)
        if (counts[g]) order.push_back((uint32_t)g);
    // std::sort, not stable_sort: the reference's own call (src/cross_cat.cc:228-231), so that tied groups land where its do
    std::sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return counts[x] > counts[y]; });
    std::copy(order.begin(), order.end(), order_out);
    *n_out = (int32_t)order.size();
    return 0;
}

extern "C" int lbio_shuffle_stream(const char *messages_in, const char *shuffled_out, uint64_t seed, double target_mem_bytes) {
    if (std::string(messages_in) == std::string(shuffled_out)) return fail("cannot shuffle file in-place: %s", messages_in);
    if (std::string(messages_in) == "-" || std::string(messages_in) == "-.gz") return fail("shuffle input is not a file: %s", messages_in);
    if (!(target_mem_bytes > 0)) return fail("expected 0 < target_mem_bytes");
    uint64_t message_count = 0;
    uint32_t max_message_size = 0;
    if (lbio_stream_stats(messages_in, &message_count, &max_message_size)) return 1;
    if (message_count > 0xFFFFFFFFull) return fail("too many messages to shuffle: %llu", (unsigned long long)message_count);
    const double index_bytes = 4.0 * (double)message_count;
    const double target = std::max(1.0, std::min((double)message_count, (target_mem_bytes - index_bytes) / std::max<uint32_t>(max_message_size, 1)));
    const size_t chunk_size = (size_t)std::llround(target);
    std::vector<uint32_t> index(message_count);
    for (size_t i = 0; i < message_count; ++i) index[i] = (uint32_t)i;
    uint64_t state = seed;
    auto next = [&]() {
        uint64_t z = (state += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    };
    for (size_t i = message_count; i > 1; --i) std::swap(index[i - 1], index[(size_t)(next() % i)]);
    pbw::OutFile shuffled(shuffled_out);
    if (!shuffled.is_open()) return fail("failed to open output file %s", shuffled_out);
    std::vector<std::vector<uint8_t>> chunk;
    std::vector<uint8_t> message;
    for (size_t begin = 0; begin < message_count; begin += chunk_size) {
        const size_t end = std::min(begin + chunk_size, (size_t)message_count);
        chunk.assign(end - begin, std::vector<uint8_t>());
        pbw::InFile messages(messages_in);
        if (!messages.is_open()) return fail("failed to open input file %s", messages_in);
        bool bad = false;
        for (size_t j = 0; j < message_count; ++j) {
            if (!messages.try_read_stream(message, bad)) return fail("failed to parse message from %s", messages_in);
            const size_t i = index[j];
            if (begin <= i && i < end) chunk[i - begin].swap(message);
        }
        for (const auto &m : chunk)
            if (!shuffled.write_stream(m)) return fail("failed to write %s", shuffled_out);
    }
    if (!shuffled.close()) return fail("failed to write %s", shuffled_out);
    return 0;
}

// One ProductValue from (feature, raw value) cells; sparsity chosen as ValueSchema::normalize_small does
// (src/product_value.hpp:396-451: none / all / sparse below 10 % / dense).
static void encode_value(Writer &w, uint32_t field, const lbio_model &m, const uint32_t *features, const uint32_t *values,
                         size_t n) {
    const size_t F = m.specs.size();
    Writer v, o;
    if (n == 0) o.u64(1, SP_NONE);
    else if (n == F) o.u64(1, SP_ALL);
    else if ((float)n < 0.1f * (float)F) {
        o.u64(1, SP_SPARSE);
        for (size_t i = 0; i < n; ++i) o.u64(3, features[i]);
    } else {
        o.u64(1, SP_DENSE);
        size_t i = 0;
        for (size_t p = 0; p < F; ++p) {
            const bool on = i < n && features[i] == p;
            o.boolean(2, on);
            i += on;
        }
    }
    v.msg(1, o);
    for (int pass = 0; pass < 3; ++pass)
        for (size_t i = 0; i < n; ++i) {
            const uint32_t p = features[i];
            const lb_feature_spec &s = m.specs[p];
            const uint32_t raw = values ? values[i] : 0;
            const int cls = s.type == LB_BB ? 0 : s.type == LB_NICH ? 2 : 1;
            if (cls != pass) continue;
            if (pass == 0) v.boolean(2, raw != 0);
            else if (pass == 1) {
                uint32_t x = raw;
                if (s.type == LB_DPD && raw != LB_DPD_OTHER)     // OTHER() of distributions' DPD is the same all-ones sentinel
                    x = m.dpd_values[m.dpd_offsets[m.dpd_ordinal[p]] + std::min<uint32_t>(raw, (uint32_t)s.dim - 1)];
                v.u64(3, x);
            } else { float x; std::memcpy(&x, &raw, 4); v.f32(4, x); }
        }
    w.msg(field, v);
}

// A rows stream written block by block (Differ::compress_rows writes as it reads, src/differ.cc:166-181)
struct lbio_rows_writer {
    pbw::OutFile file;
    std::string path;
    explicit lbio_rows_writer(const char *p) : file(p), path(p) {}
};

extern "C" int lbio_rows_writer_open(const char *path, lbio_rows_writer **out) {
    auto *w = new lbio_rows_writer(path);
    if (!w->file.is_open()) { delete w; return fail("failed to open output file %s", path); }
    *out = w;
    return 0;
}

extern "C" int lbio_rows_writer_write(lbio_rows_writer *w, const lbio_model *m, const lbio_row_block *rows,
                                      const uint32_t *tare_values) {
    Writer row, diff;
    std::vector<uint32_t> neg_values;
    for (uint64_t r = 0; r < rows->n_rows; ++r) {
        row.clear();
        diff.clear();
        const uint64_t p0 = rows->pos_ptr[r], p1 = rows->pos_ptr[r + 1];
        encode_value(diff, 1, *m, rows->pos_feature + p0, rows->pos_value + p0, (size_t)(p1 - p0));
        const uint64_t n0 = rows->neg_ptr ? rows->neg_ptr[r] : 0, n1 = rows->neg_ptr ? rows->neg_ptr[r + 1] : 0;
        // neg carries the tare's values in loom (_abs_to_rel_type, src/differ.cc:277-286); consumers only use its
        // observed positions (remove_diff re-reads the tare), so without a tare at hand zeros are written
        const uint32_t *nv = nullptr;
        if (tare_values && n1 > n0) {
            neg_values.clear();
            for (uint64_t i = n0; i < n1; ++i) neg_values.push_back(tare_values[rows->neg_feature[i]]);
            nv = neg_values.data();
        }
        encode_value(diff, 2, *m, rows->neg_feature + n0, nv, (size_t)(n1 - n0));
        if (rows->row_has_tare && rows->row_has_tare[r]) diff.u64(3, 0);
        row.u64(1, rows->rowids ? rows->rowids[r] : r);
        row.msg(2, diff);
        if (!w->file.write_stream(row.buf)) return fail("failed to write %s", w->path.c_str());
    }
    return 0;
}

extern "C" int lbio_rows_writer_close(lbio_rows_writer *w) {
    const bool ok = w->file.close();
    const std::string path = w->path;
    delete w;
    return ok ? 0 : fail("failed to write %s", path.c_str());
}

extern "C" int lbio_rows_write(const char *path, const lbio_model *m, const lbio_row_block *rows) {
    lbio_rows_writer *w = nullptr;
    if (lbio_rows_writer_open(path, &w)) return 1;
    if (lbio_rows_writer_write(w, m, rows, nullptr)) { lbio_rows_writer_close(w); return 1; }
    return lbio_rows_writer_close(w);
}

extern "C" int lbio_rows_to_columns(const lbio_model *m, const lbio_row_block *rows, const uint8_t *tare_observed,
                                    const uint32_t *tare_values, uint8_t *cat8, uint32_t *wide, uint32_t *mask) {
    const size_t F = m->specs.size();
    const uint64_t R = rows->n_rows, W = (R + 31) / 32;
    size_t F8 = 0;
    for (size_t f = 0; f < F; ++f) F8 += m->specs[f].type == LB_BB || m->specs[f].type == LB_DD;
    std::memset(cat8, 0, F8 * R);
    std::memset(wide, 0, (F - F8) * R * 4);
    std::memset(mask, 0, F * W * 4);
    auto set = [&](size_t f, uint64_t r, uint32_t raw) {
        if (f < F8) cat8[f * R + r] = (uint8_t)raw;
        else wide[(f - F8) * R + r] = raw;
        mask[f * W + (r >> 5)] |= 1u << (r & 31);
    };
    for (uint64_t r = 0; r < R; ++r) {
        if (rows->row_has_tare && rows->row_has_tare[r]) {
            if (!tare_observed) return fail("row %llu references a tare but no tares were given", (unsigned long long)r);
            for (size_t f = 0; f < F; ++f)
                if (tare_observed[f]) set(f, r, tare_values[f]);
        }
        for (uint64_t i = rows->neg_ptr[r]; i < rows->neg_ptr[r + 1]; ++i) {
            const size_t f = rows->neg_feature[i];
            mask[f * W + (r >> 5)] &= ~(1u << (r & 31));
            if (f < F8) cat8[f * R + r] = 0;
            else wide[(f - F8) * R + r] = 0;
        }
        for (uint64_t i = rows->pos_ptr[r]; i < rows->pos_ptr[r + 1]; ++i) set(rows->pos_feature[i], r, rows->pos_value[i]);
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Tares (CrossCat::tares_load src/cross_cat.cc:137-146)
extern "C" int lbio_tares_read(const char *path, const lbio_model *m, uint8_t *tare_observed, uint32_t *tare_values,
                               int32_t *n_tares) {
    const size_t F = m->specs.size();
    std::memset(tare_observed, 0, F);
    std::memset(tare_values, 0, F * 4);
    *n_tares = 0;
    pbw::InFile file(path);
    if (!file.is_open()) return fail("failed to open input file %s", path);
    std::vector<uint8_t> msg;
    bool bad = false;
    ValueScratch scratch;
    while (file.try_read_stream(msg, bad)) {
        if (*n_tares == 1) return fail("%s holds more than one tare row; one is supported", path);
        std::vector<uint32_t> features, values;
        if (const char *err = decode_value(Reader(msg), *m, scratch, features, &values)) return fail("%s: %s", path, err);
        for (size_t i = 0; i < features.size(); ++i) {
            tare_observed[features[i]] = 1;
            tare_values[features[i]] = values[i];
        }
        ++*n_tares;
    }
    if (bad) return fail("failed to parse message from %s (truncated stream)", path);
    return 0;
}

// Tares with OPEN dictionaries: a DPD value of the tare row that the dictionary lacks is appended the way
// lbio_rows_read_open appends the values of the rows (a tared column's own value is in no row's pos).
extern "C" int lbio_tares_read_open(const char *path, lbio_model *m, uint64_t seed, uint8_t *tare_observed,
                                    uint32_t *tare_values, int32_t *n_tares, int32_t *n_new) {
    if (n_new) *n_new = 0;
    {
        pbw::InFile file(path);
        if (!file.is_open()) return fail("failed to open input file %s", path);
        std::vector<uint8_t> msg;
        bool bad = false;
        ValueScratch scratch;
        std::vector<std::vector<uint32_t>> fresh(m->dpd_lookup.size());
        int32_t appended = 0;
        while (file.try_read_stream(msg, bad)) {
            std::vector<uint32_t> features, values;
            std::vector<uint64_t> misses;
            if (const char *err = decode_value(Reader(msg), *m, scratch, features, &values, &misses)) return fail("%s: %s", path, err);
            for (uint64_t at : misses) {
                auto &list = fresh[m->dpd_ordinal[features[at]]];
                if (std::find(list.begin(), list.end(), values[at]) == list.end()) { list.push_back(values[at]); ++appended; }
            }
        }
        if (bad) return fail("failed to parse message from %s (truncated stream)", path);
        if (appended && model_extend(m, fresh, seed)) return 1;
        if (n_new) *n_new = appended;
    }
    return lbio_tares_read(path, m, tare_observed, tare_values, n_tares);
}

extern "C" int lbio_tares_write(const char *path, const lbio_model *m, const uint8_t *tare_observed,
                                const uint32_t *tare_values) {
    pbw::OutFile file(path);
    if (!file.is_open()) return fail("failed to open output file %s", path);
    std::vector<uint32_t> features, values;
    for (size_t f = 0; f < m->specs.size(); ++f)
        if (tare_observed[f]) { features.push_back((uint32_t)f); values.push_back(tare_values[f]); }
    if (features.empty()) {      // "if (schema.total_size(tare)) tares.write_stream(tare)" (src/tare.cc:66-69): an empty stream
        if (!file.close()) return fail("failed to write %s", path);
        return 0;
    }
    Writer outer;
    encode_value(outer, 1, *m, features.data(), values.data(), features.size());
    // encode_value wrote a field-1 submessage; the stream holds the bare ProductValue
    Reader r(outer.buf);
    uint32_t f, wt;
    r.next(f, wt);
    Reader body = r.sub();
    std::vector<uint8_t> msg(body.p, body.end);
    if (!file.write_stream(msg) || !file.close()) return fail("failed to write %s", path);
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Assignments (src/assignments.cc:50-100)
struct AssignOwner { std::vector<uint64_t> rowids; std::vector<uint32_t> groupids; };

extern "C" int lbio_assign_read(const char *path, lbio_assign_block *out) {
    std::memset(out, 0, sizeof(*out));
    pbw::InFile file(path);
    if (!file.is_open()) return fail("failed to open input file %s", path);
    auto *o = new AssignOwner;
    std::vector<uint8_t> msg;
    std::vector<uint32_t> ids;
    bool bad = false;
    int K = -1;
    while (file.try_read_stream(msg, bad)) {
        Reader r(msg);
        uint32_t f, wt;
        uint64_t rowid = 0;
        ids.clear();
        while (r.next(f, wt)) {
            if (f == 1 && wt == pbw::VARINT) rowid = r.varint();
            else if (f == 2) r.rep_varint(wt, ids);
            else r.skip(wt);
        }
        if (!r.ok) { delete o; return fail("failed to parse message from %s", path); }
        if (K < 0) K = (int)ids.size();
        if ((int)ids.size() != K) { delete o; return fail("%s: assignments disagree on the kind count (%zu vs %d)", path, ids.size(), K); }
        o->rowids.push_back(rowid);
        o->groupids.insert(o->groupids.end(), ids.begin(), ids.end());
    }
    if (bad) { delete o; return fail("failed to parse message from %s (truncated stream)", path); }
    out->n_rows = o->rowids.size();
    out->n_kinds = std::max(K, 0);
    out->rowids = o->rowids.data();
    out->groupids = o->groupids.data();
    out->owner = o;
    return 0;
}

extern "C" void lbio_assign_free(lbio_assign_block *b) {
    if (b && b->owner) delete (AssignOwner *)b->owner;
    if (b) std::memset(b, 0, sizeof(*b));
}

extern "C" int lbio_assign_write(const char *path, uint64_t n_rows, int32_t n_kinds, const uint64_t *rowids,
                                 const uint32_t *groupids) {
    pbw::OutFile file(path);
    if (!file.is_open()) return fail("failed to open output file %s", path);
    Writer w;
    for (uint64_t r = 0; r < n_rows; ++r) {
        w.clear();
        w.u64(1, rowids[r]);
        for (int k = 0; k < n_kinds; ++k) w.u64(2, groupids[r * (uint64_t)n_kinds + k]);
        if (!file.write_stream(w.buf)) return fail("failed to write %s", path);
    }
    if (!file.close()) return fail("failed to write %s", path);
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Groups (ProductModel.Group :113-121)
static int spec_int_rows(const lb_feature_spec &s) {
    switch (s.type) {
    case LB_BB: case LB_GP: return 2;
    case LB_DD: case LB_DPD: return s.dim + 1;
    case LB_NICH: return 1;
    }
    return 0;
}
static int spec_flt_rows(const lb_feature_spec &s) { return s.type == LB_GP ? 1 : s.type == LB_NICH ? 2 : 0; }

struct GroupOwner { std::vector<uint32_t> counts, ints; std::vector<double> flts; };

struct GroupMsg {                    // one parsed ProductModel.Group: per-feature sub-messages by type
    uint64_t count = 0;
    std::vector<std::pair<const uint8_t *, size_t>> by_type[5];
};

extern "C" int lbio_groups_read(const char *path, const lbio_model *m, const uint32_t *featureids, int32_t nf,
                                lbio_group_block *out) {
    std::memset(out, 0, sizeof(*out));
    int n_int = 0, n_flt = 0;
    std::vector<int> int_off(nf), flt_off(nf);
    for (int j = 0; j < nf; ++j) {
        if (featureids[j] >= m->specs.size()) return fail("featureid out of bounds: %u", featureids[j]);
        int_off[j] = n_int; n_int += spec_int_rows(m->specs[featureids[j]]);
        flt_off[j] = n_flt; n_flt += spec_flt_rows(m->specs[featureids[j]]);
    }
    pbw::InFile file(path);
    if (!file.is_open()) return fail("failed to open input file %s", path);
    // groups arrive one message at a time; collect column-wise then transpose to [rows][G]
    std::vector<std::vector<uint32_t>> icol;   // per group: n_int values
    std::vector<std::vector<double>> fcol;
    std::vector<uint32_t> counts;
    std::vector<uint8_t> msg;
    bool bad = false;
    std::vector<uint64_t> tmp_u, keys, vals;
    while (file.try_read_stream(msg, bad)) {
        GroupMsg g;
        Reader r(msg);
        uint32_t f, wt;
        while (r.next(f, wt)) {
            if (f == 1 && wt == pbw::VARINT) g.count = r.varint();
            else if (f >= 2 && f <= 7 && wt == pbw::BYTES) {
                Reader s = r.sub();
                if (f == 6) return fail("%s: BetaNegativeBinomial groups are not supported", path);
                g.by_type[f == 7 ? LB_NICH : (int)f - 2].emplace_back(s.p, (size_t)(s.end - s.p));
            } else r.skip(wt);
        }
        if (!r.ok) return fail("failed to parse message from %s", path);
        std::vector<uint32_t> iv((size_t)n_int, 0);
        std::vector<double> fv((size_t)n_flt, 0.0);
        size_t next_of_type[5] = {0, 0, 0, 0, 0};
        for (int j = 0; j < nf; ++j) {
            const lb_feature_spec &s = m->specs[featureids[j]];
            if (next_of_type[s.type] >= g.by_type[s.type].size())
                return fail("%s: a group has fewer feature entries than the kind has features", path);
            const auto &span = g.by_type[s.type][next_of_type[s.type]++];
            Reader q(span.first, span.second);
            uint32_t *I = iv.data() + int_off[j];
            double *D = fv.data() + flt_off[j];
            switch (s.type) {
            case LB_BB:
                while (q.next(f, wt)) {
                    if (f == dist::BB_HEADS && wt == pbw::VARINT) I[0] = (uint32_t)q.varint();
                    else if (f == dist::BB_TAILS && wt == pbw::VARINT) I[1] = (uint32_t)q.varint();
                    else q.skip(wt);
                }
                break;
            case LB_DD: {
                tmp_u.clear();
                while (q.next(f, wt)) {
                    if (f == dist::DD_COUNTS) q.rep_varint(wt, tmp_u);
                    else q.skip(wt);
                }
                if ((int)tmp_u.size() != s.dim) return fail("%s: DirichletDiscrete group has %zu counts, model dim is %d", path, tmp_u.size(), s.dim);
                uint32_t total = 0;
                for (int v = 0; v < s.dim; ++v) { I[v] = (uint32_t)tmp_u[v]; total += I[v]; }
                I[s.dim] = total;
            } break;
            case LB_DPD: {
                keys.clear(); vals.clear();
                while (q.next(f, wt)) {
                    if (f == dist::DPD_KEYS) q.rep_varint(wt, keys);
                    else if (f == dist::DPD_GVALUES) q.rep_varint(wt, vals);
                    else q.skip(wt);
                }
                if (keys.size() != vals.size()) return fail("%s: DirichletProcessDiscrete group keys / values differ in length", path);
                const auto &lk = m->dpd_lookup[m->dpd_ordinal[featureids[j]]];
                uint32_t total = 0;
                for (size_t i = 0; i < keys.size(); ++i) {
                    auto it = std::lower_bound(lk.begin(), lk.end(), std::make_pair((uint32_t)keys[i], 0u));
                    if (it == lk.end() || it->first != (uint32_t)keys[i]) return fail("%s: DPD group key %llu is not in the model's dictionary", path, (unsigned long long)keys[i]);
                    I[it->second] += (uint32_t)vals[i];
                    total += (uint32_t)vals[i];
                }
                I[s.dim] = total;
            } break;
            case LB_GP:
                while (q.next(f, wt)) {
                    if (f == dist::GP_COUNT && wt == pbw::VARINT) I[0] = (uint32_t)q.varint();
                    else if (f == dist::GP_SUM && wt == pbw::VARINT) I[1] = (uint32_t)q.varint();
                    else if (f == dist::GP_LOG_PROD && wt == pbw::FIXED32) D[0] = q.f32();
                    else q.skip(wt);
                }
                break;
            case LB_NICH:
                while (q.next(f, wt)) {
                    if (f == dist::NICH_COUNT && wt == pbw::VARINT) I[0] = (uint32_t)q.varint();
                    else if (f == dist::NICH_MEAN && wt == pbw::FIXED32) D[0] = q.f32();
                    else if (f == dist::NICH_CTV && wt == pbw::FIXED32) D[1] = q.f32();
                    else q.skip(wt);
                }
                break;
            }
            if (!q.ok) return fail("failed to parse message from %s", path);
        }
        for (int t = 0; t < 5; ++t)
            if (next_of_type[t] != g.by_type[t].size())
                return fail("%s: a group has more feature entries than the kind has features", path);
        counts.push_back((uint32_t)g.count);
        icol.push_back(std::move(iv));
        fcol.push_back(std::move(fv));
    }
    if (bad) return fail("failed to parse message from %s (truncated stream)", path);
    auto *o = new GroupOwner;
    const size_t G = counts.size();
    o->counts = counts;
    o->ints.resize((size_t)n_int * G);
    o->flts.resize((size_t)n_flt * G);
    for (size_t g = 0; g < G; ++g) {
        for (int r = 0; r < n_int; ++r) o->ints[(size_t)r * G + g] = icol[g][r];
        for (int r = 0; r < n_flt; ++r) o->flts[(size_t)r * G + g] = fcol[g][r];
    }
    out->n_groups = (int32_t)G;
    out->n_int_rows = n_int;
    out->n_flt_rows = n_flt;
    out->counts = o->counts.data();
    out->ints = o->ints.data();
    out->flts = o->flts.data();
    out->owner = o;
    return 0;
}

extern "C" void lbio_groups_free(lbio_group_block *b) {
    if (b && b->owner) delete (GroupOwner *)b->owner;
    if (b) std::memset(b, 0, sizeof(*b));
}

extern "C" int lbio_groups_write(const char *path, const lbio_model *m, const uint32_t *featureids, int32_t nf,
                                 int32_t n_groups, const uint32_t *counts, const uint32_t *ints, const double *flts,
                                 const uint32_t *order, int32_t n_out) {
    pbw::OutFile file(path);
    if (!file.is_open()) return fail("failed to open output file %s", path);
    std::vector<int> int_off(nf), flt_off(nf);
    int n_int = 0, n_flt = 0;
    for (int j = 0; j < nf; ++j) {
        if (featureids[j] >= m->specs.size()) return fail("featureid out of bounds: %u", featureids[j]);
        int_off[j] = n_int; n_int += spec_int_rows(m->specs[featureids[j]]);
        flt_off[j] = n_flt; n_flt += spec_flt_rows(m->specs[featureids[j]]);
    }
    const size_t G = (size_t)n_groups;
    Writer w, s;
    for (int i = 0; i < n_out; ++i) {
        const size_t g = order ? order[i] : (size_t)i;
        if (g >= G) return fail("group id out of range: %zu", g);
        w.clear();
        w.u64(1, counts[g]);
        for (int j = 0; j < nf; ++j) {        // ascending feature id = field order bb, dd, dpd, gp, nich
            const lb_feature_spec &spec = m->specs[featureids[j]];
            auto I = [&](int row) { return ints[(size_t)(int_off[j] + row) * G + g]; };
            auto D = [&](int row) { return flts[(size_t)(flt_off[j] + row) * G + g]; };
            s.clear();
            switch (spec.type) {
            case LB_BB:
                s.u64(dist::BB_HEADS, I(0)); s.u64(dist::BB_TAILS, I(1));
                w.msg(2, s);
                break;
            case LB_DD:
                for (int v = 0; v < spec.dim; ++v) s.u64(dist::DD_COUNTS, I(v));
                w.msg(3, s);
                break;
            case LB_DPD: {
                const int at = m->dpd_offsets[m->dpd_ordinal[featureids[j]]];
                for (int v = 0; v < spec.dim; ++v)
                    if (I(v)) s.u64(dist::DPD_KEYS, m->dpd_values[at + v]);
                for (int v = 0; v < spec.dim; ++v)
                    if (I(v)) s.u64(dist::DPD_GVALUES, I(v));
                w.msg(4, s);
            } break;
            case LB_GP:
                s.u64(dist::GP_COUNT, I(0)); s.u64(dist::GP_SUM, I(1)); s.f32(dist::GP_LOG_PROD, (float)D(0));
                w.msg(5, s);
                break;
            case LB_NICH:
                s.u64(dist::NICH_COUNT, I(0)); s.f32(dist::NICH_MEAN, (float)D(0)); s.f32(dist::NICH_CTV, (float)D(1));
                w.msg(7, s);
                break;
            }
        }
        if (!file.write_stream(w.buf)) return fail("failed to write %s", path);
    }
    if (!file.close()) return fail("failed to write %s", path);
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Log (LogMessage :257-322)
struct lbio_log { pbw::OutFile *file; std::string path; };

extern "C" int lbio_log_open(const char *path, int32_t append, lbio_log **out) {
    auto *f = new pbw::OutFile(path, append != 0);
    if (!f->is_open()) { delete f; return fail("failed to open output file %s", path); }
    *out = new lbio_log{f, path};
    return 0;
}

extern "C" int lbio_log_write(lbio_log *log, const lbio_log_entry *e) {
    Writer w, rusage, args, summary, scores, status, s;
    struct timeval now;
    gettimeofday(&now, nullptr);
    w.u64(1, (uint64_t)now.tv_sec * 1000000ull + (uint64_t)now.tv_usec);
    struct rusage ru;
    getrusage(RUSAGE_SELF, &ru);
    rusage.u64(1, (uint64_t)ru.ru_maxrss);
    rusage.f64(2, ru.ru_utime.tv_sec + 1e-6 * ru.ru_utime.tv_usec);
    rusage.f64(3, ru.ru_stime.tv_sec + 1e-6 * ru.ru_stime.tv_usec);
    w.msg(2, rusage);
    args.u64(1, e->iter);
    write_py(summary, 1, e->topology_py[0], e->topology_py[1]);
    for (int k = 0; k < e->n_kinds; ++k) write_py(summary, 2, e->kind_py[2 * k], e->kind_py[2 * k + 1]);
    for (int k = 0; k < e->n_kinds; ++k) summary.u64(3, e->feature_counts[k]);
    for (int k = 0; k < e->n_kinds; ++k) summary.u64(4, e->category_counts[k]);
    args.msg(2, summary);
    scores.f32(1, e->score);
    scores.f32(2, e->kl_divergence);
    scores.u64(4, e->assigned_object_count);
    args.msg(3, scores);
    s.u64(1, e->cat_total_time_usec); status.msg(1, s); s.clear();
    s.u64(1, e->hyper_total_time_usec); status.msg(2, s); s.clear();
    s.u64(1, e->kind_total_count); s.u64(2, e->kind_change_count); s.u64(3, 0); s.u64(4, 0);
    s.u64(5, 0); s.u64(6, 0); s.u64(7, 0); s.u64(8, e->kind_total_time_usec);
    status.msg(3, s);
    args.msg(4, status);
    w.msg(3, args);
    if (!log->file->write_stream(w.buf) || !log->file->flush()) return fail("failed to write %s", log->path.c_str());
    return 0;
}

extern "C" void lbio_log_close(lbio_log *log) {
    if (!log) return;
    delete log->file;
    delete log;
}

// ---------------------------------------------------------------------------------------------
// PosteriorEnum.Sample (:326-338)
struct lbio_samples { pbw::OutFile *file; std::string path; };

extern "C" int lbio_samples_open(const char *path, lbio_samples **out) {
    auto *f = new pbw::OutFile(path);
    if (!f->is_open()) { delete f; return fail("failed to open output file %s", path); }
    *out = new lbio_samples{f, path};
    return 0;
}

extern "C" int lbio_samples_write(lbio_samples *s, int32_t n_kinds, const int32_t *kind_feature_ptr,
                                  const uint32_t *featureids, const int32_t *kind_group_ptr,
                                  const uint64_t *group_row_ptr, const uint32_t *rowids, float score) {
    Writer w, kind, group;
    for (int k = 0; k < n_kinds; ++k) {
        kind.clear();
        for (int i = kind_feature_ptr[k]; i < kind_feature_ptr[k + 1]; ++i) kind.u64(1, featureids[i]);
        for (int g = kind_group_ptr[k]; g < kind_group_ptr[k + 1]; ++g) {
            group.clear();
            for (uint64_t i = group_row_ptr[g]; i < group_row_ptr[g + 1]; ++i) group.u64(1, rowids[i]);
            kind.msg(2, group);
        }
        w.msg(1, kind);
    }
    w.f32(2, score);
    if (!s->file->write_stream(w.buf)) return fail("failed to write %s", s->path.c_str());
    return 0;
}

extern "C" int lbio_samples_close(lbio_samples *s) {
    if (!s) return 0;
    const bool ok = s->file->close();
    const std::string path = s->path;
    delete s->file;
    delete s;
    return ok ? 0 : fail("failed to write %s", path.c_str());
}

// ---------------------------------------------------------------------------------------------
// Query.Request / Query.Response (:342-418)
struct lbio_query {
    pbw::InFile *in;
    pbw::OutFile *out;
    std::string in_path, out_path, id;
    std::vector<uint8_t> msg;
    RowChunk chunk;
    RowsOwner rows;
    std::vector<uint32_t> to_sample, set_features;
    std::vector<int32_t> set_ptr;
    std::vector<std::string> invalid;
    std::vector<const char *> invalid_ptrs;
};

extern "C" int lbio_query_open(const char *requests_in, const char *responses_out, lbio_query **out) {
    auto *q = new lbio_query;
    q->in = new pbw::InFile(requests_in);
    q->out = new pbw::OutFile(responses_out);
    q->in_path = requests_in;
    q->out_path = responses_out;
    if (!q->in->is_open() || !q->out->is_open()) {
        const bool in_bad = !q->in->is_open();
        lbio_query_close(q);
        return in_bad ? fail("failed to open input file %s", requests_in) : fail("failed to open output file %s", responses_out);
    }
    *out = q;
    return 0;
}

extern "C" void lbio_query_close(lbio_query *q) {
    if (!q) return;
    delete q->in;
    delete q->out;
    delete q;
}

// ProductValue.Observed (:76-85) as ascending positions; false = ValueSchema::is_valid(observed) fails
// (src/product_value.hpp:277-298)
static bool observed_positions(Reader o, size_t F, std::vector<uint32_t> &out) {
    int sparsity = -1;
    std::vector<uint8_t> dense;
    std::vector<uint32_t> sparse;
    uint32_t f, wt;
    while (o.next(f, wt)) {
        if (f == 1 && wt == pbw::VARINT) sparsity = (int)o.varint();
        else if (f == 2) o.rep_varint(wt, dense);
        else if (f == 3) o.rep_varint(wt, sparse);
        else o.skip(wt);
    }
    if (!o.ok) return false;
    switch (sparsity) {
    case SP_ALL:
        if (!dense.empty() || !sparse.empty()) return false;
        for (size_t p = 0; p < F; ++p) out.push_back((uint32_t)p);
        return true;
    case SP_DENSE:
        if (dense.size() != F || !sparse.empty()) return false;
        for (size_t p = 0; p < F; ++p)
            if (dense[p]) out.push_back((uint32_t)p);
        return true;
    case SP_SPARSE:
        if (!dense.empty()) return false;
        for (size_t i = 0; i < sparse.size(); ++i) {
            if (sparse[i] >= F || (i && sparse[i - 1] >= sparse[i])) return false;
            out.push_back(sparse[i]);
        }
        return true;
    case SP_NONE:
        return dense.empty() && sparse.empty();
    }
    return false;
}

extern "C" int lbio_query_next(lbio_query *q, const lbio_model *m, lbio_query_request *req, int32_t *got) {
    std::memset(req, 0, sizeof(*req));
    *got = 0;
    bool bad = false;
    if (!q->in->try_read_stream(q->msg, bad)) {
        if (bad) return fail("failed to parse message from %s (truncated stream)", q->in_path.c_str());
        return 0;
    }
    *got = 1;
    q->id.clear();
    q->chunk.clear();
    q->to_sample.clear();
    q->set_features.clear();
    q->set_ptr.assign(1, 0);
    q->invalid.clear();
    req->sample_row = req->score_row = req->entropy_row = req->update_row = req->first_score_data_row = -1;
    const size_t F = m->specs.size();
    ValueScratch scratch;
    std::vector<uint32_t> tares;
    auto add_row = [&](Reader diff, const char *what) -> int {      // row index, or -1 with `invalid` extended
        const char *err = decode_diff(diff, *m, scratch, tares, q->chunk.rowids.size(), q->chunk, false);
        if (err) {      // validate(): "...data" for a value that does not fit the schema, "...data.tares" for a tare id out of range
            const bool tares_suffix = err == BAD_TARE && std::strncmp(what, "score_derivative", 16) != 0;
            q->invalid.push_back(std::string("invalid request.") + what + (tares_suffix ? ".tares" : ""));
            return -1;
        }
        return (int)q->chunk.rowids.size() - 1;
    };
    Reader r(q->msg);
    uint32_t f, wt;
    while (r.next(f, wt)) {
        if (f == 1 && wt == pbw::BYTES) { Reader b = r.sub(); q->id.assign((const char *)b.p, (size_t)(b.end - b.p)); }
        else if (f == 2 && wt == pbw::BYTES) {                    // Sample.Request
            Reader s = r.sub();
            Reader data(nullptr, 0), to_sample(nullptr, 0);
            bool has_data = false, has_to_sample = false;
            while (s.next(f, wt)) {
                if (f == 1 && wt == pbw::BYTES) { data = s.sub(); has_data = true; }
                else if (f == 2 && wt == pbw::BYTES) { to_sample = s.sub(); has_to_sample = true; }
                else if (f == 3 && wt == pbw::VARINT) req->sample_count = (uint32_t)s.varint();
                else s.skip(wt);
            }
            if (!s.ok || !has_data) { q->invalid.push_back("invalid request.sample.data"); continue; }
            const int row = add_row(data, "sample.data");
            if (row < 0) continue;
            if (!has_to_sample || !observed_positions(to_sample, F, q->to_sample)) {
                q->to_sample.clear();
                q->invalid.push_back("invalid request.sample.to_sample");
                continue;
            }
            req->sample_row = row;
        } else if (f == 3 && wt == pbw::BYTES) {                  // Score.Request
            Reader s = r.sub();
            Reader data(nullptr, 0);
            bool has_data = false;
            while (s.next(f, wt)) {
                if (f == 1 && wt == pbw::BYTES) { data = s.sub(); has_data = true; }
                else s.skip(wt);
            }
            if (!s.ok || !has_data) { q->invalid.push_back("invalid request.score.data"); continue; }
            req->score_row = add_row(data, "score.data");
        } else if (f == 4 && wt == pbw::BYTES) {                  // Entropy.Request
            Reader s = r.sub();
            std::vector<Reader> row_sets, col_sets;
            Reader conditional(nullptr, 0);
            bool has_conditional = false;
            uint32_t sample_count = 0;
            while (s.next(f, wt)) {
                if (f == 1 && wt == pbw::BYTES) row_sets.push_back(s.sub());
                else if (f == 2 && wt == pbw::BYTES) col_sets.push_back(s.sub());
                else if (f == 3 && wt == pbw::BYTES) { conditional = s.sub(); has_conditional = true; }
                else if (f == 4 && wt == pbw::VARINT) sample_count = (uint32_t)s.varint();
                else s.skip(wt);
            }
            if (!s.ok || !has_conditional) { q->invalid.push_back("invalid request.entropy.conditional"); continue; }
            const int row = add_row(conditional, "entropy.conditional");
            if (row < 0) continue;
            bool ok = true;
            for (int pass = 0; pass < 2 && ok; ++pass)
                for (Reader &o : pass ? col_sets : row_sets) {
                    if (!observed_positions(o, F, q->set_features)) {
                        q->invalid.push_back(pass ? "invalid request.entropy.col_sets" : "invalid request.entropy.row_sets");
                        ok = false;
                        break;
                    }
                    q->set_ptr.push_back((int32_t)q->set_features.size());
                }
            if (ok && sample_count <= 1) { q->invalid.push_back("invalid request.entropy.sample_count"); ok = false; }
            if (!ok) { q->set_features.clear(); q->set_ptr.assign(1, 0); continue; }
            req->entropy_row = row;
            req->n_row_sets = (int32_t)row_sets.size();
            req->n_col_sets = (int32_t)col_sets.size();
            req->entropy_sample_count = sample_count;
        } else if (f == 5 && wt == pbw::BYTES) {                  // ScoreDerivative.Request
            Reader s = r.sub();
            std::vector<Reader> score_data;
            Reader update(nullptr, 0);
            bool has_update = false;
            while (s.next(f, wt)) {
                if (f == 1 && wt == pbw::BYTES) score_data.push_back(s.sub());
                else if (f == 2 && wt == pbw::BYTES) { update = s.sub(); has_update = true; }
                else if (f == 3 && wt == pbw::VARINT) req->row_limit = (uint32_t)s.varint();
                else s.skip(wt);
            }
            if (!s.ok || !has_update) { q->invalid.push_back("invalid request.score_derivative.update_data"); continue; }
            const int row = add_row(update, "score_derivative.update_data");
            if (row < 0) continue;
            bool ok = true;
            const int first = (int)q->chunk.rowids.size();
            for (Reader &d : score_data) ok = ok && add_row(d, "score_derivative.score_data") >= 0;
            if (!ok) continue;
            req->update_row = row;
            req->first_score_data_row = first;
            req->n_score_data = (int32_t)score_data.size();
        } else r.skip(wt);
    }
    if (!r.ok) return fail("failed to parse message from %s", q->in_path.c_str());
    // publish the rows
    RowsOwner &o = q->rows;
    o.rowids = q->chunk.rowids;
    o.has_tare = q->chunk.has_tare;
    o.pos_feature = q->chunk.pos_feature;
    o.pos_value = q->chunk.pos_value;
    o.neg_feature = q->chunk.neg_feature;
    o.pos_ptr.assign(1, 0);
    o.neg_ptr.assign(1, 0);
    for (uint32_t k : q->chunk.pos_count) o.pos_ptr.push_back(o.pos_ptr.back() + k);
    for (uint32_t k : q->chunk.neg_count) o.neg_ptr.push_back(o.neg_ptr.back() + k);
    publish_rows(&o, &req->rows);
    req->rows.owner = nullptr;          // owned by the query object
    req->id = q->id.c_str();
    req->to_sample = q->to_sample.data();
    req->n_to_sample = (int32_t)q->to_sample.size();
    req->set_ptr = q->set_ptr.data();
    req->set_features = q->set_features.data();
    q->invalid_ptrs.clear();
    for (const auto &e : q->invalid) q->invalid_ptrs.push_back(e.c_str());
    req->n_invalid = (int32_t)q->invalid_ptrs.size();
    req->invalid = q->invalid_ptrs.data();
    return 0;
}

extern "C" int lbio_query_respond(lbio_query *q, const lbio_model *m, const lbio_query_response *resp) {
    Writer w, sub, diff;
    w.tag(1, pbw::BYTES);
    w.varint(q->id.size());
    w.raw(q->id.data(), q->id.size());
    for (int i = 0; i < resp->n_errors; ++i) {
        const size_t n = std::strlen(resp->errors[i]);
        w.tag(2, pbw::BYTES);
        w.varint(n);
        w.raw(resp->errors[i], n);
    }
    if (resp->has_sample) {
        sub.clear();
        for (int i = 0; i < resp->n_samples; ++i) {
            diff.clear();
            const uint64_t a = resp->sample_ptr[i], b = resp->sample_ptr[i + 1];
            encode_value(diff, 1, *m, resp->sample_feature + a, resp->sample_value + a, (size_t)(b - a));
            encode_value(diff, 2, *m, nullptr, nullptr, 0);
            sub.msg(1, diff);
        }
        w.msg(3, sub);
    }
    if (resp->has_score) { sub.clear(); sub.f32(1, resp->score); w.msg(4, sub); }
    if (resp->has_entropy) {
        sub.clear();
        for (int i = 0; i < resp->n_entropy; ++i) sub.f32(1, resp->means[i]);
        for (int i = 0; i < resp->n_entropy; ++i) sub.f32(2, resp->variances[i]);
        w.msg(5, sub);
    }
    if (resp->has_score_derivative) {
        sub.clear();
        for (int i = 0; i < resp->n_ids; ++i) sub.u64(1, resp->ids[i]);
        for (int i = 0; i < resp->n_ids; ++i) sub.f32(2, resp->score_diffs[i]);
        w.msg(6, sub);
    }
    if (!q->out->write_stream(w.buf) || !q->out->flush()) return fail("failed to write %s", q->out_path.c_str());
    return 0;
}
