// query_server.hpp -- loom::QueryServer (src/query_server.hpp:36-107, src/query_server.cc) and the
// part of loom::MultiLoom it needs (src/multi_loom.cc:36-78), re-hosted on the C ABIs of
// include/loom_b200.h (device engine) and include/loom_b200_io.h (loom's files).
//
// Same protocol: a stream of Query.Request in, one Query.Response per request out, flushed after
// each (the client waits on it).  Every latent sample of the store is one device engine holding
// the sample's groups; a request's rows are decoded into one block that all latents share, and the
// four calls become batch calls on it:
//   Score            lb_query_score of the row on every latent, log-mean-exp over latents
//   Sample           latent ~ posterior weight of the conditioning row, then lb_query_sample
//   Entropy          the reference scores each sample against each feature set with a
//                    RestrictionScorer (src/scorer.cc); here the (sample, set) pairs become rows
//                    "conditional + the sample restricted to the set" of ONE block and are scored
//                    by lb_query_score in one pass per latent -- the same quantity,
//                    -log p(x_S | conditional), averaged over samples
//   ScoreDerivative  Score of every row before and after CatKernel::add_row of the update row on
//                    every latent (lb_cat_sweep), scaled by the row count, sorted, truncated
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <string>
#include <sys/stat.h>

#include "loom.hpp"

namespace loom {
namespace b200 {

class QueryServer {
public:
    // MultiLoom(root_in, load_groups = true, load_assign = false, load_tares = true) + QueryServer(cross_cats, config, rows_in)
    // (src/query.cc:66-76); store layout of src/store.hpp:68-89
    QueryServer(const std::string &root_in, const lbio_config &config, int device = 0)
        : rows_in_(root_in + "/ingest/diffs.pbs.gz"), rng_(config.seed) {
        const std::string tares_in = root_in + "/ingest/tares.pbs.gz";
        for (size_t seed = 0;; ++seed) {
            const std::string sample_root = root_in + "/samples/sample." + std::to_string(seed);
            struct stat st;
            if (stat(sample_root.c_str(), &st) != 0) break;
            latents_.emplace_back(new Latent(device));
            Latent &l = *latents_.back();
            lbio_config sample_config;
            LOOM_B200_IO(lbio_config_read((sample_root + "/config.pb.gz").c_str(), &sample_config));
            LOOM_B200_IO(lbio_model_read((sample_root + "/model.pb.gz").c_str(), &l.model));
            LOOM_B200_IO(lbio_model_get(l.model, &l.view));
            if (latents_.size() > 1) adopt_dictionary_order(l, latents_[0]->view);
            if (sample_config.cat_empty_group_count < 1) LOOM_B200_ERROR("expected 0 < empty_group_count");
            l.cross_cat.model_load(l.view.n_features, l.view.specs, l.view.aux, l.view.n_aux, l.view.n_kinds,
                                   l.view.featureid_to_kindid, l.view.kind_py, l.view.topology_py,
                                   (int)sample_config.cat_empty_group_count);
            l.groups_in = sample_root + "/groups";
            groups_load(l);
        }
        if (latents_.empty()) LOOM_B200_ERROR("no samples were found at %s", root_in.c_str());
        const lbio_model_view &v0 = latents_[0]->view;
        for (auto &l : latents_) {   // one schema, one value dictionary: the latents decode the same request bytes
            const lbio_model_view &v = l->view;
            bool same = v.n_features == v0.n_features && v.n_dpd == v0.n_dpd;
            for (int f = 0; same && f < v.n_features; ++f)
                same = v.specs[f].type == v0.specs[f].type && v.specs[f].dim == v0.specs[f].dim;
            for (int i = 0; same && i <= v.n_dpd; ++i) same = !v.n_dpd || v.dpd_offsets[i] == v0.dpd_offsets[i];
            for (int i = 0; same && v.n_dpd && i < v.dpd_offsets[v.n_dpd]; ++i) same = v.dpd_values[i] == v0.dpd_values[i];
            if (!same) LOOM_B200_ERROR("samples disagree on the schema or on a DirichletProcessDiscrete dictionary");
        }
        const int F = v0.n_features;
        tare_observed_.assign(F, 0);
        tare_values_.assign(F, 0);
        if (std::ifstream(tares_in)) {
            int32_t n_tares = 0;
            LOOM_B200_IO(lbio_tares_read(tares_in.c_str(), latents_[0]->model, tare_observed_.data(), tare_values_.data(), &n_tares));
            tare_count_ = n_tares;
        }
    }

    size_t latent_count() const { return latents_.size(); }

    // QueryServer::serve (src/query_server.cc:36-66)
    void serve(const char *requests_in, const char *responses_out) {
        lbio_query *q = nullptr;
        LOOM_B200_IO(lbio_query_open(requests_in, responses_out, &q));
        const lbio_model *schema = latents_[0]->model;
        for (;;) {
            lbio_query_request req;
            int32_t got = 0;
            LOOM_B200_IO(lbio_query_next(q, schema, &req, &got));
            if (!got) break;
            Answer a;
            for (int i = 0; i < req.n_invalid; ++i) a.errors.push_back(req.invalid[i]);
            Rows request_rows;
            request_rows.append(req.rows, 0, req.rows.n_rows);
            const std::vector<uint8_t> inconsistent = sanitize(request_rows);
            // validate(): "for (auto id : data.tares()) if (id >= tares().size())" (src/query_server.cc:77-82, 178-183, 241-246,
            // 421-441) -- the codec rejects ids above 0; a store without a tare row has no tare 0 either
            auto tare_missing = [&](int32_t row) { return row >= 0 && !tare_count_ && request_rows.has_tare[(size_t)row]; };
            // a diff whose neg / pos do not fit the tare it references (the reference would add and subtract the
            // likelihoods it names, which is not a probability; here the data is reported invalid)
            auto bad = [&](int32_t row) { return row >= 0 && inconsistent[(size_t)row]; };
            if (bad(req.sample_row)) { a.errors.push_back("invalid request.sample.data"); req.sample_row = -1; }
            if (bad(req.score_row)) { a.errors.push_back("invalid request.score.data"); req.score_row = -1; }
            if (bad(req.entropy_row)) { a.errors.push_back("invalid request.entropy.conditional"); req.entropy_row = -1; }
            if (bad(req.update_row)) { a.errors.push_back("invalid request.score_derivative.update_data"); req.update_row = -1; }
            for (int i = 0; req.update_row >= 0 && i < req.n_score_data; ++i)
                if (bad(req.first_score_data_row + i)) {
                    a.errors.push_back("invalid request.score_derivative.score_data");
                    req.update_row = -1;
                }
            if (tare_missing(req.sample_row)) { a.errors.push_back("invalid request.sample.data.tares"); req.sample_row = -1; }
            if (tare_missing(req.score_row)) { a.errors.push_back("invalid request.score.data.tares"); req.score_row = -1; }
            if (tare_missing(req.entropy_row)) { a.errors.push_back("invalid request.entropy.conditional.tares"); req.entropy_row = -1; }
            if (tare_missing(req.update_row)) { a.errors.push_back("invalid request.score_derivative.update_data"); req.update_row = -1; }
            for (int i = 0; req.update_row >= 0 && i < req.n_score_data; ++i)
                if (tare_missing(req.first_score_data_row + i)) {
                    a.errors.push_back("invalid request.score_derivative.score_data");
                    req.update_row = -1;
                }
            if (req.sample_row >= 0) call_sample(request_rows, (uint64_t)req.sample_row, req.to_sample, req.n_to_sample, req.sample_count, a);
            if (req.score_row >= 0) {
                a.has_score = true;
                a.score = call_score(request_rows, (uint64_t)req.score_row, 1)[0];
            }
            if (req.entropy_row >= 0) call_entropy(request_rows, req, a);
            if (req.update_row >= 0) call_score_derivative(request_rows, req, a);
            respond(q, schema, a);
        }
        lbio_query_close(q);
    }

private:
    struct Latent {
        explicit Latent(int device) : cross_cat(device) {}
        ~Latent() {
            lbio_model_free(model);
            if (file_model) lbio_model_free(file_model);
        }
        CrossCat cross_cat;
        lbio_model *model = nullptr;          // the model the engine runs on (dictionaries in latent 0's order)
        lbio_model *file_model = nullptr;     // the model as the sample's files have it, when the order differs
        lbio_model_view view;
        std::vector<std::vector<int32_t>> dictionary_perm;   // per DPD feature: position in the files of the value at index i
        std::string groups_in;
    };

    // Samples of one dataset see its rows in different (shuffled) orders (loom/tasks.py:228-233), so their
    // DirichletProcessDiscrete dictionaries list the same values in different orders.  The latents share one decoded
    // row block, so every latent is put into latent 0's order: betas and the groups' count rows are permuted at load.
    static void adopt_dictionary_order(Latent &l, const lbio_model_view &v0) {
        const lbio_model_view v = l.view;
        if (v.n_dpd != v0.n_dpd || v.n_features != v0.n_features)
            LOOM_B200_ERROR("samples disagree on the schema or on a DirichletProcessDiscrete dictionary");
        bool differs = false;
        std::vector<std::vector<int32_t>> perm((size_t)v.n_dpd);
        for (int ord = 0; ord < v.n_dpd; ++ord) {
            const int32_t a = v.dpd_offsets[ord], n = v.dpd_offsets[ord + 1] - a, a0 = v0.dpd_offsets[ord];
            if (n != v0.dpd_offsets[ord + 1] - a0) LOOM_B200_ERROR("samples disagree on the schema or on a DirichletProcessDiscrete dictionary");
            perm[ord].assign((size_t)n, -1);
            for (int32_t i0 = 0; i0 < n; ++i0) {
                for (int32_t i = 0; i < n; ++i)
                    if (v.dpd_values[a + i] == v0.dpd_values[a0 + i0]) perm[ord][i0] = i;
                if (perm[ord][i0] < 0) LOOM_B200_ERROR("samples disagree on the schema or on a DirichletProcessDiscrete dictionary");
                differs = differs || perm[ord][i0] != i0;
            }
        }
        if (!differs) return;
        std::vector<float> aux(v.aux, v.aux + v.n_aux);
        int ord = 0;
        for (int f = 0; f < v.n_features; ++f) {
            if (v.specs[f].type != LB_DPD) continue;
            for (int32_t i0 = 0; i0 < v.specs[f].dim; ++i0) aux[v.specs[f].aux_off + i0] = v.aux[v.specs[f].aux_off + perm[ord][i0]];
            ++ord;
        }
        lbio_model *canonical = nullptr;
        LOOM_B200_IO(lbio_model_create(v.n_features, v.specs, aux.data(), v.n_aux, v.n_kinds, v.featureid_to_kindid, v.kind_py,
                                       v.topology_py, v.has_hyper_prior ? &v.hyper_prior : nullptr, v0.dpd_offsets, v0.dpd_values,
                                       &canonical));
        l.file_model = l.model;
        l.model = canonical;
        l.dictionary_perm = perm;
        LOOM_B200_IO(lbio_model_get(l.model, &l.view));
    }

    // a block of rows in the CSR form lb_set_rows_diff takes
    struct Rows {
        std::vector<uint8_t> has_tare;
        std::vector<uint64_t> pos_ptr{0}, neg_ptr{0};
        std::vector<uint32_t> pos_feature, pos_value, neg_feature;
        uint64_t size() const { return has_tare.size(); }
        void append(const lbio_row_block &b, uint64_t begin, uint64_t end) {
            for (uint64_t r = begin; r < end; ++r) {
                has_tare.push_back(b.row_has_tare ? b.row_has_tare[r] : 0);
                pos_feature.insert(pos_feature.end(), b.pos_feature + b.pos_ptr[r], b.pos_feature + b.pos_ptr[r + 1]);
                pos_value.insert(pos_value.end(), b.pos_value + b.pos_ptr[r], b.pos_value + b.pos_ptr[r + 1]);
                neg_feature.insert(neg_feature.end(), b.neg_feature + b.neg_ptr[r], b.neg_feature + b.neg_ptr[r + 1]);
                pos_ptr.push_back(pos_feature.size());
                neg_ptr.push_back(neg_feature.size());
            }
        }
        void append(const Rows &b, uint64_t r) {
            has_tare.push_back(b.has_tare[r]);
            pos_feature.insert(pos_feature.end(), b.pos_feature.begin() + b.pos_ptr[r], b.pos_feature.begin() + b.pos_ptr[r + 1]);
            pos_value.insert(pos_value.end(), b.pos_value.begin() + b.pos_ptr[r], b.pos_value.begin() + b.pos_ptr[r + 1]);
            neg_feature.insert(neg_feature.end(), b.neg_feature.begin() + b.neg_ptr[r], b.neg_feature.begin() + b.neg_ptr[r + 1]);
            pos_ptr.push_back(pos_feature.size());
            neg_ptr.push_back(neg_feature.size());
        }
    };

    struct Answer {
        std::vector<std::string> errors;
        bool has_sample = false, has_score = false, has_entropy = false, has_score_derivative = false;
        std::vector<uint64_t> sample_ptr{0};
        std::vector<uint32_t> sample_feature, sample_value;
        float score = 0.f;
        std::vector<float> means, variances;
        std::vector<uint64_t> ids;
        std::vector<float> score_diffs;
    };

    // Rows as lb_set_rows_diff accepts them.  A row that references no tare is scored from its pos alone
    // (mixture.score_value(model, diff.pos(), ...) when !diff.tares_size(), src/query_server.cc:121-125, 216-222): its neg
    // is dropped.  A row that references the tare must remove only cells the tare holds and set only cells that are
    // free; returns 1 for the rows that do not.
    std::vector<uint8_t> sanitize(Rows &rows) const {
        const uint64_t R = rows.size();
        std::vector<uint8_t> inconsistent(R, 0);
        Rows out;
        std::vector<uint8_t> removed(tare_observed_.size());
        for (uint64_t r = 0; r < R; ++r) {
            const uint64_t n0 = rows.neg_ptr[r], n1 = rows.neg_ptr[r + 1], p0 = rows.pos_ptr[r], p1 = rows.pos_ptr[r + 1];
            out.has_tare.push_back(rows.has_tare[r]);
            bool ok = true;
            if (rows.has_tare[r] && tare_count_) {
                std::fill(removed.begin(), removed.end(), 0);
                for (uint64_t i = n0; i < n1 && ok; ++i) {
                    const uint32_t f = rows.neg_feature[i];
                    ok = tare_observed_[f] && !removed[f];
                    removed[f] = 1;
                }
                for (uint64_t i = p0; i < p1 && ok; ++i) {
                    const uint32_t f = rows.pos_feature[i];
                    ok = !tare_observed_[f] || removed[f];
                }
            }
            inconsistent[r] = !ok;
            if (ok) {
                out.pos_feature.insert(out.pos_feature.end(), rows.pos_feature.begin() + p0, rows.pos_feature.begin() + p1);
                out.pos_value.insert(out.pos_value.end(), rows.pos_value.begin() + p0, rows.pos_value.begin() + p1);
                if (rows.has_tare[r] && tare_count_)
                    out.neg_feature.insert(out.neg_feature.end(), rows.neg_feature.begin() + n0, rows.neg_feature.begin() + n1);
            }
            out.pos_ptr.push_back(out.pos_feature.size());
            out.neg_ptr.push_back(out.neg_feature.size());
        }
        rows = out;
        return inconsistent;
    }

    double uniform() { return (double)(rng_() >> 11) * (1.0 / 9007199254740992.0); }

    // the rows every latent reads: latent 0 decodes them on its device, the others borrow the block
    void load(const Rows &rows) {
        Latent &owner = *latents_[0];
        LOOM_B200_CHECK(LB_NAME(set_rows_diff)(owner.cross_cat.handle(), rows.size(), tare_observed_.data(), tare_values_.data(),
                                               rows.has_tare.data(), rows.pos_ptr.data(), rows.pos_feature.data(),
                                               rows.pos_value.data(), rows.neg_ptr.data(), rows.neg_feature.data()));
        owner.cross_cat.rows_attached(rows.size());
        for (size_t l = 1; l < latents_.size(); ++l) latents_[l]->cross_cat.rows_share(owner.cross_cat);
    }

    // CrossCat::mixture_load (src/cross_cat.cc:148-196) without assignments
    void groups_load(Latent &l) {
        for (int k = 0; k < l.view.n_kinds; ++k) {
            std::vector<uint32_t> fids;
            for (int f = 0; f < l.view.n_features; ++f)
                if ((int)l.view.featureid_to_kindid[f] == k) fids.push_back((uint32_t)f);
            const std::string path = l.groups_in + "/mixture." + std::to_string(k) + ".pbs.gz";
            lbio_group_block g;
            LOOM_B200_IO(lbio_groups_read(path.c_str(), l.file_model ? l.file_model : l.model, fids.data(), (int32_t)fids.size(), &g));
            if (l.file_model) {     // count rows of the DPD features into latent 0's dictionary order
                std::vector<uint32_t> rows;
                size_t first_row = 0;
                for (uint32_t f : fids) {
                    const lb_feature_spec &spec = l.view.specs[f];
                    const size_t n_rows = spec.type == LB_BB || spec.type == LB_GP ? 2 : spec.type == LB_NICH ? 1 : (size_t)spec.dim + 1;
                    if (spec.type == LB_DPD) {
                        int ord = 0;
                        for (uint32_t q = 0; q < f; ++q) ord += l.view.specs[q].type == LB_DPD;
                        const size_t G = (size_t)g.n_groups;
                        rows.assign(g.ints + first_row * G, g.ints + (first_row + (size_t)spec.dim) * G);
                        for (int32_t i0 = 0; i0 < spec.dim; ++i0)
                            std::copy(rows.begin() + (size_t)l.dictionary_perm[ord][i0] * G,
                                      rows.begin() + ((size_t)l.dictionary_perm[ord][i0] + 1) * G, g.ints + (first_row + (size_t)i0) * G);
                    }
                    first_row += n_rows;
                }
            }
            LOOM_B200_CHECK(LB_NAME(set_groups)(l.cross_cat.handle(), k, g.n_groups, g.counts, g.ints, g.flts));
            lbio_groups_free(&g);
        }
    }

    static double log_sum_exp(const std::vector<double> &x) {
        double m = -INFINITY, s = 0;
        for (double v : x) m = std::max(m, v);
        if (!(m > -INFINITY)) return m;
        for (double v : x) s += std::exp(v - m);
        return m + std::log(s);
    }

    // per latent, the log predictive probability of rows [begin, begin + n) of the loaded block
    std::vector<std::vector<float>> latent_scores(uint64_t begin, uint64_t n) {
        std::vector<std::vector<float>> out(latents_.size(), std::vector<float>(n));
        for (size_t l = 0; l < latents_.size(); ++l)
            LOOM_B200_CHECK(LB_NAME(query_score)(latents_[l]->cross_cat.handle(), begin, begin + n, out[l].data()));
        return out;
    }

    // QueryServer::call(Score) (src/query_server.cc:189-231) for n rows of `rows` at once
    std::vector<float> call_score(const Rows &rows, uint64_t begin, uint64_t n) {
        load(rows);
        return mean_scores(latent_scores(begin, n));
    }
    std::vector<float> mean_scores(const std::vector<std::vector<float>> &per) {
        const size_t n = per[0].size(), L = per.size();
        std::vector<float> out(n);
        std::vector<double> s(L);
        for (size_t i = 0; i < n; ++i) {
            for (size_t l = 0; l < L; ++l) s[l] = per[l][i];
            out[i] = (float)(log_sum_exp(s) - std::log((double)L));
        }
        return out;
    }

    // QueryServer::call(Sample) (src/query_server.cc:100-176); returns values [sample_count][n_to] in latent order
    std::vector<uint32_t> sample(const Rows &rows, uint64_t row, const uint32_t *to_sample, int32_t n_to, uint32_t sample_count) {
        load(rows);
        const size_t L = latents_.size();
        const auto per = latent_scores(row, 1);
        std::vector<double> w(L);
        double m = -INFINITY, total = 0;
        for (size_t l = 0; l < L; ++l) m = std::max(m, (double)per[l][0]);
        for (size_t l = 0; l < L; ++l) { w[l] = std::exp((double)per[l][0] - m); total += w[l]; }   // scores_to_probs
        std::vector<uint32_t> latent_counts(L, 0);
        for (uint32_t s = 0; s < sample_count; ++s) {                                                  // sample_discrete
            const double t = uniform() * total;
            double c = 0;
            size_t pick = L - 1;
            for (size_t l = 0; l < L; ++l) { c += w[l]; if (c > t) { pick = l; break; } }
            ++latent_counts[pick];
        }
        std::vector<uint32_t> values((size_t)sample_count * (size_t)std::max(n_to, 0));
        uint64_t done = 0;
        for (size_t l = 0; l < L; ++l) {
            if (!latent_counts[l]) continue;
            LOOM_B200_CHECK(LB_NAME(query_sample)(latents_[l]->cross_cat.handle(), row, to_sample, n_to, latent_counts[l], rng_(), 0,
                                                  values.data() + done * (size_t)n_to, nullptr));
            done += latent_counts[l];
        }
        return values;
    }

    void call_sample(const Rows &rows, uint64_t row, const uint32_t *to_sample, int32_t n_to, uint32_t sample_count, Answer &a) {
        const std::vector<uint32_t> values = sample(rows, row, to_sample, n_to, sample_count);
        a.has_sample = true;
        for (uint32_t s = 0; s < sample_count; ++s) {
            a.sample_feature.insert(a.sample_feature.end(), to_sample, to_sample + n_to);
            a.sample_value.insert(a.sample_value.end(), values.begin() + (size_t)s * n_to, values.begin() + (size_t)(s + 1) * n_to);
            a.sample_ptr.push_back(a.sample_feature.size());
        }
    }

    // QueryServer::call(Entropy) (src/query_server.cc:306-412)
    void call_entropy(const Rows &request_rows, const lbio_query_request &req, Answer &a) {
        const int F = latents_[0]->view.n_features;
        const uint64_t cond = (uint64_t)req.entropy_row;
        const uint32_t sample_count = req.entropy_sample_count;
        // cells the conditional already holds (its pos, and the tare's cells it does not remove)
        std::vector<uint8_t> held(F, 0);
        if (request_rows.has_tare[cond])
            for (int f = 0; f < F; ++f) held[f] = tare_observed_[f];
        for (uint64_t i = request_rows.neg_ptr[cond]; i < request_rows.neg_ptr[cond + 1]; ++i) held[request_rows.neg_feature[i]] = 0;
        for (uint64_t i = request_rows.pos_ptr[cond]; i < request_rows.pos_ptr[cond + 1]; ++i) held[request_rows.pos_feature[i]] = 1;
        // to_sample = union of every set (:316-329); tasks = the distinct unions row_set | col_set (:356-380)
        std::vector<uint8_t> in_union(F, 0);
        const int n_sets = req.n_row_sets + req.n_col_sets;
        for (int s = 0; s < n_sets; ++s)
            for (int32_t i = req.set_ptr[s]; i < req.set_ptr[s + 1]; ++i) in_union[req.set_features[i]] = 1;
        std::vector<uint32_t> to_sample;
        std::vector<int32_t> slot(F, -1);
        for (int f = 0; f < F; ++f)
            if (in_union[f]) {
                if (held[f]) {
                    a.errors.push_back("invalid request.entropy: a feature set overlaps the conditional");
                    return;
                }
                slot[f] = (int32_t)to_sample.size();
                to_sample.push_back((uint32_t)f);
            }
        std::map<std::vector<uint32_t>, int> task_index;
        std::vector<std::vector<uint32_t>> tasks;
        std::vector<int> cell_task;
        for (int i = 0; i < req.n_row_sets; ++i)
            for (int j = 0; j < req.n_col_sets; ++j) {
                std::vector<uint8_t> on(F, 0);
                for (int32_t p = req.set_ptr[i]; p < req.set_ptr[i + 1]; ++p) on[req.set_features[p]] = 1;
                const int cs = req.n_row_sets + j;
                for (int32_t p = req.set_ptr[cs]; p < req.set_ptr[cs + 1]; ++p) on[req.set_features[p]] = 1;
                std::vector<uint32_t> task;
                for (int f = 0; f < F; ++f)
                    if (on[f]) task.push_back((uint32_t)f);
                auto it = task_index.find(task);
                if (it == task_index.end()) {
                    it = task_index.emplace(task, (int)tasks.size()).first;
                    tasks.push_back(task);
                }
                cell_task.push_back(it->second);
            }
        const std::vector<uint32_t> values = sample(request_rows, cond, to_sample.data(), (int32_t)to_sample.size(), sample_count);
        // one block: the conditional, then for every sample the conditional + the sample restricted to each task
        Rows block;
        block.append(request_rows, cond);
        const uint64_t c0 = request_rows.pos_ptr[cond], c1 = request_rows.pos_ptr[cond + 1];
        for (uint32_t s = 0; s < sample_count; ++s)
            for (const auto &task : tasks) {
                block.has_tare.push_back(request_rows.has_tare[cond]);
                uint64_t ci = c0;
                size_t ti = 0;
                while (ci < c1 || ti < task.size()) {            // merge by ascending feature id
                    if (ti == task.size() || (ci < c1 && request_rows.pos_feature[ci] < task[ti])) {
                        block.pos_feature.push_back(request_rows.pos_feature[ci]);
                        block.pos_value.push_back(request_rows.pos_value[ci]);
                        ++ci;
                    } else {
                        const uint32_t f = task[ti++];
                        uint32_t v = values[(size_t)s * to_sample.size() + (size_t)slot[f]];
                        if (v == LB_DPD_OTHER) continue;          // a value outside the dictionary cannot be scored: leave the cell unobserved
                        block.pos_feature.push_back(f);
                        block.pos_value.push_back(v);
                    }
                }
                block.pos_ptr.push_back(block.pos_feature.size());
                block.neg_feature.insert(block.neg_feature.end(), request_rows.neg_feature.begin() + request_rows.neg_ptr[cond],
                                         request_rows.neg_feature.begin() + request_rows.neg_ptr[cond + 1]);
                block.neg_ptr.push_back(block.neg_feature.size());
            }
        const std::vector<float> scores = call_score(block, 0, block.size());
        const double base_score = scores[0];
        const size_t T = tasks.size();
        std::vector<double> count(T, 0), mean(T, 0), ctv(T, 0);       // Accum (:268-304): NormalInverseChiSq::Group::add_value
        for (uint32_t s = 0; s < sample_count; ++s)
            for (size_t t = 0; t < T; ++t) {
                const double x = base_score - (double)scores[1 + (size_t)s * T + t];
                count[t] += 1;
                const double delta = x - mean[t];
                mean[t] += delta / count[t];
                ctv[t] += delta * (x - mean[t]);
            }
        a.has_entropy = true;
        for (int t : cell_task) {
            a.means.push_back((float)mean[t]);
            a.variances.push_back((float)(ctv[t] / (count[t] - 1) / sample_count));
        }
    }

    // QueryServer::call(ScoreDerivative) (src/query_server.cc:443-545)
    void call_score_derivative(const Rows &request_rows, const lbio_query_request &req, Answer &a) {
        uint64_t row_count = 0;     // "FIXME is there a better way to get the row count?" (:459-464): the messages of rows_in
        uint32_t max_message_size = 0;
        LOOM_B200_IO(lbio_stream_stats(rows_in_.c_str(), &row_count, &max_message_size));
        Rows block;
        block.append(request_rows, (uint64_t)req.update_row);
        std::vector<uint64_t> ids;
        if (req.n_score_data == 0) {
            lbio_row_block all;
            LOOM_B200_IO(lbio_rows_read(rows_in_.c_str(), latents_[0]->model, 0, 0, 0, &all));
            block.append(all, 0, all.n_rows);
            ids.assign(all.rowids, all.rowids + all.n_rows);
            lbio_rows_free(&all);
        } else {
            for (int i = 0; i < req.n_score_data; ++i) {
                block.append(request_rows, (uint64_t)(req.first_score_data_row + i));
                ids.push_back((uint64_t)i);
            }
        }
        const uint64_t n = ids.size();
        const std::vector<float> before = call_score(block, 1, n);
        const uint32_t add = LB_ACTION_ADD(0), remove = LB_ACTION_REMOVE(0);
        for (auto &l : latents_)                                  // cat_kernels[i]->add_row(rng, update_row, assignments[i])
            LOOM_B200_CHECK(LB_NAME(cat_sweep)(l->cross_cat.handle(), &add, 1, rng_(), 0, nullptr, nullptr, nullptr, 0));
        const std::vector<float> after = mean_scores(latent_scores(1, n));
        for (auto &l : latents_)
            LOOM_B200_CHECK(LB_NAME(cat_sweep)(l->cross_cat.handle(), &remove, 1, rng_(), 0, nullptr, nullptr, nullptr, 0));
        std::vector<std::pair<uint64_t, float>> diffs(n);
        for (uint64_t i = 0; i < n; ++i) diffs[i] = {ids[i], (float)(((double)after[i] - (double)before[i]) * (double)row_count)};
        std::stable_sort(diffs.begin(), diffs.end(), [](const std::pair<uint64_t, float> &x, const std::pair<uint64_t, float> &y) {
            return x.second > y.second;
        });
        if (diffs.size() > req.row_limit) diffs.resize(req.row_limit);
        a.has_score_derivative = true;
        for (const auto &d : diffs) {
            a.ids.push_back(d.first);
            a.score_diffs.push_back(d.second);
        }
    }

    void respond(lbio_query *q, const lbio_model *schema, const Answer &a) {
        lbio_query_response r;
        std::memset(&r, 0, sizeof(r));
        std::vector<const char *> errors;
        for (const auto &e : a.errors) errors.push_back(e.c_str());
        r.n_errors = (int32_t)errors.size();
        r.errors = errors.data();
        r.has_sample = a.has_sample;
        r.n_samples = (int32_t)a.sample_ptr.size() - 1;
        r.sample_ptr = a.sample_ptr.data();
        r.sample_feature = a.sample_feature.data();
        r.sample_value = a.sample_value.data();
        r.has_score = a.has_score;
        r.score = a.score;
        r.has_entropy = a.has_entropy;
        r.n_entropy = (int32_t)a.means.size();
        r.means = a.means.data();
        r.variances = a.variances.data();
        r.has_score_derivative = a.has_score_derivative;
        r.n_ids = (int32_t)a.ids.size();
        r.ids = a.ids.data();
        r.score_diffs = a.score_diffs.data();
        LOOM_B200_IO(lbio_query_respond(q, schema, &r));
    }

    std::string rows_in_;
    SeedStream rng_;
    std::vector<std::unique_ptr<Latent>> latents_;
    std::vector<uint8_t> tare_observed_;
    std::vector<uint32_t> tare_values_;
    int32_t tare_count_ = 0;
};

}  // namespace b200
}  // namespace loom
