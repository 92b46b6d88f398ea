// loom.hpp -- loom::Loom (src/loom.hpp:43-135, src/loom.cc) re-hosted on the C ABIs of
// include/loom_b200.h (device engine) and include/loom_b200_io.h (loom's files).
//
// Same constructor arguments, same infer_single_pass / infer_multi_pass / dump entry points, same
// files in and out, so `loom_infer` (loom_infer.cc) takes the argument list loom.runner passes
// (src/infer.cc:68-81).  What differs is where the work happens: the rows file is decoded once into
// a device-resident columnar block; the schedules (which depend only on counts) are run ahead to the
// next batch boundary and the whole batch goes to the device as one action list.
#pragma once
#include <algorithm>
#include <chrono>
#include <cstring>
#include <memory>
#include <numeric>
#include <string>
#include <sys/stat.h>
#include <unordered_map>

#include "../../include/loom_b200_io.h"
#include "loom_b200.hpp"

namespace loom {
namespace b200 {

#define LOOM_B200_ERROR(...)                                                              \
    do {                                                                                  \
        std::fprintf(stderr, "ERROR ");                                                   \
        std::fprintf(stderr, __VA_ARGS__);                                                \
        std::fprintf(stderr, "\n\t%s : %d\n\t%s\n", __FILE__, __LINE__, __func__);        \
        std::abort();                                                                     \
    } while (0)

#define LOOM_B200_IO(expr)                                                                \
    do {                                                                                  \
        if ((expr) != 0) LOOM_B200_ERROR("%s", lbio_last_error());                        \
    } while (0)

// rng_t stand-in for the host-side seeds handed to the device kernels (SplitMix64)
struct SeedStream {
    uint64_t state;
    explicit SeedStream(uint64_t seed) : state(seed) {}
    uint64_t operator()() {
        uint64_t z = (state += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
};

class Loom {
public:
    // Loom::Loom (src/loom.cc:43-86)
    Loom(const lbio_config &config, const char *model_in, const char *groups_in = nullptr,
         const char *assign_in = nullptr, const char *tares_in = nullptr, int device = 0)
        : config_(config), cross_cat_(device) {
        LOOM_B200_IO(lbio_model_read(model_in, &model_file_));
        LOOM_B200_IO(lbio_model_get(model_file_, &view_));
        if (config_.cat_empty_group_count < 1) LOOM_B200_ERROR("expected 0 < empty_group_count");
        // A model with DirichletProcessDiscrete features is installed by rows_load, once the rows have
        // completed its dictionaries (loom's init models start with empty ones, loom/generate.py:103-107).
        if (!view_.n_dpd) model_install();
        const int F = view_.n_features;
        tare_observed_.assign(F, 0);
        tare_values_.assign(F, 0);
        if (tares_in && view_.n_dpd) tares_in_ = tares_in;   // read by rows_load, once the rows have completed the dictionaries
        else if (tares_in) {
            int32_t n_tares = 0;
            LOOM_B200_IO(lbio_tares_read(tares_in, model_file_, tare_observed_.data(), tare_values_.data(), &n_tares));
        }
        if (groups_in) groups_in_ = groups_in;
        if (assign_in) {   // Assignments::load (src/assignments.cc:50-66); placed on the rows by rows_load
            lbio_assign_block b;
            LOOM_B200_IO(lbio_assign_read(assign_in, &b));
            if (b.n_rows && b.n_kinds != view_.n_kinds)
                LOOM_B200_ERROR("expected assignments.kind_count() == cross_cat.kinds.size(); actual %d vs %d", b.n_kinds,
                                view_.n_kinds);
            assign_rowids_.assign(b.rowids, b.rowids + b.n_rows);
            assign_groupids_.assign(b.groupids, b.groupids + b.n_rows * (uint64_t)std::max(b.n_kinds, 0));
            lbio_assign_free(&b);
        }
    }
    ~Loom() {
        if (log_) lbio_log_close(log_);
        lbio_model_free(model_file_);
    }
    Loom(const Loom &) = delete;

    void log_to(const char *log_out, bool append) { LOOM_B200_IO(lbio_log_open(log_out, append, &log_)); }

    // Loom::dump (src/loom.cc:91-113)
    void dump(const char *model_out, const char *groups_out, const char *assign_out = nullptr) {
        const int K = (int)cross_cat_.kind_count();
        const std::vector<uint32_t> f2k = cross_cat_.featureid_to_kindid();
        std::vector<std::vector<uint32_t>> kind_features(K);
        for (size_t f = 0; f < f2k.size(); ++f) kind_features[f2k[f]].push_back((uint32_t)f);
        lbio_model *out_model = current_model(K, f2k);
        if (model_out) LOOM_B200_IO(lbio_model_write(out_model, model_out));
        if (groups_out || assign_out) {
            if (groups_out) mkdir(groups_out, 0775);
            std::vector<std::vector<uint32_t>> packed_to_sorted(K);
            for (int k = 0; k < K; ++k) {
                lb_kind_info_t info;
                LOOM_B200_CHECK(LB_NAME(kind_info)(h(), k, &info));
                const size_t G = (size_t)info.n_groups;
                std::vector<uint32_t> counts(G), ints((size_t)std::max(info.n_int_rows, 1) * G);
                std::vector<double> flts((size_t)std::max(info.n_flt_rows, 1) * G);
                LOOM_B200_CHECK(LB_NAME(get_groups)(h(), k, counts.data(), ints.data(), flts.data()));
                // CrossCat::get_sorted_groupids (src/cross_cat.cc:212-236): non-empty groups, largest first
                std::vector<uint32_t> order(G);
                int32_t n_live = 0;
                LOOM_B200_IO(lbio_sorted_groupids(counts.data(), (int32_t)G, order.data(), &n_live));
                order.resize((size_t)n_live);
                packed_to_sorted[k].assign(G, LB_UNASSIGNED);
                for (size_t s = 0; s < order.size(); ++s) packed_to_sorted[k][order[s]] = (uint32_t)s;
                if (groups_out) {
                    const std::string path = std::string(groups_out) + "/mixture." + std::to_string(k) + ".pbs.gz";   // store.hpp:59-66
                    LOOM_B200_IO(lbio_groups_write(path.c_str(), out_model, kind_features[k].data(),
                                                   (int32_t)kind_features[k].size(), (int32_t)G, counts.data(), ints.data(),
                                                   flts.data(), order.data(), (int32_t)order.size()));
                }
            }
            if (assign_out) {   // Assignments::dump (src/assignments.cc:68-100): FIFO order, sorted group ids
                const uint64_t R = row_count(), n = assigned_count_;
                std::vector<uint64_t> rowids(n);
                std::vector<uint32_t> groupids(n * (uint64_t)K);
                for (uint64_t i = 0; i < n; ++i) rowids[i] = rowids_[(assigned_pos_ + i) % R];
                for (int k = 0; k < K; ++k) {
                    const std::vector<uint32_t> packed = cross_cat_.groupids(k);
                    for (uint64_t i = 0; i < n; ++i) {
                        const uint32_t p = packed[(assigned_pos_ + i) % R];
                        if (p == LB_UNASSIGNED || p >= packed_to_sorted[k].size() || packed_to_sorted[k][p] == LB_UNASSIGNED)
                            LOOM_B200_ERROR("bad id: %u", p);
                        groupids[i * K + k] = packed_to_sorted[k][p];
                    }
                }
                LOOM_B200_IO(lbio_assign_write(assign_out, n, K, rowids.data(), groupids.data()));
            }
        }
        lbio_model_free(out_model);
    }

    // Loom::infer_single_pass (src/loom.cc:115-137): one greedy ADD pass in stream order
    void infer_single_pass(const char *rows_in, const char *assign_out = nullptr) {
        rows_load(rows_in);
        const uint64_t R = row_count();
        SeedStream rng(config_.seed);
        CatKernel cat_kernel(cross_cat_, rng());
        sweep_range(cat_kernel, R);
        unassigned_pos_ = 0;
        assigned_count_ += R;
        if (assign_out) {   // group ids as assigned: in an ADD-only pass packed id == global id
            const int K = (int)cross_cat_.kind_count();
            std::vector<uint32_t> groupids(R * (uint64_t)K);
            for (int k = 0; k < K; ++k) {
                const std::vector<uint32_t> packed = cross_cat_.groupids(k);
                for (uint64_t r = 0; r < R; ++r) groupids[r * K + k] = packed[r];
            }
            LOOM_B200_IO(lbio_assign_write(assign_out, R, K, rowids_.data(), groupids.data()));
        }
    }

    // Loom::infer_multi_pass (src/loom.cc:167-215)
    void infer_multi_pass(const char *rows_in, const char *checkpoint_in = nullptr, const char *checkpoint_out = nullptr) {
        const bool timing = std::getenv("LOOM_B200_TIMING") != nullptr;
        const auto t_begin = std::chrono::steady_clock::now();
        rows_load(rows_in);
        if (timing)
            std::fprintf(stderr, "[LOOM_B200_TIMING]   rows decode + upload       %8.1f ms\n",
                         std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count());
        const uint64_t R = row_count();
        ScheduleConfig sc;
        sc.extra_passes = config_.extra_passes;
        sc.small_data_size = config_.small_data_size;
        sc.big_data_size = config_.big_data_size;
        sc.max_reject_iters = config_.max_reject_iters;
        AcceleratingSchedule accelerating(sc);
        AnnealingSchedule annealing(sc.extra_passes);
        annealing.set_extra_passes(accelerating.extra_passes(assigned_count_));
        BatchingSchedule batching;
        KernelDisablingSchedule disabling(sc.max_reject_iters);
        const auto start = std::chrono::steady_clock::now();
        auto checkpointing_test = [&] {   // CheckpointingSchedule (src/schedules.hpp:252-274)
            return std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count() >=
                   (double)config_.checkpoint_period_sec;
        };

        lbio_checkpoint checkpoint;
        std::memset(&checkpoint, 0, sizeof(checkpoint));
        uint64_t seed = config_.seed;
        if (checkpoint_in) {
            LOOM_B200_IO(lbio_checkpoint_read(checkpoint_in, &checkpoint));
            seed = checkpoint.seed;
            const uint64_t a = checkpoint.assigned_pos % R, u = checkpoint.unassigned_pos % R;
            if (assigned_count_ && (a != assigned_pos_ || u != unassigned_pos_))
                LOOM_B200_ERROR("checkpoint row positions (%llu, %llu) disagree with the assignments (%llu, %llu)",
                                (unsigned long long)a, (unsigned long long)u, (unsigned long long)assigned_pos_,
                                (unsigned long long)unassigned_pos_);
            assigned_pos_ = a;
            unassigned_pos_ = u;
            annealing.load(checkpoint.annealing_state);
            batching.load(checkpoint.schedule_row_count);
            disabling.load(checkpoint.reject_iters);
            checkpoint.tardis_iter += 1;
        } else {
            checkpoint.row_count = R;   // InFile::stream_stats(rows_in).message_count
            checkpoint.tardis_iter = 0;
            log_metrics(checkpoint.tardis_iter, nullptr);
        }
        if (!(assigned_count_ < checkpoint.row_count))
            LOOM_B200_ERROR("expected assignments.row_count() < checkpoint.row_count(); actual %llu vs %llu",
                            (unsigned long long)assigned_count_, (unsigned long long)checkpoint.row_count);
        if (checkpoint.row_count != R) LOOM_B200_ERROR("checkpoint.row_count differs from the rows file");
        SeedStream rng(seed);

        checkpoint.finished = 0;
        bool finished = false;
        if (config_.kind_iterations && disabling.test()) {
            finished = infer_structure(true, checkpoint, annealing, accelerating, batching, disabling, checkpointing_test, rng);
            if (!finished && !checkpointing_test())
                finished = infer_structure(false, checkpoint, annealing, accelerating, batching, disabling, checkpointing_test, rng);
        } else {
            finished = infer_structure(false, checkpoint, annealing, accelerating, batching, disabling, checkpointing_test, rng);
        }
        checkpoint.finished = finished;

        if (checkpoint_out) {
            checkpoint.seed = rng();
            checkpoint.unassigned_pos = unassigned_pos_;
            checkpoint.assigned_pos = assigned_pos_;
            checkpoint.annealing_state = annealing.state();
            checkpoint.schedule_row_count = batching.row_count();
            checkpoint.reject_iters = disabling.reject_iters();
            LOOM_B200_IO(lbio_checkpoint_write(checkpoint_out, &checkpoint));
        }
    }

    // Loom::posterior_enum (src/loom.cc:452-511): greedy ADD of every row, then sample_count times
    // { sample_skip times { REMOVE + ADD every row; [kind kernel]; hyper kernel }; dump the latent
    // state and its score }.  The protocol of loom/test/test_posterior_enum.py.
    void posterior_enum(const char *rows_in, const char *samples_out) {
        const uint32_t sample_count = config_.posterior_enum_sample_count, sample_skip = config_.posterior_enum_sample_skip;
        if (sample_count < 1) LOOM_B200_ERROR("expected 1 <= sample_count");
        if (!(sample_skip > 0 || sample_count == 1)) LOOM_B200_ERROR("zero diversity");
        rows_load(rows_in);
        const uint64_t R = row_count();
        SeedStream rng(config_.seed);
        CatKernel cat_kernel(cross_cat_, rng());
        std::unique_ptr<HyperKernel> hyper_kernel;
        if (config_.hyper_run && view_.has_hyper_prior) hyper_kernel.reset(new HyperKernel(cross_cat_, view_.hyper_prior));
        if (!assigned_count_) {
            sweep_range(cat_kernel, R);
            assigned_count_ = R;
            assigned_pos_ = unassigned_pos_ = 0;
        }
        if (assigned_count_ != R) LOOM_B200_ERROR("posterior_enum needs every row assigned or none");
        lbio_samples *samples = nullptr;
        LOOM_B200_IO(lbio_samples_open(samples_out, &samples));
        std::unique_ptr<KindKernel> kind_kernel;
        if (config_.kind_iterations > 0)
            kind_kernel.reset(new KindKernel(cross_cat_, (int)config_.kind_empty_kind_count, (int)config_.kind_iterations, rng()));
        for (uint32_t i = 0; i < sample_count; ++i) {
            for (uint32_t t = 0; t < sample_skip; ++t) {
                for (uint64_t r = 0; r < R; ++r) {
                    cat_kernel.remove_row(r);
                    cat_kernel.add_row(r);
                }
                if (kind_kernel || hyper_kernel) cat_kernel.wait();   // else the whole skip is one action list
                if (kind_kernel) kind_kernel->try_run();
                if (hyper_kernel) hyper_kernel->run(rng());
            }
            cat_kernel.wait();
            dump_posterior_enum(samples);
        }
        LOOM_B200_IO(lbio_samples_close(samples));
    }

    // Loom::generate (src/loom.cc:548-562) + generate_rows (src/generate.hpp:36-93): hypers from the prior grids,
    // then config.generate.row_count rows synthesised on the device and written as a rows stream (ids 0, 1, ...)
    void generate(const char *rows_out) {
        if (assigned_count_ || !assign_rowids_.empty()) LOOM_B200_ERROR("expected assignments.row_count() == 0");
        if (view_.n_dpd) model_install();      // generate needs the dictionaries the model file carries
        SeedStream rng(config_.seed);
        if (config_.hyper_run && view_.has_hyper_prior) HyperKernel(cross_cat_, view_.hyper_prior).run(rng());
        const uint64_t R = config_.generate_row_count;
        LOOM_B200_CHECK(LB_NAME(generate_rows)(h(), R, config_.generate_density, rng()));
        cross_cat_.rows_attached(R);
        rowids_.resize(R);
        std::iota(rowids_.begin(), rowids_.end(), (uint64_t)0);
        assigned_pos_ = unassigned_pos_ = 0;
        assigned_count_ = R;
        // the block as a rows stream: the diff encoder with an empty tare gives exactly the observed cells per row
        const std::vector<uint8_t> no_tare_obs(view_.n_features, 0);
        const std::vector<uint32_t> no_tare_val(view_.n_features, 0);
        uint64_t n_pos = 0, n_neg = 0;
        LOOM_B200_CHECK(LB_NAME(differ_compress_rows)(h(), no_tare_obs.data(), no_tare_val.data(), &n_pos, &n_neg));
        std::vector<uint64_t> pos_ptr(R + 1), neg_ptr(R + 1);
        std::vector<uint32_t> pos_feature(n_pos + 1), pos_value(n_pos + 1), neg_feature(n_neg + 1);
        LOOM_B200_CHECK(LB_NAME(differ_fetch)(h(), pos_ptr.data(), pos_feature.data(), pos_value.data(), neg_ptr.data(), neg_feature.data()));
        lbio_row_block out;
        std::memset(&out, 0, sizeof(out));
        out.n_rows = R;
        out.rowids = rowids_.data();
        out.pos_ptr = pos_ptr.data();
        out.pos_feature = pos_feature.data();
        out.pos_value = pos_value.data();
        out.neg_ptr = neg_ptr.data();
        out.neg_feature = neg_feature.data();
        LOOM_B200_IO(lbio_rows_write(rows_out, model_file_, &out));
    }

    // Loom::mix (src/loom.cc:564-586): sample_skip times { REMOVE + ADD every row; kind kernel; hyper kernel } on a
    // fully assigned state -- what loom.generate runs to decorrelate a synthetic dataset from the model that made it
    void mix(const char *rows_in) {
        const uint32_t sample_skip = config_.generate_sample_skip;
        if (!(sample_skip > 0)) LOOM_B200_ERROR("zero diversity");
        rows_load(rows_in);
        const uint64_t R = row_count();
        if (assigned_count_ != R) LOOM_B200_ERROR("mix needs every row assigned");
        SeedStream rng(config_.seed);
        CatKernel cat_kernel(cross_cat_, rng());
        std::unique_ptr<HyperKernel> hyper_kernel;
        if (config_.hyper_run && view_.has_hyper_prior) hyper_kernel.reset(new HyperKernel(cross_cat_, view_.hyper_prior));
        std::unique_ptr<KindKernel> kind_kernel;
        if (config_.kind_iterations > 0)
            kind_kernel.reset(new KindKernel(cross_cat_, (int)config_.kind_empty_kind_count, (int)config_.kind_iterations, rng()));
        for (uint32_t i = 0; i < sample_skip; ++i) {
            for (uint64_t r = 0; r < R; ++r) {
                cat_kernel.remove_row((assigned_pos_ + r) % R);
                cat_kernel.add_row((assigned_pos_ + r) % R);
            }
            cat_kernel.wait();
            if (kind_kernel) kind_kernel->try_run();
            if (hyper_kernel) hyper_kernel->run(rng());
        }
    }

    CrossCat &cross_cat() { return cross_cat_; }
    uint64_t row_count() const { return rowids_.size(); }
    uint64_t assigned_count() const { return assigned_count_; }

private:
    lb_handle h() const { return cross_cat_.handle(); }

    // The unzip / parse / split stages (src/cat_pipeline.cc:63-87) as one decode + one upload; then
    // CrossCat::mixture_load (src/cross_cat.cc:148-196) and the StreamInterval cursors
    // (src/stream_interval.hpp:66-80, 98-135).
    void model_install() {
        cross_cat_.model_load(view_.n_features, view_.specs, view_.aux, view_.n_aux, view_.n_kinds,
                              view_.featureid_to_kindid, view_.kind_py, view_.topology_py,
                              (int)config_.cat_empty_group_count);
    }

    void rows_load(const char *rows_in) {
        lbio_row_block rows;
        if (view_.n_dpd) {   // DPD Shared.add_value for every value of the stream, then the model goes to the device
            int32_t n_new = 0;
            LOOM_B200_IO(lbio_rows_read_open(rows_in, model_file_, config_.seed, 0, &rows, &n_new));
            if (!tares_in_.empty()) {   // the tare row's own DPD values are in no row's pos: they enter the dictionary here
                int32_t n_tares = 0;
                LOOM_B200_IO(lbio_tares_read_open(tares_in_.c_str(), model_file_, config_.seed, tare_observed_.data(),
                                                  tare_values_.data(), &n_tares, &n_new));
            }
            LOOM_B200_IO(lbio_model_get(model_file_, &view_));
            model_install();
        } else {
            LOOM_B200_IO(lbio_rows_read(rows_in, model_file_, 0, 0, 0, &rows));
        }
        if (!rows.n_rows) LOOM_B200_ERROR("stream is empty");
        rowids_.assign(rows.rowids, rows.rowids + rows.n_rows);
        LOOM_B200_CHECK(LB_NAME(set_rows_diff)(h(), rows.n_rows, tare_observed_.data(), tare_values_.data(),
                                               rows.row_has_tare, rows.pos_ptr, rows.pos_feature, rows.pos_value,
                                               rows.neg_ptr, rows.neg_feature));
        cross_cat_.rows_attached(rows.n_rows);
        lbio_rows_free(&rows);
        const uint64_t R = row_count();
        const int K = view_.n_kinds;
        if (!groups_in_.empty()) {
            for (int k = 0; k < K; ++k) {
                std::vector<uint32_t> fids;
                for (int f = 0; f < view_.n_features; ++f)
                    if ((int)view_.featureid_to_kindid[f] == k) fids.push_back((uint32_t)f);
                const std::string path = groups_in_ + "/mixture." + std::to_string(k) + ".pbs.gz";
                lbio_group_block g;
                LOOM_B200_IO(lbio_groups_read(path.c_str(), model_file_, fids.data(), (int32_t)fids.size(), &g));
                LOOM_B200_CHECK(LB_NAME(set_groups)(h(), k, g.n_groups, g.counts, g.ints, g.flts));
                lbio_groups_free(&g);
            }
        }
        assigned_pos_ = unassigned_pos_ = 0;
        assigned_count_ = 0;
        if (!assign_rowids_.empty()) {
            std::unordered_map<uint64_t, uint64_t> position;
            position.reserve(R * 2);
            for (uint64_t r = 0; r < R; ++r) position[rowids_[r]] = r;
            const uint64_t n = assign_rowids_.size();
            auto first = position.find(assign_rowids_.front());
            if (first == position.end()) LOOM_B200_ERROR("row.id not found: %llu", (unsigned long long)assign_rowids_.front());
            assigned_pos_ = first->second;
            std::vector<std::vector<uint32_t>> packed(K, std::vector<uint32_t>(R, LB_UNASSIGNED));
            for (uint64_t i = 0; i < n; ++i) {
                auto it = position.find(assign_rowids_[i]);
                if (it == position.end()) LOOM_B200_ERROR("row.id not found: %llu", (unsigned long long)assign_rowids_[i]);
                if (it->second != (assigned_pos_ + i) % R)
                    LOOM_B200_ERROR("assigned rows are not a contiguous window of the rows stream at assignment %llu",
                                    (unsigned long long)i);
                for (int k = 0; k < K; ++k) packed[k][it->second] = assign_groupids_[i * K + k];
            }
            for (int k = 0; k < K; ++k) {
                lb_kind_info_t info;
                LOOM_B200_CHECK(LB_NAME(kind_info)(h(), k, &info));
                if (n > info.sample_size)   // src/loom.cc:73-77
                    LOOM_B200_ERROR("expected assignments.row_count() <= sample_size; actual %llu vs %llu",
                                    (unsigned long long)n, (unsigned long long)info.sample_size);
                LOOM_B200_CHECK(LB_NAME(set_assignments)(h(), k, packed[k].data()));
            }
            assigned_count_ = n;
            unassigned_pos_ = (assigned_pos_ + n) % R;
        }
    }

    // protobuf::CrossCat of the current state: hypers and kind structure from the device, grids and
    // dictionaries from the input model
    lbio_model *current_model(int K, const std::vector<uint32_t> &f2k) {
        const int F = view_.n_features;
        std::vector<lb_feature_spec> specs(F);
        std::vector<float> aux((size_t)std::max(view_.n_aux, 1)), kind_py((size_t)2 * K);
        float topology[2];
        LOOM_B200_CHECK(LB_NAME(get_hypers)(h(), specs.data(), aux.data(), kind_py.data(), topology));
        lbio_model *m = nullptr;
        LOOM_B200_IO(lbio_model_create(F, specs.data(), aux.data(), view_.n_aux, K, f2k.data(), kind_py.data(), topology,
                                       view_.has_hyper_prior ? &view_.hyper_prior : nullptr, view_.dpd_offsets,
                                       view_.dpd_values, &m));
        return m;
    }

    void sweep_range(CatKernel &cat_kernel, uint64_t n_adds) {
        const uint64_t R = row_count(), CHUNK = 1ull << 24;
        for (uint64_t i = 0; i < n_adds; ++i) {
            cat_kernel.add_row((unassigned_pos_ + i) % R);
            if ((i + 1) % CHUNK == 0) cat_kernel.wait();
        }
        cat_kernel.wait();
    }

    // infer_kind_structure_* / infer_cat_structure_* (src/loom.cc:217-450), one loop for both: the
    // schedules run ahead to the batch boundary, the batch is swept on the device, then the batch
    // kernels run in the reference's order.  Returns true when every row is assigned.
    template <class CheckpointingTest>
    bool infer_structure(bool with_kinds, lbio_checkpoint &checkpoint, AnnealingSchedule &annealing,
                         AcceleratingSchedule &accelerating, BatchingSchedule &batching,
                         KernelDisablingSchedule &disabling, CheckpointingTest &checkpointing_test, SeedStream &rng) {
        const uint64_t R = row_count();
        std::unique_ptr<KindKernel> kind_kernel;
        if (with_kinds)
            kind_kernel.reset(new KindKernel(cross_cat_, (int)config_.kind_empty_kind_count, (int)config_.kind_iterations, rng()));
        CatKernel cat_kernel(cross_cat_, rng());
        std::unique_ptr<HyperKernel> hyper_kernel;
        if (config_.hyper_run && view_.has_hyper_prior) hyper_kernel.reset(new HyperKernel(cross_cat_, view_.hyper_prior));
        KernelCounters counters;
        while (assigned_count_ != checkpoint.row_count) {
            if (annealing.next_action_is_add()) {
                cat_kernel.add_row(unassigned_pos_);
                unassigned_pos_ = (unassigned_pos_ + 1) % R;
                ++assigned_count_;
                batching.add();
            } else {
                cat_kernel.remove_row(assigned_pos_);
                assigned_pos_ = (assigned_pos_ + 1) % R;
                --assigned_count_;
                batching.remove();
            }
            if (batching.test()) {
                counters.cat_usec += timed([&] { cat_kernel.wait(); });
                annealing.set_extra_passes(accelerating.extra_passes(assigned_count_));
                if (kind_kernel) {
                    bool changed = false;
                    counters.kind_usec += timed([&] { changed = kind_kernel->try_run(); });
                    ++counters.kind_total;
                    counters.kind_change += changed;
                    disabling.run(changed);
                }
                if (hyper_kernel) counters.hyper_usec += timed([&] { hyper_kernel->run(rng()); });
                checkpoint.tardis_iter += 1;
                log_metrics(checkpoint.tardis_iter, &counters);
                if (checkpointing_test()) return false;
                if (kind_kernel && !disabling.test()) return false;
            }
        }
        counters.cat_usec += timed([&] { cat_kernel.wait(); });
        checkpoint.tardis_iter += 1;
        kind_kernel.reset();   // ~KindKernel drops the ephemeral kinds (src/kind_kernel.cc:75-81)
        log_metrics(checkpoint.tardis_iter, &counters);
        return true;
    }

    // Loom::dump_posterior_enum (src/loom.cc:513-546): kinds with features, their row partitions, the score
    void dump_posterior_enum(lbio_samples *samples) {
        const int K = (int)cross_cat_.kind_count();
        const uint64_t R = row_count();
        const std::vector<uint32_t> f2k = cross_cat_.featureid_to_kindid();
        std::vector<int32_t> kind_feature_ptr(1, 0), kind_group_ptr(1, 0);
        std::vector<uint32_t> featureids, rowids;
        std::vector<uint64_t> group_row_ptr(1, 0);
        for (int k = 0; k < K; ++k) {
            const size_t before = featureids.size();
            for (size_t f = 0; f < f2k.size(); ++f)
                if ((int)f2k[f] == k) featureids.push_back((uint32_t)f);
            if (featureids.size() == before) continue;        // featureless kinds are not part of the latent state
            kind_feature_ptr.push_back((int32_t)featureids.size());
            const std::vector<uint32_t> packed = cross_cat_.groupids(k);
            std::unordered_map<uint32_t, std::vector<uint32_t>> groups;
            for (uint64_t r = 0; r < R; ++r)
                if (packed[r] != LB_UNASSIGNED) groups[packed[r]].push_back((uint32_t)rowids_[r]);
            for (const auto &g : groups) {
                rowids.insert(rowids.end(), g.second.begin(), g.second.end());
                group_row_ptr.push_back(rowids.size());
            }
            kind_group_ptr.push_back((int32_t)group_row_ptr.size() - 1);
        }
        LOOM_B200_IO(lbio_samples_write(samples, (int32_t)kind_feature_ptr.size() - 1, kind_feature_ptr.data(),
                                        featureids.data(), kind_group_ptr.data(), group_row_ptr.data(), rowids.data(),
                                        cross_cat_.score_data()));
    }

    struct KernelCounters { uint64_t cat_usec = 0, hyper_usec = 0, kind_usec = 0, kind_total = 0, kind_change = 0; };

    template <class Fun> static uint64_t timed(Fun &&fun) {
        const auto t0 = std::chrono::steady_clock::now();
        fun();
        return (uint64_t)std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count();
    }

    // Loom::log_metrics (src/loom.cc:139-165)
    void log_metrics(uint64_t iter, const KernelCounters *counters) {
        if (!log_) return;
        const int K = (int)cross_cat_.kind_count();
        std::vector<lb_feature_spec> specs(view_.n_features);
        std::vector<float> aux((size_t)std::max(view_.n_aux, 1)), all_py((size_t)2 * K), kind_py;
        lbio_log_entry e;
        std::memset(&e, 0, sizeof(e));
        LOOM_B200_CHECK(LB_NAME(get_hypers)(h(), specs.data(), aux.data(), all_py.data(), e.topology_py));
        std::vector<uint32_t> feature_counts, category_counts;
        for (int k = 0; k < K; ++k) {
            lb_kind_info_t info;
            LOOM_B200_CHECK(LB_NAME(kind_info)(h(), k, &info));
            if (!info.n_features) continue;
            feature_counts.push_back((uint32_t)info.n_features);
            category_counts.push_back((uint32_t)info.n_groups - config_.cat_empty_group_count);
            kind_py.push_back(all_py[2 * k]);
            kind_py.push_back(all_py[2 * k + 1]);
        }
        e.iter = (uint32_t)iter;
        e.n_kinds = (int32_t)feature_counts.size();
        e.kind_py = kind_py.data();
        e.feature_counts = feature_counts.data();
        e.category_counts = category_counts.data();
        e.score = cross_cat_.score_data();
        e.assigned_object_count = assigned_count_;
        e.kl_divergence = assigned_count_ ? (float)((-e.score - std::log((double)assigned_count_)) / (double)assigned_count_) : 0.f;
        if (counters) {
            e.cat_total_time_usec = counters->cat_usec;
            e.hyper_total_time_usec = counters->hyper_usec;
            e.kind_total_time_usec = counters->kind_usec;
            e.kind_total_count = counters->kind_total;
            e.kind_change_count = counters->kind_change;
        }
        LOOM_B200_IO(lbio_log_write(log_, &e));
    }

    lbio_config config_;
    CrossCat cross_cat_;
    lbio_model *model_file_ = nullptr;
    lbio_model_view view_;
    std::vector<uint8_t> tare_observed_;
    std::vector<uint32_t> tare_values_;
    std::string groups_in_, tares_in_;
    std::vector<uint64_t> assign_rowids_;
    std::vector<uint32_t> assign_groupids_;
    std::vector<uint64_t> rowids_;              // Row.id by stream position
    uint64_t assigned_pos_ = 0, unassigned_pos_ = 0, assigned_count_ = 0;   // StreamInterval + Assignments::row_count
    lbio_log *log_ = nullptr;
};

}  // namespace b200
}  // namespace loom
