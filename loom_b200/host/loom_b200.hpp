// loom_b200.hpp -- C++ shell that re-hosts loom's engine classes on the C ABI of
// include/loom_b200.h.  Header only; link with -lloomb200.
//
// Class and method names follow the reference so a loom maintainer can swap
// the bodies of src/cat_kernel.hpp, src/kind_kernel.{hpp,cc} and
// src/hyper_kernel.{hpp,cc} for these calls:
//   loom::CrossCat        src/cross_cat.hpp:43-75      -> b200::CrossCat (device-resident state)
//   loom::CatKernel       src/cat_kernel.hpp:41-103    -> b200::CatKernel
//   loom::KindKernel      src/kind_kernel.hpp:41-124   -> b200::KindKernel
//   loom::HyperKernel     src/hyper_kernel.hpp:43-111  -> b200::HyperKernel
// Errors follow loom's convention: message to stderr + abort()
// (LOOM_ERROR, src/common.hpp:52-59); there are no return codes across this API.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

#include "../../include/loom_b200.h"

namespace loom {
namespace b200 {

#define LOOM_B200_CHECK(expr)                                                              \
    do {                                                                                   \
        if ((expr) != 0) {                                                                 \
            std::fprintf(stderr, "ERROR %s\n\t%s : %d\n\t%s\n", LB_NAME(last_error)(), __FILE__, \
                         __LINE__, __func__);                                              \
            std::abort();                                                                  \
        }                                                                                  \
    } while (0)

// One batch of rows, already split into typed columns (the product of loom's unzip /
// parse / ValueSplitter::split stages, src/cat_pipeline.cc:63-87).
struct RowBlock {
    uint64_t row_count = 0;
    std::vector<uint8_t> cat8;    // [F8][R]
    std::vector<uint32_t> wide;   // [FW][R]
    std::vector<uint32_t> mask;   // [F][ceil(R/32)]
};

class CrossCat {
public:
    explicit CrossCat(int device = 0) { LOOM_B200_CHECK(LB_NAME(create)(device, &handle_)); }
    ~CrossCat() { LB_NAME(destroy)(handle_); }
    CrossCat(const CrossCat &) = delete;
    CrossCat &operator=(const CrossCat &) = delete;

    // CrossCat::model_load + mixture_init_unobserved (src/cross_cat.cc:50-120)
    void model_load(const std::vector<lb_feature_spec> &features, const std::vector<float> &aux,
                    const std::vector<uint32_t> &featureid_to_kindid, const std::vector<float> &kind_py,
                    const float topology_py[2], int empty_group_count = 1) {
        feature_count_ = (int)features.size();
        LOOM_B200_CHECK(LB_NAME(set_model)(handle_, (int32_t)features.size(), features.data(), aux.data(),
                                     (int32_t)aux.size(), (int32_t)(kind_py.size() / 2),
                                     featureid_to_kindid.data(), kind_py.data(), topology_py,
                                     empty_group_count));
    }
    // the same from the arrays of a model file (lbio_model_view, include/loom_b200_io.h)
    void model_load(int32_t n_features, const lb_feature_spec *features, const float *aux, int32_t n_aux,
                    int32_t n_kinds, const uint32_t *featureid_to_kindid, const float *kind_py,
                    const float *topology_py, int empty_group_count) {
        feature_count_ = n_features;
        LOOM_B200_CHECK(LB_NAME(set_model)(handle_, n_features, features, aux, n_aux, n_kinds, featureid_to_kindid,
                                           kind_py, topology_py, empty_group_count));
    }
    // rows uploaded through the C ABI directly (lb_set_rows_diff)
    void rows_attached(uint64_t row_count) { row_count_ = row_count; }
    void rows_load(const RowBlock &rows) {
        row_count_ = rows.row_count;
        LOOM_B200_CHECK(LB_NAME(set_rows)(handle_, rows.row_count, rows.cat8.data(), rows.wide.data(), rows.mask.data()));
    }
    // another sample over the same rows file (loom/tasks.py:173-194): read `owner`'s block
    void rows_share(const CrossCat &owner) {
        row_count_ = owner.row_count_;
        LOOM_B200_CHECK(LB_NAME(share_rows)(handle_, owner.handle_));
    }
    // CrossCat::score_data (src/cross_cat.cc:238-250)
    float score_data() const {
        double s = 0;
        LOOM_B200_CHECK(LB_NAME(score_data)(handle_, &s));
        return (float)s;
    }
    size_t kind_count() const {
        int32_t k = 0;
        LOOM_B200_CHECK(LB_NAME(get_kinds)(handle_, &k, nullptr));
        return (size_t)k;
    }
    std::vector<uint32_t> featureid_to_kindid() const {
        std::vector<uint32_t> f2k(feature_count_);
        int32_t k = 0;
        LOOM_B200_CHECK(LB_NAME(get_kinds)(handle_, &k, f2k.data()));
        return f2k;
    }
    // Assignments::groupids(kind) as packed ids (src/assignments.hpp:108-111)
    std::vector<uint32_t> groupids(size_t kindid) const {
        std::vector<uint32_t> out(row_count_);
        LOOM_B200_CHECK(LB_NAME(get_assignments)(handle_, (int32_t)kindid, out.data()));
        return out;
    }
    // ProductMixture::score_value for a batch of rows (src/product_mixture.cc:421-434)
    std::vector<float> score_rows(size_t kindid, uint64_t row_begin, uint64_t row_end) const {
        lb_kind_info_t info;
        LOOM_B200_CHECK(LB_NAME(kind_info)(handle_, (int32_t)kindid, &info));
        std::vector<float> out((size_t)(row_end - row_begin) * info.n_groups);
        LOOM_B200_CHECK(LB_NAME(score_rows)(handle_, (int32_t)kindid, row_begin, row_end, out.data()));
        return out;
    }
    // QueryServer::call(Score) for this sample (src/query_server.cc:189-231): log predictive of rows [row_begin, row_end)
    std::vector<float> query_score(uint64_t row_begin, uint64_t row_end) const {
        std::vector<float> out((size_t)(row_end - row_begin));
        LOOM_B200_CHECK(LB_NAME(query_score)(handle_, row_begin, row_end, out.data()));
        return out;
    }
    // ProductMixture::sample_value per kind (src/product_mixture.cc:637-649): n_samples joint draws of `to_sample`
    // (ascending feature ids) given the loaded row `row`; values [n_samples][to_sample.size()] as raw cells
    std::vector<uint32_t> query_sample(uint64_t row, const std::vector<uint32_t> &to_sample, uint32_t n_samples, uint64_t seed,
                                       uint64_t draw_offset = 0) const {
        std::vector<uint32_t> out((size_t)n_samples * to_sample.size());
        LOOM_B200_CHECK(LB_NAME(query_sample)(handle_, row, to_sample.data(), (int32_t)to_sample.size(), n_samples, seed,
                                              draw_offset, out.data(), nullptr));
        return out;
    }
    // generate_rows (src/generate.hpp:36-93): the engine then holds the synthesised rows, every row assigned
    void generate_rows(uint64_t row_count, float density, uint64_t seed) {
        LOOM_B200_CHECK(LB_NAME(generate_rows)(handle_, row_count, density, seed));
        row_count_ = row_count;
    }
    lb_handle handle() const { return handle_; }
    uint64_t row_count() const { return row_count_; }

private:
    lb_handle handle_ = nullptr;
    int feature_count_ = 0;
    uint64_t row_count_ = 0;
};

// CatKernel (src/cat_kernel.hpp:41-103).  Rows are addressed by their position in the loaded
// block; add_row / remove_row queue actions and flush() runs them in order on the device,
// which is what CatPipeline's ring buffer does with threads (src/cat_pipeline.cc:103-125).
class CatKernel {
public:
    CatKernel(CrossCat &cross_cat, uint64_t seed) : cross_cat_(cross_cat), seed_(seed) {}
    void add_row(uint64_t row_position) { actions_.push_back(LB_ACTION_ADD(row_position)); }
    void remove_row(uint64_t row_position) { actions_.push_back(LB_ACTION_REMOVE(row_position)); }
    // pipeline.wait() (src/loom.cc:425-429)
    void wait() {
        if (actions_.empty()) return;
        LOOM_B200_CHECK(LB_NAME(cat_sweep)(cross_cat_.handle(), actions_.data(), actions_.size(), seed_, offset_,
                                     nullptr, nullptr, nullptr, 0));
        offset_ += actions_.size();
        actions_.clear();
    }

private:
    CrossCat &cross_cat_;
    uint64_t seed_, offset_ = 0;
    std::vector<uint32_t> actions_;
};

// loom.tasks.infer_many (loom/tasks.py:173-194) runs one process per sample; on the device the
// samples are swept together: the same queued actions on every chain, ONE launch, chain c with
// seeds[c].  This is the throughput mode (hundreds of chains fill the GPU); a single CatKernel is
// the latency mode.
class MultiCatKernel {
public:
    MultiCatKernel(const std::vector<CrossCat *> &chains, const std::vector<uint64_t> &seeds) : seeds_(seeds) {
        for (CrossCat *c : chains) handles_.push_back(c->handle());
    }
    void add_row(uint64_t row_position) { actions_.push_back(LB_ACTION_ADD(row_position)); }
    void remove_row(uint64_t row_position) { actions_.push_back(LB_ACTION_REMOVE(row_position)); }
    void wait() {
        if (actions_.empty()) return;
        LOOM_B200_CHECK(LB_NAME(cat_sweep_multi)(handles_.data(), (int32_t)handles_.size(), actions_.data(), actions_.size(),
                                           seeds_.data(), offset_));
        offset_ += actions_.size();
        actions_.clear();
    }

private:
    std::vector<lb_handle> handles_;
    std::vector<uint64_t> seeds_;
    uint64_t offset_ = 0;
    std::vector<uint32_t> actions_;
};

// KindKernel (src/kind_kernel.hpp:41-124): the cat sweep runs over real + ephemeral kinds;
// try_run scores every feature against every kind and re-assigns features.
class KindKernel {
public:
    KindKernel(CrossCat &cross_cat, int empty_kind_count, int iterations, uint64_t seed)
        : cross_cat_(cross_cat), empty_kind_count_(empty_kind_count), iterations_(iterations), seed_(seed) {
        LOOM_B200_CHECK(LB_NAME(init_featureless_kinds)(cross_cat_.handle(), empty_kind_count_, next_seed()));
    }
    ~KindKernel() { LB_NAME(init_featureless_kinds)(cross_cat_.handle(), 0, 0); }  // kind_kernel.cc:75-81
    // KindKernel::try_run (src/kind_kernel.cc:118-157); returns true when a feature moved
    bool try_run() {
        LOOM_B200_CHECK(LB_NAME(kind_proposer_score)(cross_cat_.handle(), nullptr));
        int32_t changed = 0;
        LOOM_B200_CHECK(LB_NAME(kind_run)(cross_cat_.handle(), iterations_, next_seed(), nullptr, &changed));
        LOOM_B200_CHECK(LB_NAME(init_featureless_kinds)(cross_cat_.handle(), empty_kind_count_, next_seed()));
        return changed > 0;
    }

private:
    uint64_t next_seed() { return seed_ * 0x9E3779B97F4A7C15ull + (++calls_); }
    CrossCat &cross_cat_;
    int empty_kind_count_, iterations_;
    uint64_t seed_, calls_ = 0;
};

// HyperKernel (src/hyper_kernel.hpp:43-111)
class HyperKernel {
public:
    HyperKernel(CrossCat &cross_cat, const lb_hyper_prior &prior) : cross_cat_(cross_cat) {
        LOOM_B200_CHECK(LB_NAME(set_hyper_prior)(cross_cat_.handle(), &prior));
    }
    void run(uint64_t seed) { LOOM_B200_CHECK(LB_NAME(hyper_run)(cross_cat_.handle(), seed)); }

private:
    CrossCat &cross_cat_;
};

// ---------------------------------------------------------------------------------------------
// Schedules (src/schedules.hpp:41-321) and the multi-pass driver (src/loom.cc:167-450).  The
// schedules depend only on counts, so a whole batch of ADD / REMOVE decisions is made up front and
// handed to the device as one action list; the batch kernels run at the boundary in loom's order.
struct ScheduleConfig {           // protobuf::Config::Schedule, defaults of loom/config.py:37-43
    double extra_passes = 500.0, small_data_size = 4e3, big_data_size = 1e9;
    uint64_t max_reject_iters = 100;
};

class AnnealingSchedule {         // REMOVE : ADD = N : (1 + N)
public:
    explicit AnnealingSchedule(double extra_passes) { set_extra_passes(extra_passes); state_ = 2 * add_rate_; }
    void set_extra_passes(double n) { add_rate_ = 1.0 + n; remove_rate_ = n; }
    bool next_action_is_add() {
        if (state_ >= 0) { state_ -= remove_rate_; return true; }
        state_ += add_rate_;
        return false;
    }
    void load(double state) { state_ = state; }     // Checkpoint.Schedule.annealing_state (schedules.hpp:88-96)
    double state() const { return state_; }

private:
    double add_rate_ = 1, remove_rate_ = 0, state_ = 0;
};

class AcceleratingSchedule {      // piecewise linear in log(rows)-log(passes)
public:
    explicit AcceleratingSchedule(const ScheduleConfig &c) : c_(c) {}
    double extra_passes(uint64_t row_count) const {
        const double n = (double)row_count;
        if (n <= c_.small_data_size) return c_.extra_passes;
        if (n > c_.big_data_size) return 0.0;
        const double power = std::log(c_.big_data_size / n) / std::log(c_.big_data_size / c_.small_data_size);
        const double passes = std::pow(1.0 + c_.extra_passes, power);
        return passes > 2.0 ? passes - 1.0 : 0.0;
    }

private:
    ScheduleConfig c_;
};

class BatchingSchedule {          // a batch ends when every previously assigned row was replaced
public:
    void add() { ++fresh_; }
    void remove() { --stale_; }
    bool test() {
        if (stale_ == 0 && fresh_) { stale_ = fresh_; fresh_ = 0; return true; }
        return false;
    }
    void load(uint64_t row_count) { stale_ = row_count; fresh_ = 0; }   // schedules.hpp:184-193
    uint64_t row_count() const { return stale_ + fresh_; }

private:
    uint64_t stale_ = 0, fresh_ = 0;
};

class KernelDisablingSchedule {
public:
    explicit KernelDisablingSchedule(uint64_t max_reject_iters) : max_(max_reject_iters) {}
    void run(bool accepted) { rejects_ = accepted ? 0 : rejects_ + 1; }
    bool test() const { return rejects_ <= max_; }
    void load(uint64_t reject_iters) { rejects_ = reject_iters; }      // schedules.hpp:224-232
    uint64_t reject_iters() const { return rejects_; }

private:
    uint64_t max_, rejects_ = 0;
};

// Loom::infer_multi_pass for one chain.  `kind` / `hyper` may be null (kernel off).  Rows are the
// positions of the loaded block; the assigned rows form a window of the cyclic stream
// (src/stream_interval.hpp:36-140).  Returns the number of batches.
inline uint64_t infer_multi_pass(CrossCat &cross_cat, CatKernel &cat, KindKernel *kind, HyperKernel *hyper,
                                 const ScheduleConfig &config, uint64_t seed) {
    const uint64_t n_rows = cross_cat.row_count();
    AcceleratingSchedule accelerating(config);
    AnnealingSchedule annealing(accelerating.extra_passes(0));
    BatchingSchedule batching;
    KernelDisablingSchedule disabling(config.max_reject_iters);
    uint64_t assigned = 0, front = 0, back = 0, batches = 0;
    while (assigned != n_rows) {
        if (annealing.next_action_is_add()) {
            cat.add_row(front);
            front = (front + 1) % n_rows;
            ++assigned;
            batching.add();
        } else {
            cat.remove_row(back);
            back = (back + 1) % n_rows;
            --assigned;
            batching.remove();
        }
        if (batching.test()) {
            cat.wait();
            ++batches;
            annealing.set_extra_passes(accelerating.extra_passes(assigned));
            if (kind) {
                disabling.run(kind->try_run());
                if (!disabling.test()) kind = nullptr;   // infer_kind_structure returns false (loom.cc:203-209)
            }
            if (hyper) hyper->run(seed * 7919 + batches);
        }
    }
    cat.wait();
    return batches;
}

}  // namespace b200
}  // namespace loom
