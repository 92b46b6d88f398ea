// TEST INFRASTRUCTURE.  Stand-in for the protoc-generated <loom/schema.pb.h> (protoc is not in this image) covering the
// messages the reference's value path touches -- ProductValue, ProductValue.Observed, ProductValue.Diff, Row
// (src/schema.proto) -- as plain structs with the accessor names protoc generates, over a minimal RepeatedField.
// ProductModel / HyperPrior exist only so that src/protobuf.hpp's Fields<> declarations parse.  Written from
// schema.proto's field list and protobuf's public C++ API; nothing of the reference is copied.
#pragma once
#include <cstddef>
#include <cstdint>
#include <exception>
#include <ostream>
#include <string>
#include <vector>
namespace google { namespace protobuf {
struct FatalException : std::exception {};          // what the reference's try blocks around protobuf calls catch
template <class T> class RepeatedField {
    std::vector<T> v_;

public:
    typedef typename std::vector<T>::iterator iterator;
    typedef typename std::vector<T>::const_iterator const_iterator;
    int size() const { return (int)v_.size(); }
    const T &Get(int i) const { return v_[i]; }
    void Set(int i, const T &x) { v_[i] = x; }
    void Add(const T &x) { v_.push_back(x); }
    void AddAlreadyReserved(const T &x) { v_.push_back(x); }
    void Reserve(int n) { v_.reserve(n); }
    void Clear() { v_.clear(); }
    void CopyFrom(const RepeatedField &o) { v_ = o.v_; }
    iterator begin() { return v_.begin(); }
    iterator end() { return v_.end(); }
    const_iterator begin() const { return v_.begin(); }
    const_iterator end() const { return v_.end(); }
    bool operator==(const RepeatedField &o) const { return v_ == o.v_; }
};
// std::vector<bool> has proxy iterators; the reference writes `*it = true` and reads `*it++` through protobuf's plain
// bool arrays, so booleans are stored as bytes
template <> class RepeatedField<bool> {
    std::vector<unsigned char> v_;

public:
    typedef bool *iterator;
    typedef const bool *const_iterator;
    static_assert(sizeof(bool) == 1, "bool is one byte");
    int size() const { return (int)v_.size(); }
    bool Get(int i) const { return v_[i] != 0; }
    void Set(int i, bool x) { v_[i] = x; }
    void Add(bool x) { v_.push_back(x); }
    void AddAlreadyReserved(bool x) { v_.push_back(x); }
    void Reserve(int n) { v_.reserve(n); }
    void Clear() { v_.clear(); }
    iterator begin() { return reinterpret_cast<bool *>(v_.data()); }
    iterator end() { return begin() + v_.size(); }
    const_iterator begin() const { return reinterpret_cast<const bool *>(v_.data()); }
    const_iterator end() const { return begin() + v_.size(); }
    bool operator==(const RepeatedField &o) const { return v_ == o.v_; }
};
template <class T> inline std::ostream &operator<<(std::ostream &os, const RepeatedField<T> &f) {   // error messages only
    os << "[";
    for (int i = 0; i < f.size(); ++i) os << (i ? "," : "") << f.Get(i);
    return os << "]";
}
}}  // namespace google::protobuf

namespace protobuf { namespace loom {
using google::protobuf::RepeatedField;

struct ProductValue {
    struct Observed {
        enum Sparsity { NONE = 0, SPARSE = 1, DENSE = 2, ALL = 3 };          // src/schema.proto:76-81
        static const char *Sparsity_Name(Sparsity s) {
            static const char *names[] = {"NONE", "SPARSE", "DENSE", "ALL"};
            return names[(int)s];
        }
        Sparsity sparsity_ = NONE;
        RepeatedField<bool> dense_;
        RepeatedField<uint32_t> sparse_;
        Sparsity sparsity() const { return sparsity_; }
        void set_sparsity(Sparsity s) { sparsity_ = s; }
        const RepeatedField<bool> &dense() const { return dense_; }
        RepeatedField<bool> *mutable_dense() { return &dense_; }
        bool dense(int i) const { return dense_.Get(i); }
        void add_dense(bool x) { dense_.Add(x); }
        void set_dense(int i, bool x) { dense_.Set(i, x); }
        int dense_size() const { return dense_.size(); }
        bool has_dense() const { return dense_.size() != 0; }
        void clear_dense() { dense_.Clear(); }
        const RepeatedField<uint32_t> &sparse() const { return sparse_; }
        RepeatedField<uint32_t> *mutable_sparse() { return &sparse_; }
        uint32_t sparse(int i) const { return sparse_.Get(i); }
        void add_sparse(uint32_t x) { sparse_.Add(x); }
        int sparse_size() const { return sparse_.size(); }
        void clear_sparse() { sparse_.Clear(); }
        void Clear() { sparsity_ = NONE; dense_.Clear(); sparse_.Clear(); }
    };
    struct Diff;
    Observed observed_;
    RepeatedField<bool> booleans_;
    RepeatedField<uint32_t> counts_;
    RepeatedField<float> reals_;
    const Observed &observed() const { return observed_; }
    Observed *mutable_observed() { return &observed_; }
#define LB_STUB_REPEATED(T, name)                                            \
    const RepeatedField<T> &name() const { return name##_; }                 \
    RepeatedField<T> *mutable_##name() { return &name##_; }                  \
    T name(int i) const { return name##_.Get(i); }                           \
    void add_##name(T x) { name##_.Add(x); }                                 \
    void set_##name(int i, T x) { name##_.Set(i, x); }                       \
    int name##_size() const { return name##_.size(); }                       \
    void clear_##name() { name##_.Clear(); }
    LB_STUB_REPEATED(bool, booleans)
    LB_STUB_REPEATED(uint32_t, counts)
    LB_STUB_REPEATED(float, reals)
#undef LB_STUB_REPEATED
    void Clear() { observed_.Clear(); booleans_.Clear(); counts_.Clear(); reals_.Clear(); }
    std::string ShortDebugString() const { return "<ProductValue>"; }
};
struct ProductValue::Diff {
    ProductValue pos_, neg_;
    RepeatedField<uint32_t> tares_;
    const ProductValue &pos() const { return pos_; }
    ProductValue *mutable_pos() { return &pos_; }
    const ProductValue &neg() const { return neg_; }
    ProductValue *mutable_neg() { return &neg_; }
    const RepeatedField<uint32_t> &tares() const { return tares_; }
    RepeatedField<uint32_t> *mutable_tares() { return &tares_; }
    uint32_t tares(int i) const { return tares_.Get(i); }
    void add_tares(uint32_t x) { tares_.Add(x); }
    int tares_size() const { return tares_.size(); }
    void clear_tares() { tares_.Clear(); }
    void Clear() { pos_.Clear(); neg_.Clear(); tares_.Clear(); }
};
struct Row {
    uint64_t id_ = 0;
    ProductValue::Diff diff_;
    uint64_t id() const { return id_; }
    void set_id(uint64_t x) { id_ = x; }
    const ProductValue::Diff &diff() const { return diff_; }
    ProductValue::Diff *mutable_diff() { return &diff_; }
    void Clear() { id_ = 0; diff_.Clear(); }
};
struct Assignment {                                   // src/schema.proto:160-163
    uint64_t rowid_ = 0;
    std::vector<uint32_t> groupids_;
    uint64_t rowid() const { return rowid_; }
    void set_rowid(uint64_t x) { rowid_ = x; }
    int groupids_size() const { return (int)groupids_.size(); }
    uint32_t groupids(int i) const { return groupids_[i]; }
    void add_groupids(uint32_t g) { groupids_.push_back(g); }
    void clear_groupids() { groupids_.clear(); }
};
struct Config {                                     // Config.Schedule as src/schedules.hpp reads it, Config.Kernels as src/kind_kernel.cc:40-43 does
    struct Schedule {
        double extra_passes_ = 0, small_data_size_ = 0, big_data_size_ = 0, checkpoint_period_sec_ = 1e9;
        uint32_t max_reject_iters_ = 0;
        double extra_passes() const { return extra_passes_; }
        double small_data_size() const { return small_data_size_; }
        double big_data_size() const { return big_data_size_; }
        uint32_t max_reject_iters() const { return max_reject_iters_; }
        double checkpoint_period_sec() const { return checkpoint_period_sec_; }
    };
    struct Kernels {
        struct Hyper { bool run_ = true, parallel_ = false; bool run() const { return run_; } bool parallel() const { return parallel_; } } hyper_;
        const Hyper &hyper() const { return hyper_; }
        struct Cat { uint32_t empty_group_count_ = 1; uint32_t empty_group_count() const { return empty_group_count_; } } cat_;
        struct Kind {
            uint32_t empty_kind_count_ = 32, iterations_ = 32;
            bool score_parallel_ = false;
            uint32_t empty_kind_count() const { return empty_kind_count_; }
            uint32_t iterations() const { return iterations_; }
            bool score_parallel() const { return score_parallel_; }
        } kind_;
        const Cat &cat() const { return cat_; }
        const Kind &kind() const { return kind_; }
    };
};
struct Checkpoint {                                 // Checkpoint.Schedule: what the schedules load and dump
    struct Schedule {
        double annealing_state_ = 0;
        uint64_t row_count_ = 0;
        uint32_t reject_iters_ = 0;
        double annealing_state() const { return annealing_state_; }
        void set_annealing_state(double x) { annealing_state_ = x; }
        uint64_t row_count() const { return row_count_; }
        void set_row_count(uint64_t x) { row_count_ = x; }
        uint32_t reject_iters() const { return reject_iters_; }
        void set_reject_iters(uint32_t x) { reject_iters_ = x; }
    };
};
// only so that the Fields<> specialisations of src/protobuf.hpp and the load / dump code of src/product_mixture.cc parse:
// per-model-type lists of empty messages
struct MessageListStub;
struct MessageStub {
    const MessageListStub &alphas() const;            // DirichletDiscrete.Shared.alphas: product_model.cc:60 reads its size
};
struct MessageListStub {
    int size() const { return 0; }
    const MessageStub &Get(int) const { static MessageStub m; return m; }
    MessageStub *Add() { static MessageStub m; return &m; }
    const MessageStub *begin() const { return nullptr; }
    const MessageStub *end() const { return nullptr; }
};
inline const MessageListStub &MessageStub::alphas() const { static MessageListStub l; return l; }
struct FieldListStub {
    MessageListStub list_;
    const MessageListStub &bb() const { return list_; } const MessageListStub &dd() const { return list_; }
    const MessageListStub &dpd() const { return list_; } const MessageListStub &gp() const { return list_; }
    const MessageListStub &nich() const { return list_; }
    const MessageStub &bb(int) const { return list_.Get(0); } const MessageStub &dd(int) const { return list_.Get(0); }
    const MessageStub &dpd(int) const { return list_.Get(0); } const MessageStub &gp(int) const { return list_.Get(0); }
    const MessageStub &nich(int) const { return list_.Get(0); }
    int bb_size() const { return 0; } int dd_size() const { return 0; } int dpd_size() const { return 0; }
    int gp_size() const { return 0; } int nich_size() const { return 0; }
    MessageListStub *mutable_bb() { return &list_; } MessageListStub *mutable_dd() { return &list_; }
    MessageListStub *mutable_dpd() { return &list_; } MessageListStub *mutable_gp() { return &list_; }
    MessageListStub *mutable_nich() { return &list_; }
};
struct ProductModel {
    struct Shared : FieldListStub {
        MessageStub clustering_;
        const MessageStub &clustering() const { return clustering_; }
        MessageStub *mutable_clustering() { return &clustering_; }
    };
    struct Group : FieldListStub {
        uint64_t count_ = 0;
        uint64_t count() const { return count_; }
        void set_count(uint64_t x) { count_ = x; }
        void Clear() { count_ = 0; }
    };
};
typedef ProductModel::Shared ProductModel_Shared;      // protoc's flat names for the nested messages
typedef ProductModel::Group ProductModel_Group;
struct HyperPrior {                                   // src/schema.proto:34-71: one grid per hyperparameter
    typedef google::protobuf::RepeatedField<float> Grid;
    struct PitmanYorPoint { float alpha_ = 0, d_ = 0; float alpha() const { return alpha_; } float d() const { return d_; } };
    struct PitmanYorGrid {
        std::vector<PitmanYorPoint> points;
        size_t size() const { return points.size(); }
        const PitmanYorPoint &Get(int i) const { return points[i]; }
    };
#define LB_STUB_GRID(name) Grid name##_; const Grid &name() const { return name##_; } int name##_size() const { return name##_.size(); } float name(int i) const { return name##_.Get(i); }
    struct BetaBernoulli { LB_STUB_GRID(alpha) LB_STUB_GRID(beta) };
    struct DirichletDiscrete { LB_STUB_GRID(alpha) };
    struct DirichletProcessDiscrete { LB_STUB_GRID(gamma) LB_STUB_GRID(alpha) };
    struct GammaPoisson { LB_STUB_GRID(alpha) LB_STUB_GRID(inv_beta) };
    struct BetaNegativeBinomial { LB_STUB_GRID(alpha) LB_STUB_GRID(beta) LB_STUB_GRID(r) };
    struct NormalInverseChiSq { LB_STUB_GRID(mu) LB_STUB_GRID(kappa) LB_STUB_GRID(sigmasq) LB_STUB_GRID(nu) };
#undef LB_STUB_GRID
    PitmanYorGrid topology_, clustering_;             // EMPTY grids: the kernel keeps the hypers (new kinds copy kind 0's, kind_kernel.cc:167-171)
    BetaBernoulli bb_; DirichletDiscrete dd_; DirichletProcessDiscrete dpd_; GammaPoisson gp_; NormalInverseChiSq nich_;
    const PitmanYorGrid &topology() const { return topology_; }
    const PitmanYorGrid &clustering() const { return clustering_; }
    const BetaBernoulli &bb() const { return bb_; }
    const DirichletDiscrete &dd() const { return dd_; }
    const DirichletProcessDiscrete &dpd() const { return dpd_; }
    const GammaPoisson &gp() const { return gp_; }
    const NormalInverseChiSq &nich() const { return nich_; }
};
struct CrossCat {                                      // what CrossCat::model_load / model_dump touch (never called by the harnesses)
    struct Kind {
        std::vector<uint32_t> featureids_;
        ProductModel::Shared product_model_;
        size_t featureids_size() const { return featureids_.size(); }
        uint32_t featureids(int i) const { return featureids_[i]; }
        void add_featureids(uint32_t f) { featureids_.push_back(f); }
        const ProductModel::Shared &product_model() const { return product_model_; }
        ProductModel::Shared *mutable_product_model() { return &product_model_; }
    };
    std::vector<Kind> kinds_;
    MessageStub topology_;
    HyperPrior hyper_prior_;
    size_t kinds_size() const { return kinds_.size(); }
    const Kind &kinds(int i) const { return kinds_[i]; }
    Kind *add_kinds() { kinds_.emplace_back(); return &kinds_.back(); }
    const MessageStub &topology() const { return topology_; }
    MessageStub *mutable_topology() { return &topology_; }
    const HyperPrior &hyper_prior() const { return hyper_prior_; }
    HyperPrior *mutable_hyper_prior() { return &hyper_prior_; }
};
inline std::ostream &operator<<(std::ostream &os, const ProductValue &v) { return os << v.ShortDebugString(); }
}}  // namespace protobuf::loom
