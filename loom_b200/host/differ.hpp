// differ.hpp -- loom::Differ (src/differ.hpp:36-125, src/differ.cc) re-hosted on the C ABIs of
// include/loom_b200.h (lb_differ_*: histograms and diff encoding on the device) and
// include/loom_b200_io.h (loom's files): what `loom_tare` and `loom_sparsify` run.
//
// The rows file is decoded window by window into the device-resident columnar block, so a stream
// larger than one block is summarised / compressed in pieces (the summaries add up across windows;
// compress_rows is row-local).
#pragma once
#include <string>
#include <vector>

#include "loom.hpp"

namespace loom {
namespace b200 {

class Differ {
public:

    // Differ::Differ(schema) / Differ(schema, tare) (src/differ.cc:56-86); the schema is a schema row (src/tare.cc:52-55)
    explicit Differ(const char *schema_row_in, int device = 0) : cross_cat_(device) {
        if (const char *w = std::getenv("LOOM_B200_WINDOW_ROWS")) window_rows_ = std::max<uint64_t>(1, std::strtoull(w, nullptr, 10));
        LOOM_B200_IO(lbio_model_from_schema_row(schema_row_in, &schema_));
        LOOM_B200_IO(lbio_model_get(schema_, &view_));
        cross_cat_.model_load(view_.n_features, view_.specs, view_.aux, view_.n_aux, view_.n_kinds,
                              view_.featureid_to_kindid, view_.kind_py, view_.topology_py, 1);
        tare_observed_.assign(view_.n_features, 0);
        tare_values_.assign(view_.n_features, 0);
        LOOM_B200_CHECK(LB_NAME(differ_reset)(cross_cat_.handle()));
    }
    ~Differ() { lbio_model_free(schema_); }
    Differ(const Differ &) = delete;

    // set_tare from a tares file: none = the blank tare, more than one is the reference's TODO (src/sparsify.cc:62-68)
    void tares_load(const char *tares_in) {
        int32_t n_tares = 0;
        LOOM_B200_IO(lbio_tares_read(tares_in, schema_, tare_observed_.data(), tare_values_.data(), &n_tares));
    }

    // Differ::add_rows (src/differ.cc:93-127) + _make_tare
    void add_rows(const char *rows_in) {
        const std::vector<uint8_t> no_tare_obs(view_.n_features, 0);
        const std::vector<uint32_t> no_tare_val(view_.n_features, 0);
        for (uint64_t first = 0;; first += window_rows_) {
            lbio_row_block rows;
            LOOM_B200_IO(lbio_rows_read(rows_in, schema_, first, window_rows_, 0, &rows));
            const uint64_t n = rows.n_rows;
            for (uint64_t r = 0; r < n; ++r)
                if (rows.row_has_tare[r]) LOOM_B200_ERROR("row is already sparsified");
            if (n) {
                upload(rows, no_tare_obs, no_tare_val);
                LOOM_B200_CHECK(LB_NAME(differ_add_rows)(cross_cat_.handle(), nullptr, nullptr));
            }
            lbio_rows_free(&rows);
            if (n < window_rows_) break;
            need_file(rows_in);
        }
        LOOM_B200_CHECK(LB_NAME(differ_make_tare)(cross_cat_.handle(), nullptr, nullptr, tare_observed_.data(), tare_values_.data()));
    }

    bool has_tare() const {
        for (uint8_t o : tare_observed_)
            if (o) return true;
        return false;
    }

    // tare.cc:64-69
    void tares_dump(const char *tares_out) const {
        LOOM_B200_IO(lbio_tares_write(tares_out, schema_, tare_observed_.data(), tare_values_.data()));
    }

    // Differ::compress_rows (src/differ.cc:148-189)
    void compress_rows(const char *rows_in, const char *diffs_out) {
        if (std::string(rows_in) == std::string(diffs_out) && std::string(rows_in) != "-" && std::string(rows_in) != "-.gz")
            LOOM_B200_ERROR("in-place sparsify is not supported");
        const bool tared = has_tare();
        const std::vector<uint8_t> no_tare_obs(view_.n_features, 0);
        const std::vector<uint32_t> no_tare_val(view_.n_features, 0);
        lbio_rows_writer *writer = nullptr;
        LOOM_B200_IO(lbio_rows_writer_open(diffs_out, &writer));
        for (uint64_t first = 0;; first += window_rows_) {
            lbio_row_block rows;
            LOOM_B200_IO(lbio_rows_read(rows_in, schema_, first, window_rows_, 0, &rows));
            const uint64_t n = rows.n_rows;
            if (n) {
                upload(rows, no_tare_obs, no_tare_val);
                uint64_t n_pos = 0, n_neg = 0;
                LOOM_B200_CHECK(LB_NAME(differ_compress_rows)(cross_cat_.handle(), tare_observed_.data(), tare_values_.data(), &n_pos, &n_neg));
                std::vector<uint64_t> pos_ptr(n + 1), neg_ptr(n + 1);
                std::vector<uint32_t> pos_feature(n_pos + 1), pos_value(n_pos + 1), neg_feature(n_neg + 1);
                std::vector<uint8_t> has_tare(n, tared ? 1 : 0);
                LOOM_B200_CHECK(LB_NAME(differ_fetch)(cross_cat_.handle(), pos_ptr.data(), pos_feature.data(), pos_value.data(),
                                                      neg_ptr.data(), neg_feature.data()));
                lbio_row_block out;
                std::memset(&out, 0, sizeof(out));
                out.n_rows = n;
                out.rowids = rows.rowids;
                out.row_has_tare = has_tare.data();
                out.pos_ptr = pos_ptr.data();
                out.pos_feature = pos_feature.data();
                out.pos_value = pos_value.data();
                out.neg_ptr = neg_ptr.data();
                out.neg_feature = neg_feature.data();
                LOOM_B200_IO(lbio_rows_writer_write(writer, schema_, &out, tare_values_.data()));
            }
            lbio_rows_free(&rows);
            if (n < window_rows_) break;
            need_file(rows_in);
        }
        LOOM_B200_IO(lbio_rows_writer_close(writer));
    }

private:
    // further windows re-open the stream at their first row
    void need_file(const char *rows_in) const {
        if (std::string(rows_in) == "-" || std::string(rows_in) == "-.gz")
            LOOM_B200_ERROR("a piped rows stream must fit one block of %llu rows", (unsigned long long)window_rows_);
    }

    void upload(const lbio_row_block &rows, const std::vector<uint8_t> &tare_obs, const std::vector<uint32_t> &tare_val) {
        LOOM_B200_CHECK(LB_NAME(set_rows_diff)(cross_cat_.handle(), rows.n_rows, tare_obs.data(), tare_val.data(), rows.row_has_tare,
                                               rows.pos_ptr, rows.pos_feature, rows.pos_value, rows.neg_ptr, rows.neg_feature));
        cross_cat_.rows_attached(rows.n_rows);
    }

    uint64_t window_rows_ = 1ull << 22;     // rows per device block (LOOM_B200_WINDOW_ROWS overrides)
    CrossCat cross_cat_;
    lbio_model *schema_ = nullptr;
    lbio_model_view view_;
    std::vector<uint8_t> tare_observed_;
    std::vector<uint32_t> tare_values_;
};

}  // namespace b200
}  // namespace loom
