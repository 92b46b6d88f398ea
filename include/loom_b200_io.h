/*
 * loom_b200_io.h -- C ABI of the file side of the drop-in: loom's protobuf files
 * (config, model, groups, assignments, rows, tares, checkpoint, log) decoded straight
 * into the arrays include/loom_b200.h consumes, and encoded back.
 *
 * This is the process/file boundary of SURVEY.md 8(b1) -- what loom.runner hands to
 * `loom_infer` (src/infer.cc:68-81) -- and row n1 of 8(f): the unzip / parse / split
 * stages of src/cat_pipeline.cc:63-87 become one columnar decode.  Host code only
 * (libloomb200_io.so: g++ + zlib, no CUDA, no protobuf runtime): the wire format is
 * read and written directly.
 *
 * Framing (src/protobuf_stream.hpp:100-144, 258-282): `*.pb[.gz]` holds one message,
 * `*.pbs[.gz]` a stream of messages each preceded by its size as a little-endian
 * uint32; a name ending in `.gz` is gzip-compressed; `-` / `-.gz` mean stdin / stdout.
 *
 * Field numbers of loom's own messages are those of src/schema.proto (cited at each
 * codec).  The per-feature Shared / Group messages belong to the un-vendored
 * distributions 2.0.28 (`distributions/io/schema.proto`, not in the reference tree):
 * their field numbers here are a stand-in restated from that package's published
 * schema and are NOT verifiable offline; they live in one table in
 * loom_b200/io/lb_io.cc.
 *
 * All functions return 0 on success; lbio_last_error() gives the message.  Blocks
 * returned through *_read own their arrays until the matching *_free.
 */
#ifndef LOOM_B200_IO_H_
#define LOOM_B200_IO_H_

#include <stdint.h>

#include "loom_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

const char *lbio_last_error(void);

/* ------------------------------------------------------------------------- */
/* protobuf::Config (src/schema.proto:167-229; defaults loom/config.py:34-74). */
typedef struct {
    uint64_t seed;
    float extra_passes, small_data_size, big_data_size;
    uint32_t max_reject_iters;
    float checkpoint_period_sec;
    uint32_t cat_empty_group_count, cat_row_queue_capacity, cat_parser_threads;
    int32_t hyper_run, hyper_parallel;
    uint32_t kind_iterations, kind_empty_kind_count, kind_row_queue_capacity, kind_parser_threads;
    int32_t kind_score_parallel;
    uint32_t posterior_enum_sample_count, posterior_enum_sample_skip;
    uint64_t generate_row_count;
    float generate_density;
    uint32_t generate_sample_skip;
    float target_mem_bytes;
    int32_t has_query, query_parallel;
} lbio_config;

void lbio_config_defaults(lbio_config *out);               /* loom/config.py:34-74 */
int lbio_config_read(const char *path, lbio_config *out);   /* protobuf_load<Config>, src/infer.cc:91 */
int lbio_config_write(const char *path, const lbio_config *in);

/* ------------------------------------------------------------------------- */
/* protobuf::Checkpoint (src/schema.proto:234-253), read and written by
 * Loom::infer_multi_pass (src/loom.cc:178-214). */
typedef struct {
    int32_t finished;
    uint64_t seed, tardis_iter;
    double annealing_state;            /* AnnealingSchedule::state_   src/schedules.hpp:88-96  */
    uint64_t schedule_row_count;       /* BatchingSchedule            src/schedules.hpp:184-193 */
    uint64_t reject_iters;             /* KernelDisablingSchedule     src/schedules.hpp:224-232 */
    uint64_t row_count;
    uint64_t unassigned_pos, assigned_pos;   /* StreamInterval, message indices (src/stream_interval.hpp:47-64) */
} lbio_checkpoint;

int lbio_checkpoint_read(const char *path, lbio_checkpoint *out);
int lbio_checkpoint_write(const char *path, const lbio_checkpoint *in);

/* ------------------------------------------------------------------------- */
/* protobuf::CrossCat, the model file (src/schema.proto:126-135; CrossCat::model_load /
 * model_dump src/cross_cat.cc:50-135; ProductModel::load src/product_model.cc:33-93).
 * The view is exactly the argument list of lb_set_model plus the hyper-prior grids
 * (lb_set_hyper_prior) and the DPD dictionaries the row decoder needs. */
typedef struct lbio_model lbio_model;

typedef struct {
    int32_t n_features;
    const lb_feature_spec *specs;          /* loom's global feature order */
    const float *aux;                      /* DD alphas / DPD betas */
    int32_t n_aux;
    int32_t n_kinds;
    const uint32_t *featureid_to_kindid;   /* [F] */
    const float *kind_py;                  /* [K][2] */
    float topology_py[2];
    int32_t has_hyper_prior;
    lb_hyper_prior hyper_prior;            /* pointers into the model object */
    int32_t n_dpd;
    const int32_t *dpd_offsets;            /* [n_dpd + 1] into dpd_values */
    const uint32_t *dpd_values;            /* Shared.values of each DPD feature, ascending feature id:
                                              dictionary index i of a feature <-> raw value */
} lbio_model_view;

int lbio_model_read(const char *path, lbio_model **out);
int lbio_model_create(int32_t n_features, const lb_feature_spec *specs, const float *aux, int32_t n_aux,
                      int32_t n_kinds, const uint32_t *featureid_to_kindid, const float *kind_py,
                      const float *topology_py, const lb_hyper_prior *hyper_prior /* may be NULL */,
                      const int32_t *dpd_offsets /* may be NULL: identity dictionaries */,
                      const uint32_t *dpd_values, lbio_model **out);
/* A model from a schema row (ValueSchema::load, src/product_value.hpp:125-130; written by loom/format.py:79-101): the
 * file only holds the number of boolean, count and real columns, which is all `loom_tare` / `loom_sparsify` need
 * (src/tare.cc:52-55).  Count columns are typed GammaPoisson (any 32-bit value), one kind. */
int lbio_model_from_schema_row(const char *path, lbio_model **out);
int lbio_model_get(const lbio_model *m, lbio_model_view *out);
int lbio_model_write(const lbio_model *m, const char *path);
void lbio_model_free(lbio_model *m);

/* ------------------------------------------------------------------------- */
/* Rows: a stream of protobuf::Row{id, diff{pos, neg, tares}} (src/schema.proto:74-99,
 * 153-156), decoded by read_value_all / _dense / _sparse (src/product_value.hpp:586-767)
 * into the CSR form lb_set_rows_diff takes -- one pass replaces the unzip, parse and
 * ValueSplitter::split stages (src/cat_pipeline.cc:63-87).  Feature f of the model is
 * position f of the concatenated booleans | counts | reals space.  pos_value is the raw
 * cell value: BB 0/1, DD category, DPD DICTIONARY INDEX, GP count, NICH float bits.
 * Rows [first_row, first_row + max_rows) of the file are decoded (max_rows = 0: to the
 * end), so a long stream can be ingested window by window.  n_threads <= 0 = all cores
 * (gunzip and framing are serial, message decoding is parallel). */
typedef struct {
    uint64_t n_rows;
    uint64_t *rowids;        /* [R]   Row.id */
    uint8_t *row_has_tare;   /* [R]   diff.tares contains tare 0 */
    uint64_t *pos_ptr;       /* [R+1] */
    uint32_t *pos_feature;   /* [nnz] ascending inside a row */
    uint32_t *pos_value;     /* [nnz] */
    uint64_t *neg_ptr;       /* [R+1] */
    uint32_t *neg_feature;
    uint64_t stream_rows;    /* messages in the whole file (InFile::stream_stats), when read to the end */
    void *owner;
} lbio_row_block;

int lbio_rows_read(const char *path, const lbio_model *m, uint64_t first_row, uint64_t max_rows,
                   int32_t n_threads, lbio_row_block *out);
void lbio_rows_free(lbio_row_block *b);

/* The same with OPEN dictionaries: a DirichletProcessDiscrete value that the model's dictionary does
 * not hold is appended to it in order of first appearance in the stream, and its beta is drawn by
 * stick breaking, beta_v = beta0 * Beta(1, gamma), beta0 -= beta_v -- what distributions'
 * DPD Shared.add_value does when inference first meets a value (loom's init models
 * start with EMPTY dictionaries, loom/generate.py:103-107).  Here every value of the file is
 * materialised at ingest, so the engine's DPD tables have a fixed shape; the end state (all rows
 * assigned) has the same law.  A DPD feature no row observes gets the single value 0 so that no
 * feature has an empty table.  Draws: Philox stream 4, counter a = feature id, b = draw index.
 * The model object is MODIFIED: fetch lbio_model_get again afterwards.  *n_new = values appended. */
int lbio_rows_read_open(const char *path, lbio_model *m, uint64_t seed, int32_t n_threads, lbio_row_block *out,
                        int32_t *n_new);
/* Give `dst` the dictionary of `src`: values of src that dst lacks are appended in src's order with
 * dst's own stick-breaking draws (several samples share one decoded row block but have their own
 * hypers).  Fails unless dst's dictionary is then identical to src's. */
int lbio_model_adopt_dictionaries(lbio_model *dst, const lbio_model *src, uint64_t seed);
/* InFile::stream_stats (src/protobuf_stream.hpp:156-184) */
int lbio_stream_stats(const char *path, uint64_t *message_count, uint32_t *max_message_size);
/* CrossCat::get_sorted_groupids for one kind (src/cross_cat.cc:212-236): the packed ids of the non-empty groups, largest
 * first -- the order groups are dumped in (ProductMixture::dump) and assignments renumbered to (Assignments::dump).  The
 * reference sorts with std::sort and a `counts[x] > counts[y]` comparator, which leaves the order of EQUAL counts to the
 * library's introsort; this entry point runs the same call on the same initial sequence (ascending packed id), so a
 * build against the same libstdc++ orders tied groups identically (pinned by tests/golden/reference_sorted_groupids.json).
 * order_out has room for n_groups entries; *n_out = number of non-empty groups. */
int lbio_sorted_groupids(const uint32_t *counts, int32_t n_groups, uint32_t *order_out, int32_t *n_out);
/* shuffle_stream (src/shuffle.hpp:38-86): a seeded permutation of the messages of a stream file, produced in chunks
 * of at most target_mem_bytes (the result does not depend on the chunk size).  What `loom_shuffle` runs. */
int lbio_shuffle_stream(const char *messages_in, const char *shuffled_out, uint64_t seed, double target_mem_bytes);
/* the inverse, for tests / generators: rows as Row messages, dense rows as diff{pos = row, neg = NONE} */
int lbio_rows_write(const char *path, const lbio_model *m, const lbio_row_block *rows);
/* The same block by block (Differ::compress_rows streams, src/differ.cc:166-181).  tare_values [F] (may be NULL): the
 * tare row's cell values, written as the values of neg the way _abs_to_rel_type does (src/differ.cc:277-286). */
typedef struct lbio_rows_writer lbio_rows_writer;
int lbio_rows_writer_open(const char *path, lbio_rows_writer **out);
int lbio_rows_writer_write(lbio_rows_writer *w, const lbio_model *m, const lbio_row_block *rows,
                           const uint32_t *tare_values);
int lbio_rows_writer_close(lbio_rows_writer *w);

/* Expand a decoded block into the dense columnar layout of lb_set_rows on the HOST
 * (cat8 [F8][R], wide [FW][R], mask [F][ceil(R/32)], caller-allocated, zeroed here).
 * lb_set_rows_diff does this on the device; this form exists for hosts that keep a
 * dense block and for checking the decoder. */
int lbio_rows_to_columns(const lbio_model *m, const lbio_row_block *rows, const uint8_t *tare_observed,
                         const uint32_t *tare_values, uint8_t *cat8, uint32_t *wide, uint32_t *mask);

/* Tares: a stream of ProductValue (CrossCat::tares_load, src/cross_cat.cc:137-146).  Only a
 * single tare row is supported, which is all loom's tare pass emits (src/differ.cc:93-189).
 * tare_observed [F], tare_values [F] as lb_set_rows_diff takes them; *n_tares = 0 or 1. */
int lbio_tares_read(const char *path, const lbio_model *m, uint8_t *tare_observed,
                    uint32_t *tare_values, int32_t *n_tares);
/* The same with OPEN dictionaries (see lbio_rows_read_open): a DirichletProcessDiscrete value of the tare row that
 * the model's dictionary lacks is appended with its stick-breaking beta -- the value a tared column holds in most
 * rows appears in no row's pos.  Call after lbio_rows_read_open; the model object is MODIFIED. */
int lbio_tares_read_open(const char *path, lbio_model *m, uint64_t seed, uint8_t *tare_observed,
                         uint32_t *tare_values, int32_t *n_tares, int32_t *n_new);
int lbio_tares_write(const char *path, const lbio_model *m, const uint8_t *tare_observed,
                     const uint32_t *tare_values);

/* ------------------------------------------------------------------------- */
/* Assignments: a stream of protobuf::Assignment{rowid, groupids[K]} (src/schema.proto:160-163;
 * Assignments::load / dump src/assignments.cc:50-100).  groupids is [R][K] row-major. */
typedef struct {
    uint64_t n_rows;
    int32_t n_kinds;
    uint64_t *rowids;
    uint32_t *groupids;
    void *owner;
} lbio_assign_block;

int lbio_assign_read(const char *path, lbio_assign_block *out);
void lbio_assign_free(lbio_assign_block *b);
int lbio_assign_write(const char *path, uint64_t n_rows, int32_t n_kinds, const uint64_t *rowids,
                      const uint32_t *groupids);

/* ------------------------------------------------------------------------- */
/* Groups of one kind: a stream of ProductModel.Group (src/schema.proto:113-121;
 * ProductMixture::load_step_1_of_3 / dump src/product_mixture.cc:739-761, 836-856) in the
 * table layout of lb_set_groups / lb_get_groups ([rows][n_groups], group index fastest;
 * the derived DD / DPD count_sum row is filled on read and ignored on write).
 * `featureids` are the kind's features in ascending id (n_kind_features of them). */
typedef struct {
    int32_t n_groups, n_int_rows, n_flt_rows;
    uint32_t *counts;
    uint32_t *ints;
    double *flts;
    void *owner;
} lbio_group_block;

int lbio_groups_read(const char *path, const lbio_model *m, const uint32_t *featureids,
                     int32_t n_kind_features, lbio_group_block *out);
void lbio_groups_free(lbio_group_block *b);
/* Writes groups order[0..n_out) (packed ids into the tables, whose row stride is
 * n_groups): loom dumps the non-empty groups in descending-count order
 * (CrossCat::get_sorted_groupids src/cross_cat.cc:212-236). */
int lbio_groups_write(const char *path, const lbio_model *m, const uint32_t *featureids,
                      int32_t n_kind_features, int32_t n_groups, const uint32_t *counts,
                      const uint32_t *ints, const double *flts, const uint32_t *order, int32_t n_out);

/* ------------------------------------------------------------------------- */
/* Log: a stream of LogMessage (src/schema.proto:257-322; Loom::log_metrics src/loom.cc:139-165,
 * Logger src/logger.hpp).  One call appends one message with summary + scores. */
typedef struct {
    uint32_t iter;
    float topology_py[2];
    int32_t n_kinds;               /* kinds with features */
    const float *kind_py;          /* [n_kinds][2] */
    const uint32_t *feature_counts, *category_counts;   /* [n_kinds] */
    float score, kl_divergence;
    uint64_t assigned_object_count;
    uint64_t cat_total_time_usec, hyper_total_time_usec, kind_total_time_usec;
    uint64_t kind_total_count, kind_change_count;
} lbio_log_entry;

typedef struct lbio_log lbio_log;
int lbio_log_open(const char *path, int32_t append, lbio_log **out);
int lbio_log_write(lbio_log *log, const lbio_log_entry *entry);
void lbio_log_close(lbio_log *log);

/* ------------------------------------------------------------------------- */
/* Posterior-enumeration samples: a stream of PosteriorEnum.Sample{kinds{featureids, groups{rowids}},
 * score} (src/schema.proto:326-338; Loom::dump_posterior_enum src/loom.cc:513-546) -- what
 * loom/test/test_posterior_enum.py reads back.  One call appends one sample:
 *   kind_feature_ptr [n_kinds + 1] into featureids; kind_group_ptr [n_kinds + 1] into the group list;
 *   group_row_ptr [n_groups + 1] into rowids. */
typedef struct lbio_samples lbio_samples;
int lbio_samples_open(const char *path, lbio_samples **out);
int lbio_samples_write(lbio_samples *s, int32_t n_kinds, const int32_t *kind_feature_ptr, const uint32_t *featureids,
                       const int32_t *kind_group_ptr, const uint64_t *group_row_ptr, const uint32_t *rowids,
                       float score);
int lbio_samples_close(lbio_samples *s);

/* ------------------------------------------------------------------------- */
/* Query protocol: a stream of Query.Request in, one Query.Response out per request
 * (src/schema.proto:342-418; QueryServer::serve src/query_server.cc:36-66).  A request's rows
 * are decoded into ONE CSR block (sample.data | score.data | entropy.conditional |
 * score_derivative.update_data, score_data...) in the order the sub-requests appear in the
 * message; the *_row fields index it (-1 = that sub-request is absent).  A
 * sub-request whose data does not fit the schema is reported through `invalid` exactly as the
 * reference's validate() does ("invalid request.score.data", ...) and is then marked absent. */
typedef struct lbio_query lbio_query;

typedef struct {
    const char *id;
    lbio_row_block rows;
    int32_t sample_row;              /* Query.Sample.Request :346-351 */
    const uint32_t *to_sample;       /* observed positions of to_sample */
    int32_t n_to_sample;
    uint32_t sample_count;
    int32_t score_row;               /* Query.Score.Request :360-363 */
    int32_t entropy_row;             /* Query.Entropy.Request :372-378: the conditional */
    int32_t n_row_sets, n_col_sets;
    const int32_t *set_ptr;          /* [n_row_sets + n_col_sets + 1] into set_features, row sets first */
    const uint32_t *set_features;    /* ascending feature positions of each set */
    uint32_t entropy_sample_count;
    int32_t update_row;              /* Query.ScoreDerivative.Request :388-393 */
    int32_t first_score_data_row, n_score_data;
    uint32_t row_limit;
    int32_t n_invalid;
    const char *const *invalid;      /* error strings produced while decoding */
} lbio_query_request;

typedef struct {
    int32_t n_errors;
    const char *const *errors;
    int32_t has_sample;              /* Sample.Response :352-355: n_samples values of the to_sample cells */
    int32_t n_samples;
    const uint64_t *sample_ptr;      /* [n_samples + 1] */
    const uint32_t *sample_feature;
    const uint32_t *sample_value;    /* raw cell values (DPD: dictionary index) */
    int32_t has_score;               /* Score.Response :364-367 */
    float score;
    int32_t has_entropy;             /* Entropy.Response :379-382: one estimate per (row set, col set) */
    int32_t n_entropy;
    const float *means, *variances;
    int32_t has_score_derivative;    /* ScoreDerivative.Response :394-398 */
    int32_t n_ids;
    const uint64_t *ids;
    const float *score_diffs;
} lbio_query_response;

int lbio_query_open(const char *requests_in, const char *responses_out, lbio_query **out);
/* *got = 0 at the end of the stream.  The request stays valid until the next call. */
int lbio_query_next(lbio_query *q, const lbio_model *m, lbio_query_request *req, int32_t *got);
/* writes Response{id of the current request, errors, ...} and flushes (the client waits on it) */
int lbio_query_respond(lbio_query *q, const lbio_model *m, const lbio_query_response *resp);
void lbio_query_close(lbio_query *q);

#ifdef __cplusplus
}
#endif
#endif /* LOOM_B200_IO_H_ */
