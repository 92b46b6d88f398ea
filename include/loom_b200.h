/*
 * loom_b200.h -- C ABI of the B200-native CrossCat inference core.
 *
 * This is the drop-in boundary for loom's row ("cat") / kind / hyper Gibbs hot
 * path.  The reference (priorknowledge/loom) has no FFI; its boundary is the
 * C++ class API (src/loom.hpp, src/cat_kernel.hpp, src/kind_kernel.hpp,
 * src/hyper_kernel.hpp, src/product_mixture.hpp).  Every entry point below
 * names the reference interface it replaces (file:line into /root/reference).
 * The C++ shell in loom_b200/host/ re-hosts those classes on top of this ABI;
 * see INTEGRATION.md for the binding a loom maintainer would add.
 *
 * Plain pointers and sizes only; no torch / C++ types.  All functions return
 * 0 on success, non-zero on failure; LB_NAME(last_error)() gives the message
 * (the C++ shell turns non-zero into LOOM_ERROR-style abort, common.hpp:52-59).
 *
 * The CPU oracle (oracle/loom_oracle.c, test infrastructure only) exports the
 * same functions with the prefix lo_ by including this header with
 * LB_NAME(x) = lo_##x, so parity tests drive both through identical calls.
 */
#ifndef LOOM_B200_H_
#define LOOM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef LB_NAME
#define LB_NAME(x) lb_##x
#endif

/* ------------------------------------------------------------------------- */
/* Feature models (src/models.hpp:116-121).  DD16 and DD256 are one type here:
 * dim <= 256, value stored in one byte. */
enum {
    LB_BB = 0,   /* BetaBernoulli        hp = {alpha, beta}                    */
    LB_DD = 1,   /* DirichletDiscrete    aux[aux_off .. +dim) = alphas         */
    LB_DPD = 2,  /* DirichletProcessDiscrete, dictionary-encoded to [0,dim):
                    hp = {gamma, alpha, beta0}; aux[aux_off .. +dim) = betas   */
    LB_GP = 3,   /* GammaPoisson         hp = {alpha, inv_beta}                */
    LB_NICH = 4  /* NormalInverseChiSq   hp = {mu, kappa, sigmasq, nu}         */
};

typedef struct {
    int32_t type;     /* LB_BB .. LB_NICH */
    int32_t dim;      /* DD: number of categories (<=256); DPD: dictionary size; else 0 */
    float hp[4];
    int32_t aux_off;  /* offset into the aux float array (DD alphas / DPD betas) */
} lb_feature_spec;

/* Features must be sorted by type rank bb < dd < dpd < gp < nich (and dd by
 * ascending dim): that IS loom's global feature-id order
 * (loom/schema.py:32-72, src/product_value.hpp:586-623).
 *
 * Row block (columnar, feature-major), for R rows, F8 = #BB + #DD features,
 * FW = #DPD + #GP + #NICH features:
 *   cat8 : uint8  [F8][R]   value of BB (0/1) and DD (0..dim-1) cells
 *   wide : uint32 [FW][R]   DPD dictionary id / GP count / NICH float bits
 *   mask : uint32 [F][ceil(R/32)]  bit (r & 31) of word r>>5 = cell observed
 * Row position = position in the shuffled stream (src/stream_interval.hpp),
 * so FIFO order of loom's Assignments (src/assignments.hpp:56-74) is row order.
 */

/* Per-kind group storage, identical for upload and download (the order is the
 * field order of ProductModel.Group, src/schema.proto:113-121, features in
 * ascending feature id inside the kind):
 *   counts : uint32 [G]                 clustering counts
 *   ints   : uint32 [n_int_rows][G]     BB: heads, tails | DD/DPD: counts[dim], count_sum
 *                                       GP: count, sum   | NICH: count
 *   flts   : double [n_flt_rows][G]     GP: log_prod     | NICH: mean, count_times_variance
 */
typedef struct {
    int32_t n_features;
    int32_t n_groups;      /* G, including the empty groups */
    int32_t n_int_rows;
    int32_t n_flt_rows;
    uint64_t sample_size;  /* rows currently assigned in this kind */
    float py_alpha, py_d;
} lb_kind_info_t;

/* Hyper-prior grids (src/schema.proto:34-71; defaults loom/hyperprior.py:31-67).
 * Each pointer may be NULL with length 0 = "do not infer this coordinate". */
typedef struct {
    const float *topology;   int32_t n_topology;    /* [n][2] = (alpha, d) */
    const float *clustering; int32_t n_clustering;  /* [n][2] */
    const float *bb_alpha;   int32_t n_bb_alpha;
    const float *bb_beta;    int32_t n_bb_beta;
    const float *dd_alpha;   int32_t n_dd_alpha;
    const float *dpd_gamma;  int32_t n_dpd_gamma;
    const float *dpd_alpha;  int32_t n_dpd_alpha;
    const float *gp_alpha;   int32_t n_gp_alpha;
    const float *gp_inv_beta; int32_t n_gp_inv_beta;
    const float *nich_mu;    int32_t n_nich_mu;
    const float *nich_kappa; int32_t n_nich_kappa;
    const float *nich_sigmasq; int32_t n_nich_sigmasq;
    const float *nich_nu;    int32_t n_nich_nu;
} lb_hyper_prior;

#define LB_UNASSIGNED 0xFFFFFFFFu
/* cat-sweep action encoding: (row_position << 1) | is_remove */
#define LB_ACTION_ADD(r)    (((uint32_t)(r)) << 1)
#define LB_ACTION_REMOVE(r) ((((uint32_t)(r)) << 1) | 1u)

typedef void *lb_handle;

/* Random numbers: every draw is Philox4x32-10 keyed by the caller's 64-bit seed
 * with counter (a_lo, a_hi, b, stream); u = (x0 >> 8) * 2^-24.  Streams:
 *   0 cat sweep     a = action_offset + action index, b = kind id
 *   1 block-PY      a = iteration * F + featureid
 *   2 hyper         a = task id (hyper_kernel.cc:203-207), b = draw index
 *   3 ephemeral kind partition   a = row, b = kind id
 *   4 DPD stick-breaking at ingest (loom_b200_io.h)
 *   5 / 6 query sampler: group pick / value draw (query_sample below)
 *   7 / 8 generate_rows: seating / observed flag and value
 * (replaces distributions::rng_t; sampled values are only statistically
 *  comparable with the reference, see SURVEY.md section 7 "RNG".) */

const char *LB_NAME(last_error)(void);

/* Engine life cycle.  One engine = one CrossCat + Assignments on one device
 * (Loom::Loom, src/loom.cc:43-86). */
int LB_NAME(create)(int device, lb_handle *out);
void LB_NAME(destroy)(lb_handle h);

/* CrossCat::model_load (src/cross_cat.cc:62-120) followed by
 * CrossCat::mixture_init_unobserved (src/cross_cat.cc:50-60): installs the
 * schema, hypers, kind structure and `empty_group_count` empty groups per kind.
 * kind_py is [K][2] = (alpha, d); topology_py is [2].  Kinds may be
 * featureless (the kind kernel's ephemeral kinds, src/kind_kernel.cc:159-189). */
int LB_NAME(set_model)(lb_handle h, int32_t n_features, const lb_feature_spec *specs,
                       const float *aux, int32_t n_aux, int32_t n_kinds,
                       const uint32_t *featureid_to_kindid, const float *kind_py,
                       const float *topology_py, int32_t empty_group_count);

/* Replaces the unzip/parse/split stages (src/cat_pipeline.cc:63-87,
 * ValueSplitter::split src/product_value.hpp:1071-1092): upload one columnar
 * row block from HOST memory.  Clears all assignments. */
int LB_NAME(set_rows)(lb_handle h, uint64_t n_rows, const uint8_t *cat8,
                      const uint32_t *wide, const uint32_t *mask);
/* Stream the SAME block's cells in again from host buffers (same n_rows; the engine must own its rows): the
 * reference re-reads its rows file on every pass over the data (StreamInterval, src/stream_interval.hpp:38-75;
 * Loom::infer_multi_pass, src/loom.cc:217-243).  Assignments and group statistics are untouched. */
int LB_NAME(reload_rows)(lb_handle h, uint64_t n_rows, const uint8_t *cat8,
                         const uint32_t *wide, const uint32_t *mask);

/* Tare-compressed rows: ProductValue::Diff{pos, neg, tares} (src/schema.proto:87-91) as
 * produced by loom's tare + sparsify passes (src/differ.cc:191-298; worked example
 * doc/adapting.md:323-414) with ONE tare row, which is all the automatic pass emits.
 * true row = (row_has_tare ? tare : nothing) + pos - neg.  The block is decoded on the device
 * into the same dense columns set_rows uploads, so score_diff / add_diff / remove_diff
 * (src/product_mixture.cc:241-306, 436-463) reduce to score_value / add_value / remove_value
 * of the reconstructed row -- identical sufficient statistics and scores, no tare caches.
 *   tare_observed [F] (0/1), tare_values [F] raw cell value of the tare row
 *   row_has_tare  [R] (0/1), NULL = every row references the tare
 *   pos_ptr [R+1], pos_feature [nnz], pos_value [nnz] : cells set by the row (CSR by row)
 *   neg_ptr [R+1], neg_feature [nnz]                  : tare cells the row removes */
int LB_NAME(set_rows_diff)(lb_handle h, uint64_t n_rows, const uint8_t *tare_observed,
                           const uint32_t *tare_values, const uint8_t *row_has_tare,
                           const uint64_t *pos_ptr, const uint32_t *pos_feature,
                           const uint32_t *pos_value, const uint64_t *neg_ptr,
                           const uint32_t *neg_feature);
/* ProductMixture::score_diff (src/product_mixture.cc:436-463) for rows [row_begin, row_end) of one kind, computed
 * from the tare-compressed form set_rows_diff left on the device: clustering prior + the kind's tare score (the
 * TareCache entry, product_mixture.hpp:53-57, cc:58-81) + score(pos cells) - score(neg cells).  O(nnz x G) per row
 * instead of O(F x G); same layout and values as score_rows (which scores the decoded row).  Fails unless the loaded
 * block came from set_rows_diff. */
int LB_NAME(score_rows_diff)(lb_handle h, int32_t kind, uint64_t row_begin, uint64_t row_end,
                             float *out_scores);

/* loom's tare / sparsify passes on the LOADED row block -- loom::Differ
 * (src/differ.hpp, src/differ.cc), what `loom_tare` and `loom_sparsify` run before inference.
 *
 * differ_reset     Differ::Differ(schema): zero the per-column summaries and the row count.
 * differ_add_rows  Differ::add_rows (src/differ.cc:93-127): add every row of the loaded block to the
 *                  summaries -- per boolean column the counts of {0,1}, per count column (DD, DPD,
 *                  GP) the counts of the values in [0,16) (CountSummary::max_count,
 *                  src/differ.hpp:63-95); reals are not summarised ("do not sparsify reals").  May
 *                  be called once per window of a long stream.
 * differ_make_tare Differ::_make_tare (src/differ.cc:129-146, 191-206): a column is tared at its
 *                  mode (first maximum) when more than half of ALL rows added hold the mode.
 *                  Output as set_rows_diff takes it: tare_observed [F], tare_values [F].
 * differ_compress_rows  Differ::compress_rows -> _abs_to_rel (src/differ.cc:148-189, 248-298): the
 *                  loaded block relative to the tare.  For a tared column a row that agrees stores
 *                  nothing, one that differs stores pos = its value and neg = the tare cell, one
 *                  that lacks the cell stores neg; untared columns go to pos when observed.  The
 *                  diff stays on the device; *n_pos / *n_neg give its sizes.
 * differ_fetch     copy that diff to HOST arrays in the CSR form of set_rows_diff
 *                  (pos_ptr [R+1], pos_feature / pos_value [n_pos], neg_ptr [R+1], neg_feature [n_neg]).
 * The [0,16) window is over RAW count values as in the reference; DPD cells hold dictionary
 * indices, so the dictionaries are passed the way lbio_model_view holds them (dpd_offsets
 * [n_dpd + 1] into dpd_values, ascending feature id; NULL = identity). */
int LB_NAME(differ_reset)(lb_handle h);
int LB_NAME(differ_add_rows)(lb_handle h, const int32_t *dpd_offsets, const uint32_t *dpd_values);
int LB_NAME(differ_make_tare)(lb_handle h, const int32_t *dpd_offsets, const uint32_t *dpd_values,
                              uint8_t *tare_observed, uint32_t *tare_values);
int LB_NAME(differ_compress_rows)(lb_handle h, const uint8_t *tare_observed, const uint32_t *tare_values,
                                  uint64_t *n_pos, uint64_t *n_neg);
int LB_NAME(differ_fetch)(lb_handle h, uint64_t *pos_ptr, uint32_t *pos_feature, uint32_t *pos_value,
                          uint64_t *neg_ptr, uint32_t *neg_feature);

/* Use `owner`'s row block instead of holding a copy: loom's samples all read the
 * same shuffled rows file (loom/tasks.py:173-194 passes one `rows` path to every
 * seed).  `h` must carry the same schema; its assignments are reset as by
 * set_rows.  The block stays valid until `owner` replaces its rows with a
 * different row count or is destroyed; after the owner re-uploads rows, call
 * share_rows again. */
int LB_NAME(share_rows)(lb_handle h, lb_handle owner);

/* CatKernel::add_row / remove_row over a list of actions, every kind
 * (src/cat_kernel.hpp:177-299; scheduling of src/cat_pipeline.cc:103-125):
 * exact single-site Gibbs, sequential in action order inside each kind, kinds
 * independent.  ADD = model.add_value -> score_value -> sample -> add_value ->
 * push global id; REMOVE = pop -> remove_value.  `kind_mask` selects kinds
 * (NULL = all; used to shard kinds across devices).  Parity-test taps (HOST,
 * may be NULL): debug_picks [n_actions][n_kinds] receives the packed group id
 * chosen by each ADD / vacated by each REMOVE; debug_scores
 * [n_actions][n_kinds][debug_stride] the score vector of every ADD. */
int LB_NAME(cat_sweep)(lb_handle h, const uint32_t *actions, uint64_t n_actions,
                       uint64_t seed, uint64_t action_offset, const uint8_t *kind_mask,
                       uint32_t *debug_picks, float *debug_scores, int32_t debug_stride);

/* The same action list on `n_handles` independent chains at once -- loom's
 * per-sample parallelism (loom/tasks.py:173-194 infer_many -> parallel_map over
 * seeds; each sample is its own CrossCat state and its own rng).  Equivalent to
 * calling cat_sweep(handles[c], actions, n_actions, seeds[c], action_offset,
 * NULL, NULL, NULL, 0) for every c; the device library runs all (chain, kind)
 * pairs in ONE launch, which is how many chains fill a GPU.  All handles must
 * live on the same device. */
int LB_NAME(cat_sweep_multi)(const lb_handle *handles, int32_t n_handles, const uint32_t *actions,
                             uint64_t n_actions, const uint64_t *seeds, uint64_t action_offset);

/* cat_sweep_multi with a row order per chain: chain c runs actions[c * n_actions .. + n_actions).
 * loom shuffles the rows separately for every sample (loom/tasks.py:228-233 infer_one ->
 * loom.runner.shuffle(seed = sample seed)), so samples visit the same rows in different orders;
 * here the chains share ONE row block (share_rows) and each brings its own permutation as row
 * positions in its action list.  The ADD / REMOVE pattern is normally the same for every chain
 * (the schedules depend only on counts); only n_actions must be.  Equivalent to
 * cat_sweep(handles[c], actions + c * n_actions, n_actions, seeds[c], action_offset, ...) for
 * every c, in one launch. */
int LB_NAME(cat_sweep_streams)(const lb_handle *handles, int32_t n_handles, const uint32_t *actions,
                               uint64_t n_actions, const uint64_t *seeds, uint64_t action_offset);

/* Assignments (src/assignments.hpp): packed group id per row position for one
 * kind, LB_UNASSIGNED where the row is not assigned. */
int LB_NAME(get_assignments)(lb_handle h, int32_t kind, uint32_t *out_packed);
int LB_NAME(set_assignments)(lb_handle h, int32_t kind, const uint32_t *packed);

/* ProductMixture load/dump (src/product_mixture.cc:739-856) without the
 * protobuf framing. set_groups appends the empty groups and rebuilds the
 * cached scorers (load_step_2/3_of_3). */
int LB_NAME(kind_info)(lb_handle h, int32_t kind, lb_kind_info_t *out);
int LB_NAME(get_groups)(lb_handle h, int32_t kind, uint32_t *counts, uint32_t *ints,
                        double *flts);
int LB_NAME(set_groups)(lb_handle h, int32_t kind, int32_t n_nonempty_groups,
                        const uint32_t *counts, const uint32_t *ints, const double *flts);

/* ProductMixture_<true>::score_value for a batch of rows of one kind with the
 * tables frozen (src/product_mixture.cc:421-434): out[r - row_begin][g],
 * g < G, row stride = G.  `out` is HOST memory; pass NULL to leave the result
 * in the engine's device scratch (bench use). */
int LB_NAME(score_rows)(lb_handle h, int32_t kind, uint64_t row_begin, uint64_t row_end,
                        float *out);

/* QueryServer::call(Score) for one latent sample (src/query_server.cc:189-231):
 * out[r - row_begin] = sum over kinds of log_sum_exp_g score_value(row r)[g], the
 * log predictive probability of the observed cells of row r under this CrossCat
 * state (kinds in which the row observes nothing contribute 0).  The caller
 * averages over samples: log_sum_exp_l(out_l) - log(#samples).  out: HOST float
 * [row_end - row_begin]. */
int LB_NAME(query_score)(lb_handle h, uint64_t row_begin, uint64_t row_end, float *out);

/* The per-latent half of QueryServer::call(Sample) (src/query_server.cc:100-176 ->
 * ProductMixture_::sample_value src/product_mixture.cc:637-649): conditioned on the loaded row
 * `row`, draw n_samples joint values of the features to_sample[n_to_sample] (feature ids,
 * ascending).  Per kind holding any of them: probs = scores_to_probs(score_value(row)); per
 * sample: group ~ probs, then every requested feature of the kind ~ that group's posterior
 * predictive (Group::sample_value of distributions 2.0.28: BB Bernoulli, DD / DPD categorical
 * with DPD's unseen mass alpha*beta0 reported as LB_DPD_OTHER, GP Gamma-Poisson, NICH Student-t).
 *   out_values [n_samples][n_to_sample]  raw cell values as in the row block (NICH: float bits)
 *   out_groups [n_samples][n_kinds]      parity tap, may be NULL: packed group drawn per kind,
 *                                        LB_UNASSIGNED for kinds none of to_sample lives in
 * Random numbers: stream 5 (a = draw_offset + sample, b = kind) picks the group; stream 6
 * (a = draw_offset + sample, b = featureid << 8 | draw index) draws the value. */
#define LB_DPD_OTHER 0xFFFFFFFFu
int LB_NAME(query_sample)(lb_handle h, uint64_t row, const uint32_t *to_sample, int32_t n_to_sample,
                          uint32_t n_samples, uint64_t seed, uint64_t draw_offset,
                          uint32_t *out_values, uint32_t *out_groups);

/* Loom::generate -> generate_rows (src/loom.cc:548-562, src/generate.hpp:36-93): synthesise n_rows rows from
 * the model.  Per kind the row's group comes from the kind's Pitman-Yor process (sequential seating), every
 * feature is observed with probability `density`, and an observed value is drawn from the group's posterior
 * predictive given the values the group already holds (ProductMixture::sample_value), then added to the group.
 * Afterwards the engine holds the rows (as after set_rows), every row assigned, and the groups' statistics.
 * The seating is sequential per kind (host, float64); the values are independent chains per (group, feature)
 * -- one device thread each, walking its group's rows in order -- which is exactly the reference's order of
 * draws inside a group.  A DPD draw outside the dictionary leaves the cell unobserved.
 * Random numbers: stream 7 (a = row, b = kind) seats the row; stream 8 (a = row, b = featureid << 10 | draw)
 * decides observed (draw 0) and draws the value (draws 1..). */
int LB_NAME(generate_rows)(lb_handle h, uint64_t n_rows, float density, uint64_t seed);
/* The loaded row block as the dense columns set_rows takes (HOST buffers of those sizes). */
int LB_NAME(get_rows)(lb_handle h, uint8_t *cat8, uint32_t *wide, uint32_t *mask);

/* ProductMixture_::add_value for rows [row_begin,row_end) with their CURRENT
 * assignments, all kinds (kind < 0) or one kind: rebuilds/extends sufficient
 * statistics in bulk (src/product_mixture.cc:165-184; bulk form of
 * CrossCat::mixture_load src/cross_cat.cc:148-196).  sign = +1 add, -1 remove.
 * Groups must already exist (packed ids < G); cached scorers are refreshed. */
int LB_NAME(accumulate_rows)(lb_handle h, int32_t kind, uint64_t row_begin,
                             uint64_t row_end, int32_t sign);

/* CrossCat::score_data (src/cross_cat.cc:238-250). */
int LB_NAME(score_data)(lb_handle h, double *out);

/* KindProposer score phase (src/kind_proposer.cc:336-349 ->
 * ProductMixture_::score_feature src/product_mixture.cc:611-620) after the
 * bulk form of KindKernel::add_to_kind_proposer (src/kind_kernel.hpp:193-217):
 * accumulates every assigned row into every kind's all-feature proposer tables
 * and writes out_scores[f][k] = sum_g log marginal, [F][K] row-major (HOST;
 * NULL = keep on device).  The scores are NOT yet passed through
 * scores_to_likelihoods. */
int LB_NAME(kind_proposer_score)(lb_handle h, float *out_scores);

/* Row-sharded form of kind_proposer_score for multi-GPU runs: every
 * device accumulates ITS rows [row_begin,row_end) into zeroed all-feature tables
 * (proposer_accumulate), the caller all-reduces (sum) the flat buffers exposed by
 * proposer_tables -- uint32 counts and float64 sums (sum x, sum x^2, sum log x!), DEVICE
 * pointers for the CUDA library -- and proposer_finish converts sums to moments and
 * scores.  kind_proposer_score == accumulate(0, R) + finish on one device. */
int LB_NAME(proposer_accumulate)(lb_handle h, uint64_t row_begin, uint64_t row_end);
int LB_NAME(proposer_tables)(lb_handle h, int32_t kind, uint32_t **ints, uint64_t *n_ints,
                             double **flts, uint64_t *n_flts);
int LB_NAME(proposer_finish)(lb_handle h, float *out_scores);

/* KindKernel::try_run (src/kind_kernel.cc:118-157) given proposer tables from
 * kind_proposer_score: block Pitman-Yor sampler (src/kind_proposer.cc:269-300)
 * then move_features (src/kind_kernel.cc:83-116).  Writes the new
 * featureid_to_kindid and returns the number of moved features in *n_changed.
 * Featureless kinds are kept (call set_model / add_ephemeral_kinds to manage
 * them, src/kind_kernel.cc:208-227). */
int LB_NAME(kind_run)(lb_handle h, int32_t iterations, uint64_t seed,
                      uint32_t *new_featureid_to_kindid, int32_t *n_changed);

/* KindKernel::init_featureless_kinds (src/kind_kernel.cc:208-227): drop
 * featureless kinds (swap-remove, patching featureid_to_kindid) and append
 * `count` ephemeral kinds, each with a Pitman-Yor prior partition of all
 * assigned rows (src/kind_kernel.cc:159-189). */
int LB_NAME(init_featureless_kinds)(lb_handle h, int32_t count, uint64_t seed);
/* Start drawing the partitions of the NEXT init_featureless_kinds(count, seed) on background host threads.  The
 * seating depends on the seed, the kind ids and WHICH rows are assigned, not on where they sit, so a driver calls this
 * before it sweeps a pass that keeps the assigned set (the reference's KindKernel seats its ephemeral kinds on the
 * critical path, src/kind_kernel.cc:159-189); init_featureless_kinds uses the draw when it was made for exactly its
 * situation and draws afresh otherwise -- the result is the same either way. */
int LB_NAME(prepare_featureless_kinds)(lb_handle h, int32_t count, uint64_t seed);
int LB_NAME(get_kinds)(lb_handle h, int32_t *n_kinds, uint32_t *featureid_to_kindid);

/* HyperKernel::run (src/hyper_kernel.cc:195-235): grid Gibbs over topology,
 * per-kind clustering and per-feature hyper-parameters. */
int LB_NAME(set_hyper_prior)(lb_handle h, const lb_hyper_prior *prior);
int LB_NAME(hyper_run)(lb_handle h, uint64_t seed);
int LB_NAME(get_hypers)(lb_handle h, lb_feature_spec *specs, float *aux, float *kind_py,
                        float *topology_py);
/* The same with the tasks sharded (they are independent: src/hyper_kernel.cc:205-234 runs them as OpenMP
 * tasks with rng_t(seed + taskid)).  Task 0 = topology, 1..K = clustering of kind k-1, 1+K+f = feature f;
 * task_mask [1 + K + F] (0/1) selects the tasks THIS call runs, NULL = all.  Draws are keyed by
 * (seed, task id), so a masked run chooses exactly what the full run chooses for those tasks: ranks that
 * each run a share and then exchange the chosen values (set_hypers) end where hyper_run ends. */
int LB_NAME(hyper_run_tasks)(lb_handle h, uint64_t seed, const uint8_t *task_mask);
/* Install hyper-parameters chosen elsewhere -- the gather step of the sharded hyper kernel.  Same arrays
 * as get_hypers (only hp[] of each spec is read; type / dim / aux_off must match the model); cached
 * scorers are rebuilt as after hyper_run. */
int LB_NAME(set_hypers)(lb_handle h, const lb_feature_spec *specs, const float *aux, const float *kind_py,
                        const float *topology_py);

/* Device timing for bench.py (CUDA events on the engine's stream; the oracle
 * uses a monotonic clock). */
int LB_NAME(sync)(lb_handle h);
int LB_NAME(timer_start)(lb_handle h);
int LB_NAME(timer_stop)(lb_handle h, float *elapsed_ms);
/* Launch / kernel-time counters since the last reset:
 * out[0] = kernels launched, out[1..] = per-kernel-family counts. */
int LB_NAME(launch_count)(lb_handle h, uint64_t *out, int32_t n, int32_t reset);
/* Device time (ms, CUDA events around each launch) accumulated per kernel
 * family since the last reset, same indexing as launch_count. */
int LB_NAME(kernel_ms)(lb_handle h, double *out, int32_t n, int32_t reset);

#ifdef __cplusplus
}
#endif
#endif /* LOOM_B200_H_ */
