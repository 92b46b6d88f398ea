/* loom_b200_comm.h -- multi-GPU entry points of libloomb200.so: the collectives are NCCL calls issued from
 * INSIDE the library on the engine's own stream (no torch, no Python in the data path).
 *
 * The reference has no multi-GPU code (zero NCCL / MPI call sites); what these replace is
 * its OpenMP parallelism over the same independent units:
 *   kind proposer   src/kind_proposer.cc:323-349   (#pragma omp parallel for over kinds / features)
 *   hyper kernel    src/hyper_kernel.cc:203-234    (1 + K + F independent tasks, rng_t(seed + taskid))
 * One process per GPU; every rank holds the same model, the same row block and (after the replicated cat sweep)
 * the same assignments.  A proposal round is then sharded
 *   rows      -> each rank adds rows [R r / N, R (r + 1) / N) into zeroed all-feature tables (K4),
 *   exchange  -> ONE ncclAllReduce(sum) over the uint32 slab and one over the float64 slab,
 *   features  -> each rank scores a contiguous block of ceil(F / N) features (K5),
 *   exchange  -> ONE ncclAllGather of the [F][K] score matrix,
 * after which every rank runs the identical block Pitman-Yor sampler (lb_kind_run, same seed).
 * Status / error convention as in loom_b200.h. */
#ifndef LOOM_B200_COMM_H_
#define LOOM_B200_COMM_H_

#include "loom_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#define LB_COMM_ID_BYTES 128

/* ncclGetUniqueId: called by rank 0; the bytes travel to the other ranks by any host channel (bench.py: a file). */
int lb_comm_unique_id(uint8_t *id_out);
/* ncclCommInitRank on the engine's device; one communicator per engine. */
int lb_comm_init(lb_handle h, const uint8_t *id, int32_t rank, int32_t world);
int lb_comm_destroy(lb_handle h);
int lb_comm_rank(lb_handle h, int32_t *rank, int32_t *world);
/* all ranks' streams have reached this point (an all-reduce of one word + stream sync) */
int lb_comm_barrier(lb_handle h);
/* inout[i] = max over ranks (multi-GPU timings are the max over ranks) */
int lb_comm_max_f64(lb_handle h, double *inout, int32_t n);

/* stats of one sharded proposal round, all in ms / bytes, measured with CUDA events on the engine's stream */
typedef struct {
    double accumulate_ms;      /* K4 over this rank's rows */
    double allreduce_ms;       /* both slabs */
    double allreduce_bytes;    /* 4 * ints + 8 * floats */
    double score_ms;           /* moments + K5 over this rank's features */
    double allgather_ms;
    double allgather_bytes;
    double row_begin, row_end, feature_begin, feature_end;
} lb_comm_round_stats;

/* KindProposer::infer_assignments tare + score phases (src/kind_proposer.cc:323-349) sharded as described above;
 * out_scores [F][K] on the HOST (NULL = keep on device), identical on every rank.  With a world of 1 this is
 * lb_kind_proposer_score. */
int lb_comm_kind_proposer_score(lb_handle h, float *out_scores, lb_comm_round_stats *stats);

/* HyperKernel::run with task t run by rank t % world (src/hyper_kernel.cc:203-234), then one all-reduce of the
 * chosen hyper-parameters (owners contribute their choice, everyone else zeros) and lb_set_hypers on every rank:
 * ends where lb_hyper_run ends, bit for bit. */
int lb_comm_hyper_run(lb_handle h, uint64_t seed, double *allreduce_ms);

/* Kind-sharded cat sweep (doc/adapting.md:144-147: the kinds of a row are independent tasks -- the reference runs one
 * thread per kind, src/cat_pipeline.cc:103-125): rank owner[k] has swept kind k with lb_cat_sweep(kind_mask); this
 * publishes every kind's groups, id maps and assignment column from its owner to all ranks (one all-reduce of the
 * headers + one NCCL group of broadcasts), after which every rank holds the state an unsharded sweep leaves.
 * bytes_out / ms_out (optional): payload bytes and device time of the broadcast group. */
int lb_comm_sync_kinds(lb_handle h, const int32_t *owner, double *bytes_out, double *ms_out);

/* Row-sharded bulk accumulate (ProductMixture::add_value over every row with its given assignment,
 * src/product_mixture.cc:165-184, as lb_accumulate_rows does on one device): every rank holds the same groups and the
 * assignments of all rows, zeroes the statistics, adds its share of the rows; counts and integer statistics are summed as
 * uint32, float statistics as float64 sums (moments are formed after the reduce); one NCCL group of all-reduces.  Every
 * rank ends with the statistics of all rows. */
int lb_comm_accumulate_rows(lb_handle h, double *allreduce_ms, double *allreduce_bytes);

#ifdef __cplusplus
}
#endif
#endif
