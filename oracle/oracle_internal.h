/* oracle_internal.h -- shared structs of the CPU oracle (test infrastructure). */
#ifndef ORACLE_INTERNAL_H_
#define ORACLE_INTERNAL_H_

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "loom_oracle.h"

typedef struct {
    int nf;            /* features in this kind */
    int *fids;         /* global feature ids, ascending */
    double alpha, d;   /* Pitman-Yor hypers of the row clustering */
    int G, Gcap;       /* packed groups incl. empty ones; column capacity */
    uint32_t *counts;  /* [Gcap] */
    uint64_t sample_size;
    uint32_t *p2g;     /* packed -> global (MixtureIdTracker) */
    uint32_t *g2p;     /* global -> packed */
    uint32_t g2p_cap, next_global;
    int n_int_rows, n_flt_rows, n_cache_rows;
    int *int_off, *flt_off, *cache_off; /* [nf] first row of each feature */
    uint32_t *ints;    /* [n_int_rows][Gcap] */
    double *flts;      /* [n_flt_rows][Gcap] */
    double *cache;     /* [n_cache_rows][Gcap] cached scorers (FastMixture) */
    double *shifted;   /* [Gcap] cached clustering scores */
    uint32_t *assign;  /* [R] global group id per row position, LB_UNASSIGNED */
} kind_t;

typedef struct {
    int F, F8, FW;
    lb_feature_spec *specs;
    float *aux;
    int n_aux;
    int K;
    kind_t *kinds;
    uint32_t *f2k;
    double topo_alpha, topo_d;
    int empty_group_count;
    uint64_t R, W; /* rows, mask words per feature */
    int rows_borrowed; /* cat8 / wide / mask belong to another engine (lo_share_rows) */
    uint8_t *cat8;
    uint32_t *wide;
    uint32_t *mask;
    /* Differ summaries (src/differ.hpp:49-95): [F][16] counts of the observed raw values in [0,16) */
    uint64_t *differ_hist;
    uint64_t differ_rows;
    /* the last compress_rows result */
    uint64_t *diff_pos_ptr, *diff_neg_ptr, diff_rows;
    uint32_t *diff_pos_feature, *diff_pos_value, *diff_neg_feature;
    /* hyper prior (owned copies) */
    lb_hyper_prior hp;
    float *hp_store;
    /* kind proposer tables: per kind, all F features */
    int prop_K;
    int *prop_int_off, *prop_flt_off; /* [F] first row of feature f in the all-feature layout */
    int prop_int_rows, prop_flt_rows;
    uint32_t **prop_ints; /* [K] -> [prop_int_rows][G_k] */
    double **prop_flts;   /* [K] -> [prop_flt_rows][G_k] */
    int *prop_G;
    float *prop_scores;   /* [F][K] */
    /* timer */
    double t0;
} engine_t;

int lo_fail(const char *fmt, ...);

static inline int spec_int_rows(const lb_feature_spec *s) {
    switch (s->type) {
    case LB_BB: return 2;
    case LB_DD: case LB_DPD: return s->dim + 1;
    case LB_GP: return 2;
    case LB_NICH: return 1;
    }
    return 0;
}
static inline int spec_flt_rows(const lb_feature_spec *s) {
    return s->type == LB_GP ? 1 : (s->type == LB_NICH ? 2 : 0);
}
static inline int spec_cache_rows(const lb_feature_spec *s) {
    switch (s->type) {
    case LB_BB: return 2;
    case LB_DD: case LB_DPD: return s->dim + 1;
    case LB_GP: return 2;
    case LB_NICH: return 4;
    }
    return 0;
}

static inline int cell_observed(const engine_t *e, int f, uint64_t r) {
    return (e->mask[(uint64_t)f * e->W + (r >> 5)] >> (r & 31)) & 1u;
}
static inline uint32_t cell_raw(const engine_t *e, int f, uint64_t r) {
    return f < e->F8 ? (uint32_t)e->cat8[(uint64_t)f * e->R + r]
                     : e->wide[(uint64_t)(f - e->F8) * e->R + r];
}

/* closed forms (oracle_math.c) */
double om_score_value(const lb_feature_spec *s, const float *aux, const uint32_t *ints,
                      const double *flts, int stride, uint32_t raw);
double om_score_data(const lb_feature_spec *s, const float *aux, const uint32_t *ints,
                     const double *flts, int stride);
void om_add_value(const lb_feature_spec *s, uint32_t *ints, double *flts, int stride,
                  uint32_t raw, int sign);
double om_py_score_counts(double alpha, double d, const uint32_t *counts, int n);
int om_sample_from_scores(double u, double *scores, int n);
int om_sample_from_likelihoods(double u, const double *p, int n, double total);
uint32_t om_sample_value(const lb_feature_spec *s, const float *aux, const uint32_t *ints, const double *flts,
                         int stride, uint64_t seed, uint64_t a, uint32_t featureid, uint32_t stream, uint32_t first_draw);

/* kind helpers (loom_oracle.c) */
void kind_free(kind_t *k);
void kind_init(engine_t *e, kind_t *k, const int *fids, int nf, double alpha, double d,
               const uint32_t *counts, int G);
void kind_refresh_group(engine_t *e, kind_t *k, int g);
void kind_refresh_all(engine_t *e, kind_t *k);
void kind_refresh_clustering(kind_t *k);
void kind_ensure_cap(kind_t *k, int G);

#endif
