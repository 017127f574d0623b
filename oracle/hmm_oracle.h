/*
 * oracle/hmm_oracle.h -- TEST INFRASTRUCTURE ONLY (see hmm_oracle.c).
 * CPU restatement of the reference E-step / M-step, and (ref_*) the same entry points served by the
 * UNMODIFIED reference objects when oracle/_ref/libref_harness.so has been built (oracle/ref_harness.c).
 */
#ifndef HFG_ORACLE_H
#define HFG_ORACLE_H
#include "../include/hfg.h"
#ifdef __cplusplus
extern "C" {
#endif

/* The negative-binomial model (reference `--modelType negative_binomial`, MODEL_NEGATIVE_BINOMIAL) is restated by the
 * oracle ahead of the product: cfg->model_type == ORC_MODEL_NEGATIVE_BINOMIAL selects it in every orc_ and ref_ entry point.
 * hfg_region_params.mean[s][c] then holds theta, .var[s][c] lambda, .weight[s][c] the mixture weights (NegativeBinomial,
 * hmm_utils.h); hfg_region_stats.mean_* hold the theta estimator, .var_* the lambda estimator, .weight_* the weights'. */
#define ORC_MODEL_NEGATIVE_BINOMIAL 2
long double orc_digammal(long double x);
void orc_nb_histogram_out(double *out); /* test hook, see hmm_oracle.c */

double orc_beta(const hfg_config *cfg, const hfg_chunk_desc *ch, int i);

int orc_estep(const hfg_config *cfg, int n_chunks, const hfg_chunk_desc *chunks, const uint16_t *cov,
              const uint16_t *mapq, const uint16_t *clip, const uint8_t *region, const double *alpha,
              const hfg_region_params *params, hfg_region_stats *stats, double *loglik, double *chunk_logliks,
              int8_t *labels, double *posteriors, double *fwd, double *bwd, double *scales_out, int forward_only);

int orc_best_num_collapsed_comps(int max_coverage, const int32_t *region_coverages, int n_regions);

int orc_model_init(const hfg_config *cfg, const int32_t *region_coverages, int window_len, int start_only_mode,
                   hfg_region_params *params);

int orc_mstep(const hfg_config *cfg, hfg_region_params *params, const hfg_region_stats *stats, double tol,
              int *converged_out);

int orc_run_em(const hfg_config *cfg, int n_chunks, const hfg_chunk_desc *chunks, const uint16_t *cov,
               const uint16_t *mapq, const uint16_t *clip, const uint8_t *region, const double *alpha,
               hfg_region_params *params, int max_iterations, double tol, double *logliks, int *n_esteps,
               int8_t *labels);

int orc_feasible(const hfg_config *cfg, const hfg_region_params *params);
int orc_squarem(const hfg_config *cfg, const hfg_region_params *p0, const hfg_region_params *p1,
                const hfg_region_params *p2, int n_shrinks, double margin, hfg_region_params *prime, double *alpha_rate);
int orc_run_em_accelerated(const hfg_config *cfg, int n_chunks, const hfg_chunk_desc *chunks, const uint16_t *cov,
                           const uint16_t *mapq, const uint16_t *clip, const uint8_t *region, const double *alpha,
                           hfg_region_params *params, int max_iterations, double tol, double *logliks,
                           double *alpha_rates, int *n_outer, int8_t *labels);

#ifdef __cplusplus
}
#endif
#endif
