/*
 * hfg.h -- C-ABI of libhfg (HMM-Flagger on B200): the drop-in boundary.
 *
 * This library replaces ONE path of mobinasri/flagger v1.2.0: the per-chunk HMM
 * E-step of the `hmm_flagger` binary (scaled forward/backward, pair statistics,
 * posterior-argmax decode), i.e. the two functions
 *
 *     void EM_runOneIterationForList(stList *emList, HMM *model, int threads);  // hmm.h:109, hmm.c:739-780
 *     void EM_runForwardForList   (stList *emList, HMM *model, int threads);   // hmm.h:113, hmm.c:790-816
 *
 * plus thin host mirrors of the O(#parameters) functions that sit on either side of
 * that seam (model construction and the M-step) so that a complete EM run can be
 * driven through this header alone.  Everything is plain C: opaque handle, plain
 * pointers and sizes, caller-owned host buffers, int status codes.  There is no CPU
 * fallback: every compute entry point fails with HFG_ERR_CUDA when no sm_100 device /
 * CUDA runtime is usable.
 *
 * Citations are file:line under the reference tree (programs/...).
 */
#ifndef HFG_H
#define HFG_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HFG_NUM_STATES 4   /* Err, Dup, Hap, Col  (submodules/hmm_utils/hmm_utils.h:19-25; MSJ excluded, src/hmm_flagger.c:235) */
#define HFG_MAX_COMPS 16   /* mixture components per state; the CLI clamps the auto value to [2,10] (src/hmm_flagger.c:1012-1013) */
#define HFG_MAX_REGIONS 64 /* region index lives in 6 bits of the annotation flag (submodules/ptBlock/ptBlock.c:294-304) */

enum hfg_state { HFG_STATE_ERR = 0, HFG_STATE_DUP = 1, HFG_STATE_HAP = 2, HFG_STATE_COL = 3 };

/* submodules/hmm_utils/hmm_utils.h:43-48.  HFG_MODEL_NEGATIVE_BINOMIAL is served by the blocking E-steps on one GPU
 * (hfg_em_iteration, hfg_forward_only, hfg_get_posteriors, hfg_run_em through the host loop): the emission table is built on
 * the host, the device returns per-tile pair masses, the host folds them into the histogram of hmm.c:615-617 and runs the
 * estimator update (hfg_nb_*).  The device-resident loop (hfg_em_*) and the multi-GPU exchange do not serve it. */
enum hfg_model_type { HFG_MODEL_TRUNC_EXP_GAUSSIAN = 0, HFG_MODEL_GAUSSIAN = 1, HFG_MODEL_NEGATIVE_BINOMIAL = 2 };

enum hfg_status {
    HFG_OK = 0,
    HFG_ERR_INVALID = 1,         /* bad argument / call order */
    HFG_ERR_CUDA = 2,            /* CUDA runtime error, or no usable device (no CPU fallback) */
    HFG_ERR_SCALE_UNDERFLOW = 3, /* a forward scale fell below 1e-50: the reference exits here (hmm.c:412-415,521-524) */
    HFG_ERR_NAN = 4,             /* an emission pdf evaluated to NaN: the reference exits here (hmm_utils.c:782-786) */
    HFG_ERR_NOMEM = 5
};

typedef struct hfg_ctx hfg_ctx;

/* Run-constant configuration: what EM_construct / HMM_construct / EM_setMinReadFractionAtEnds fix
 * for the lifetime of a run (hmm.c:22-77,253-278; src/hmm_flagger.c:164-237). */
typedef struct hfg_config {
    int32_t model_type;                 /* enum hfg_model_type */
    int32_t n_regions;                  /* 1..HFG_MAX_REGIONS */
    int32_t n_comps[HFG_NUM_STATES];    /* {1,1,1,K} for the shipped models */
    int32_t adjust_contig_ends;         /* 0 = --disableAdjustContigEnds (beta == 1) */
    int32_t mean_read_length;           /* header #avg_alignment_len (em->meanReadLength) */
    double min_read_fraction_at_ends;   /* --minReadFractionAtEnds */
    double max_high_mapq_ratio;         /* --maxHighMapqRatio, Dup invalid above it (default 0.25) */
    double min_high_mapq_ratio;         /* --minHighMapqRatio, Col invalid below it (default 0.75) */
    double min_highly_clipped_ratio;    /* END column valid at or above it (fixed 1.0, src/hmm_flagger.c:222) */
    int32_t device;                     /* CUDA device ordinal */
    int32_t reserved;
} hfg_config;

/* One chain (the reference's Chunk, submodules/chunk/chunk.h:11-34). `offset` is the index of the
 * chunk's first window in the flat observation arrays handed to hfg_set_chunks. */
typedef struct hfg_chunk_desc {
    int32_t ctg_len;
    int32_t s;          /* 0-based inclusive start on the contig */
    int32_t e;          /* 0-based inclusive end */
    int32_t window_len;
    int32_t n_windows;  /* coverageInfoSeqLen */
    int32_t reserved;
    int64_t offset;
} hfg_chunk_desc;

/* Parameters of one region (EmissionDistSeries + Transition of that region; hmm.h:14-25).
 * Row index of mean/var/weight is the state; row HFG_STATE_ERR is used only by HFG_MODEL_GAUSSIAN and
 * HFG_MODEL_NEGATIVE_BINOMIAL.  For the negative binomial, mean[s][c] holds theta and var[s][c] lambda of the
 * (theta, lambda) parametrisation (NegativeBinomial, hmm_utils.h), and hfg_region_stats.mean_x / var_x their estimators.
 * trans is the 5x5 matrix with row 4 = start and column 4 = end (hmm_utils.c:2109-2128). */
typedef struct hfg_region_params {
    double lambda;      /* TruncExponential.lambda     (hmm_utils.h:393-397) */
    double trunc_point; /* TruncExponential.truncPoint */
    double mean[HFG_NUM_STATES][HFG_MAX_COMPS];
    double var[HFG_NUM_STATES][HFG_MAX_COMPS];
    double weight[HFG_NUM_STATES][HFG_MAX_COMPS];
    double trans[HFG_NUM_STATES + 1][HFG_NUM_STATES + 1];
} hfg_region_params;

/* E-step sufficient statistics of one region, laid out as the reference's estimator arrays
 * (ParameterEstimator.{numeratorPerComp,denominatorPerComp}, hmm_utils.h:~90; TransitionCountData.countMatrix,
 * hmm_utils.h:768-778).  hfg_em_iteration OVERWRITES these with this call's sums; the adapter adds them
 * into the model's estimators (EM_updateModelEstimators, hmm.c:548-560). */
typedef struct hfg_region_stats {
    double trans_count[HFG_NUM_STATES][HFG_NUM_STATES]; /* [pre][state] */
    double lambda_num, lambda_den;
    double mean_num[HFG_NUM_STATES][HFG_MAX_COMPS], mean_den[HFG_NUM_STATES][HFG_MAX_COMPS];
    double var_num[HFG_NUM_STATES][HFG_MAX_COMPS], var_den[HFG_NUM_STATES][HFG_MAX_COMPS];
    double weight_num[HFG_NUM_STATES][HFG_MAX_COMPS], weight_den[HFG_NUM_STATES][HFG_MAX_COMPS];
} hfg_region_stats;

/* ---- lifetime ---------------------------------------------------------------------------------- */

/* Creates a context bound to cfg->device.  Replaces the per-chunk EM_construct loop (src/hmm_flagger.c:320-330). */
int hfg_create(hfg_ctx **ctx, const hfg_config *cfg);
void hfg_destroy(hfg_ctx *ctx);
/* Message of the last failing call on this context ("" if none); ctx may be NULL for hfg_create failures. */
const char *hfg_last_error(const hfg_ctx *ctx);

/* Uploads all chains once.  cov / cov_high_mapq / cov_high_clip are the u16 window values of CoverageInfo
 * (ptBlock.h:79-92; already averaged, rounded and clipped to 250 by Chunk_addWindow, chunk.c:393-441) and
 * region is CoverageInfo_getRegionIndex per window; each array has sum(n_windows) entries, chunk after chunk
 * in list order.  Replaces the CoverageInfo** walk of hmm.c:344-345,382-385. */
int hfg_set_chunks(hfg_ctx *ctx, int32_t n_chunks, const hfg_chunk_desc *chunks,
                   const uint16_t *cov, const uint16_t *cov_high_mapq, const uint16_t *cov_high_clip,
                   const uint8_t *region);

int64_t hfg_num_windows(const hfg_ctx *ctx);

/* Page-locked host memory (cudaMallocHost / cudaFreeHost behind a plain-C face).  A `labels` buffer obtained here is
 * written by the device directly; any other host pointer works too and costs one staging copy per call. */
void *hfg_host_alloc(size_t bytes);
void hfg_host_free(void *p);

/* ---- the hot path ------------------------------------------------------------------------------ */

/* One E-step over all chunks == EM_runOneIterationForList (hmm.c:739-780):
 *   forward + backward + pair statistics + per-window posterior-argmax label.
 * alpha: 4x4 row-major [pre][state] (model->alpha).  params/stats: n_regions entries.
 * *loglik = sum over chunks of sum_i log(scale_i)  (model->loglikelihood).
 * labels: sum(n_windows) int8 (Inference.prediction, hmm.c:730-736) or NULL to skip the copy.
 * All pointers are HOST memory; the call blocks until the results are in them. */
int hfg_em_iteration(hfg_ctx *ctx, const double *alpha, const hfg_region_params *params,
                     hfg_region_stats *stats, double *loglik, int8_t *labels);

/* Forward pass only == EM_runForwardForList (hmm.c:790-816), used by SQUAREM (hmm.c:900,912). */
int hfg_forward_only(hfg_ctx *ctx, const double *alpha, const hfg_region_params *params, double *loglik);

/* Normalised posteriors (EM_getPosterior, hmm.c:671-685) of the LAST hfg_em_iteration, W x 4 row-major,
 * for --writePosteriorProbs (src/hmm_flagger.c:240-282). */
int hfg_get_posteriors(hfg_ctx *ctx, double *posteriors);

/* Per-chunk forward log-likelihoods (em->loglikelihood, hmm.c:428) of the last E-step / forward pass. */
int hfg_get_chunk_logliks(hfg_ctx *ctx, double *logliks);

/* Device-resident variant for multi-GPU drivers: enqueues the E-step on `stream` (a cudaStream_t, may be NULL
 * for the context's own stream) and leaves  [n_regions x hfg_region_stats | loglik]  (doubles) in the DEVICE
 * buffer stats_dev, ready for a sum all-reduce; labels stay on the device.  Does not synchronise. */
int hfg_em_iteration_device(hfg_ctx *ctx, const double *alpha, const hfg_region_params *params,
                            void *stats_dev, void *stream);
size_t hfg_stats_device_bytes(const hfg_ctx *ctx);
/* Copies the labels of the last E-step to the host (blocking). */
int hfg_get_labels(hfg_ctx *ctx, int8_t *labels);

/* Multi-GPU, one process per GPU: after this set-up every hfg_em_iteration / hfg_forward_only call returns statistics
 * and log-likelihood already SUMMED over all ranks -- the E-step kernel ends with a sum all-reduce of its result block
 * through peer memory (NVLink P2P stores into the peers' mailboxes, counters with system scope), in rank order, so all
 * ranks hold identical bits.  Labels / posteriors stay per-rank (each rank's own chunks).  Protocol: every rank calls
 * hfg_peer_export, the hfg_peer_handle_bytes()-sized handles are all-gathered by the host program (torch.distributed,
 * MPI, ...), every rank calls hfg_peer_connect with all of them in rank order.  All ranks must then make the same
 * sequence of E-step calls.  A rank that never arrives turns into HFG_ERR_CUDA after ~2 s, not a hang. */
size_t hfg_peer_handle_bytes(void);
int hfg_peer_export(hfg_ctx *ctx, void *handle_out);
int hfg_peer_connect(hfg_ctx *ctx, int n_ranks, int rank, const void *handles);
/* Device-side rendezvous of all ranks, queued on the context's stream (no host synchronisation): what a benchmark puts in
 * front of a timed iteration so that every rank's interval starts together.  No-op for a single rank. */
int hfg_peer_barrier(hfg_ctx *ctx);

/* ---- host mirrors of the O(#parameters) neighbours of the seam ---------------------------------- */

/* max(2, min(10, maxCov / min(regionCov) + 1))  (src/hmm_flagger.c:105-111,1008-1013). */
int hfg_best_num_collapsed_comps(int max_coverage, const int32_t *region_coverages, int n_regions);

/* Initial model == createModel + HMM_construct (src/hmm_flagger.c:164-237, hmm.c:22-77,
 * hmm_utils.c:1605-1652,2109-2128) with --initialRandomDev 0.  start_only_mode scales the baseline by
 * window_len / avg_alignment_len (src/hmm_flagger.c:190-195). */
int hfg_model_init(const hfg_config *cfg, const int32_t *region_coverages, int window_len,
                   int start_only_mode, hfg_region_params *params);

/* M-step == HMM_estimateParameters (hmm.c:120-127; hmm_utils.c:1791-1903,2185-2219).
 * Updates params in place from stats; *converged is the reference's return value. */
int hfg_mstep(const hfg_config *cfg, hfg_region_params *params, const hfg_region_stats *stats,
              double convergence_tol, int *converged);

/* Whole EM loop of runHMMFlagger without its file outputs (src/hmm_flagger.c:337-467): up to max_iterations x
 * (E-step, M-step) while not converged, then the final-inference E-step.  logliks receives one value per E-step
 * (capacity max_iterations + 1); *n_esteps the number run.  labels (nullable) are those of the final E-step. */
int hfg_run_em(hfg_ctx *ctx, const double *alpha, hfg_region_params *params, int max_iterations,
               double convergence_tol, double *logliks, int *n_esteps, int8_t *labels);

/* Device-resident EM loop: what hfg_run_em is made of.  The parameters stay in device memory between iterations and
 * the M-step (the same code as hfg_mstep, compiled for the device) runs in the tail of the E-step kernel, so successive
 * iterations are back-to-back kernel launches with no host round trip (the reference's loop, src/hmm_flagger.c:337-431,
 * pays a host M-step and a thread-pool start per iteration).
 *   hfg_em_begin    uploads alpha / parameters, clears the loop state (max_esteps = log-likelihood slots to keep);
 *   hfg_em_enqueue  queues one iteration on the context's stream WITHOUT waiting: E-step + M-step, or with final_pass != 0
 *                   the final inference E-step (no M-step).  Once an iteration has converged (the reference's test,
 *                   convergence_tol) or hit a fatal condition, later non-final iterations return at once on the device;
 *   hfg_em_finish   waits, then returns the parameters, the log-likelihood of every E-step that ran, their number,
 *                   whether the loop converged, and (nullable) the labels of the last E-step.
 * In a multi-GPU job (hfg_peer_connect) every rank queues the same sequence; the statistics are all-reduced inside the
 * kernel and every rank's M-step sees identical bits. */
int hfg_em_begin(hfg_ctx *ctx, const double *alpha, const hfg_region_params *params, double convergence_tol, int max_esteps);
int hfg_em_enqueue(hfg_ctx *ctx, int final_pass);
int hfg_em_finish(hfg_ctx *ctx, hfg_region_params *params, double *logliks, int *n_esteps, int *converged, int8_t *labels);
/* Device time (ms) of the i-th hfg_em_enqueue since hfg_em_begin (CUDA events on the context's stream around the
 * kernel); valid after hfg_em_finish. */
double hfg_em_enqueued_ms(hfg_ctx *ctx, int i);

/* Batched EM runs: many (alpha, start parameters) candidates over the SAME windows, several at a time on one GPU.  The alpha
 * tuner of the reference (programs/src/tune_alpha_hmm_flagger.py:243-267) starts one hmm_flagger process per candidate and
 * input; here a batch holds `n_lanes` contexts over the same chunks, each confined to num_SMs / n_lanes CTAs with its own
 * stream, so that n_lanes device-resident EM loops (hfg_em_*) run side by side and the fixed per-iteration cost of a small
 * input (grid barriers, scans, the M-step tail: ~55 us however few the windows) is paid by all of them at once.
 *   hfg_batch_create      n_lanes in 1..16;
 *   hfg_batch_set_chunks  the arguments of hfg_set_chunks, given to every lane;
 *   hfg_batch_run_em      n_runs candidates: alphas [n_runs][16], params [n_runs][n_regions] (in: start, out: fitted),
 *                         logliks [n_runs][max_iterations + 1], n_esteps [n_runs], labels [n_runs][n_windows] (nullable);
 *                         each run is exactly hfg_run_em (same loop, same stopping rule) on a smaller grid: labels identical,
 *                         log-likelihoods equal to rounding (the segment sums associate differently).
 * Not for HFG_MODEL_NEGATIVE_BINOMIAL (no alpha to tune). */
typedef struct hfg_batch hfg_batch;
int hfg_batch_create(hfg_batch **out, const hfg_config *cfg, int n_lanes);
int hfg_batch_set_chunks(hfg_batch *b, int32_t n_chunks, const hfg_chunk_desc *chunks, const uint16_t *cov,
                         const uint16_t *cov_high_mapq, const uint16_t *cov_high_clip, const uint8_t *region);
int hfg_batch_run_em(hfg_batch *b, int n_runs, const double *alphas, hfg_region_params *params, int max_iterations,
                     double convergence_tol, double *logliks, int *n_esteps, int8_t *labels);
const char *hfg_batch_last_error(const hfg_batch *b);
void hfg_batch_destroy(hfg_batch *b);
/* Upper bound on the CTAs (one per SM) the next hfg_set_chunks of this context may use; 0 = all SMs. */
int hfg_set_max_blocks(hfg_ctx *ctx, int max_blocks);

/* --accelerate (SQUAREM; SquareAccelerator, hmm.c:820-1098): feasibility of a parameter set (HMM_isFeasible, hmm.c:80-87),
 * the step length from three successive parameter sets (SquareAccelerator_computeRates, hmm.c:1000-1098), the
 * extrapolated + renormalised parameters for a step length (SquareAccelerator_computeValuesForModelPrime, hmm.c:921-997)
 * and the step-halving rule (SquareAccelerator_shrinkAlphaAndRecomputeModelPrime, hmm.c:869-883). */
int hfg_params_feasible(const hfg_config *cfg, const hfg_region_params *params);
double hfg_squarem_alpha_rate(const hfg_config *cfg, const hfg_region_params *p0, const hfg_region_params *p1,
                              const hfg_region_params *p2);
int hfg_squarem_prime(const hfg_config *cfg, const hfg_region_params *p0, const hfg_region_params *p1,
                      const hfg_region_params *p2, double alpha_rate, hfg_region_params *prime);
int hfg_squarem_shrink(const hfg_config *cfg, const hfg_region_params *p0, const hfg_region_params *p1,
                       const hfg_region_params *p2, double margin, double *alpha_rate, hfg_region_params *prime);

/* One outer iteration of the `acceleration` branch (src/hmm_flagger.c:344-416) without its closing M-step:
 * E(p0) -> M -> E(p1) -> M -> p' (SquareAccelerator_getModelPrime, hmm.c:885-918) -> E(p').  In: params = p0.
 * Out: params = p', stats = the statistics of E(p'), *loglik0 = log-likelihood of p0, *alpha_rate (may be NULL). */
int hfg_squarem_iteration(hfg_ctx *ctx, const double *alpha, hfg_region_params *params, hfg_region_stats *stats,
                          double convergence_tol, double *loglik0, double *alpha_rate);

/* The accelerated EM loop of runHMMFlagger (src/hmm_flagger.c:337-467 with the `acceleration` branch :382-416): per outer
 * iteration E(p0) -> M -> E(p1) -> M -> extrapolate p' (forward-only passes choose the step, hmm.c:885-918) -> E(p') -> M.
 * logliks[k] = log-likelihood of p0 at outer iteration k (what loglikelihood.tsv holds), alpha_rates[k] (may be NULL)
 * the accepted step; both need max_iterations + 1 slots; the last loglik is the final inference pass. */
int hfg_run_em_accelerated(hfg_ctx *ctx, const double *alpha, hfg_region_params *params, int max_iterations,
                           double convergence_tol, double *logliks, double *alpha_rates, int *n_outer, int8_t *labels);

/* ---- negative-binomial model: host side (hmm_utils.c:320-640); the device side is opt-in, see enum hfg_model_type ---- */
#define HFG_NB_TABLE_X 251 /* coverage values 0..MAX_COVERAGE_VALUE (hmm_utils.h:15) */
#define HFG_NB_BINS 250    /* bins of the per-state count histogram; x = 250 falls into bin 249 (count_data.c:56-64) */
/* pmf of every (region, state, x), summed over the weighted components and floored at 1e-40 per component
 * (NegativeBinomial_getProb / _getComponentProbs, hmm_utils.c:479-515): table[(r * 4 + s) * HFG_NB_TABLE_X + x].
 * HFG_ERR_NAN where the reference would exit ("prob is NAN"). */
int hfg_nb_emission_table(const hfg_config *cfg, const hfg_region_params *params, double *table);
/* The theta / lambda / weight estimator sums (stats[r].mean_x / var_x / weight_x, OVERWRITTEN; the other fields are left
 * alone) from histogram[(r * 4 + s) * HFG_NB_BINS + x] = pair mass of state s at coverage x in region r, i.e. what
 * EmissionDistSeries_updateAllEstimatorsUsingCountData (hmm_utils.c:1662-1672) makes of one chunk's CountData through
 * NegativeBinomial_updateEstimator (:536-563) with the digamma table of :392-406. */
int hfg_nb_stats_from_histogram(const hfg_config *cfg, const hfg_region_params *params, const double *histogram,
                                hfg_region_stats *stats);
/* digamma in long double (the routine behind the table above; submodules/digamma/digamma.c:36-116 restated) */
long double hfg_digammal(long double x);

/* A destroyed context leaves its device arena (one per process) for the next context on the same device, so that a
 * process running job after job does not pay cudaMalloc per job; this returns it to the driver. */
void hfg_release_cached_memory(void);

/* Creates the CUDA context of `device` (about a second on a B200 box; the reference has no counterpart) so that a host
 * program can do it on a thread of its own while it reads its input; hfg_create then finds the context in place.
 * Returns HFG_OK or HFG_ERR_CUDA.  Optional: hfg_create does the same on first use. */
int hfg_device_warmup(int device);

/* ---- instrumentation ---------------------------------------------------------------------------- */

/* Number of kernels this context has launched so far. */
int64_t hfg_kernel_launches(const hfg_ctx *ctx);
/* Device time (ms, CUDA events on the launching stream) of the E-step kernel in the last hfg_em_iteration*. */
double hfg_last_estep_kernel_ms(hfg_ctx *ctx);
/* Device time (ms, CUDA events on the context's stream) of the whole last blocking call: parameter upload, kernel,
 * read-back of statistics (and labels). */
double hfg_last_call_device_ms(hfg_ctx *ctx);

/* on != 0: the blocking calls record the CUDA events the two functions above read.  Off by default: a single-region model
 * then takes the one-launch path (parameters in the kernel arguments, completion word polled in pinned memory) and the two
 * functions return -1. */
int hfg_debug_set_timing(hfg_ctx *ctx, int on);

/* Test / profiling hooks (no reference counterpart): phase timeline of the last E-step kernel ([grid][12]: eight clock64
 * values + SM id) and the kernel's exponential applied to n host values. */
int hfg_debug_phase_clocks(hfg_ctx *ctx, long long *out, int *grid);
/* Benchmark hook: queues a write of `bytes` of scratch device memory on the context's stream (evicts the L2 between
 * timed iterations of the device-resident loop). */
int hfg_debug_l2_flush(hfg_ctx *ctx, size_t bytes);
/* Benchmark hook: what a C host (the drop-in binding, integration/hmm_estep_cuda.c) does per EM iteration, n_steps times:
 * hfg_em_iteration (host parameters in; host statistics, log-likelihood and labels out) + hfg_mstep; before every step
 * `flush_bytes` of scratch are written and waited for (0 = no flush), outside the step's interval.  step_seconds[i] = wall
 * time of step i (CLOCK_MONOTONIC around the two calls), logliks[i] its log-likelihood; params are updated in place. */
int hfg_debug_blocking_steps(hfg_ctx *ctx, const double *alpha, hfg_region_params *params, hfg_region_stats *stats,
                             int8_t *labels, int n_steps, size_t flush_bytes, double convergence_tol, double *step_seconds,
                             double *logliks);
int hfg_debug_exp(hfg_ctx *ctx, const double *in, double *out, int n);
/* Host-only (no GPU): self-check of the segment layout and observation-key builder for `capacity` segment slots;
 * summary[6] = {segments, windows per slot, edge windows, windows, distinct observation keys, statistics tiles}.  And EM_computeAdjustmentBeta (hmm.c:301-316) for one window of a chunk. */
int hfg_debug_layout_check(const hfg_config *cfg, int32_t n_chunks, const hfg_chunk_desc *chunks, const uint16_t *cov,
                           const uint16_t *cov_high_mapq, const uint16_t *cov_high_clip, const uint8_t *region,
                           int32_t capacity, int64_t *summary);
double hfg_debug_beta(const hfg_config *cfg, const hfg_chunk_desc *chunk, int window);
/* GPU: hfg_set_chunks builds the observation keys, their window lists and tiles on the device; this rebuilds them with the
 * host builder from the same inputs and compares every table bit for bit (HFG_OK = identical). */
int hfg_debug_layout_compare(hfg_ctx *ctx, int32_t n_chunks, const hfg_chunk_desc *chunks, const uint16_t *cov,
                             const uint16_t *cov_high_mapq, const uint16_t *cov_high_clip, const uint8_t *region);

#ifdef __cplusplus
}
#endif
#endif /* HFG_H */
