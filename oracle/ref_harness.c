/*
 * oracle/ref_harness.c -- TEST INFRASTRUCTURE ONLY.  NOT PART OF THE PRODUCT.
 *
 * Thin driver that exposes the UNMODIFIED reference implementation (compiled from
 * /root/reference/programs by oracle/Makefile into oracle/_ref/libref_harness.so) behind the
 * same flat-array signatures as oracle/hmm_oracle.c, so that tests can compare
 *      reference  <->  restatement  <->  CUDA path
 * on identical inputs at full double precision (the reference binary itself only prints %.4f / %.5e).
 *
 * This file contains no HMM arithmetic: it builds the reference's own structs (Chunk, CoverageInfo,
 * HMM, EM), calls the reference's own functions (createModel, EM_runOneIterationForList,
 * EM_runForwardForList, EM_getPosterior, HMM_estimateParameters) and copies the results out.
 */
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <time.h>

#include "hmm.h"
#include "hmm_utils.h"
#include "chunk.h"
#include "data_types.h"
#include "track_reader.h"
#include "../include/hfg.h"

/* defined in the reference's src/hmm_flagger.c (compiled with -Dmain=ref_hmm_flagger_main) */
HMM *createModel(ModelType modelType, int numberOfCollapsedComps, CoverageHeader *header, MatrixDouble *alphaMatrix,
                 double maxHighMapqRatio, double minHighMapqRatio, int windowLen, double initialDeviation);
int getBestNumberOfCollapsedComps(ChunksCreator *chunksCreator);

#define NS HFG_NUM_STATES

static CoverageHeader *make_header(const hfg_config *cfg, const int32_t *regionCov, int startOnly) {
    CoverageHeader *h = calloc(1, sizeof(CoverageHeader));
    h->numberOfRegions = cfg->n_regions;
    h->regionCoverages = malloc(sizeof(int) * cfg->n_regions);
    for (int r = 0; r < cfg->n_regions; r++) h->regionCoverages[r] = regionCov ? regionCov[r] : 40;
    h->startOnlyMode = startOnly;
    h->averageAlignmentLength = cfg->mean_read_length;
    h->numberOfLabels = 4;
    return h;
}

static MatrixDouble *make_alpha(const double *alpha) {
    MatrixDouble *m = MatrixDouble_construct0(NS, NS);
    for (int i = 0; i < NS; i++)
        for (int j = 0; j < NS; j++) m->data[i][j] = alpha ? alpha[i * NS + j] : 0.0;
    return m;
}

/* model_type 2 = negative binomial (oracle/hmm_oracle.h ORC_MODEL_NEGATIVE_BINOMIAL; not part of the product ABI yet):
 * hfg_region_params.mean holds theta, .var holds lambda, .weight the mixture weights (NegativeBinomial, hmm_utils.h) */
static ModelType model_type(const hfg_config *cfg) {
    if (cfg->model_type == 2) return MODEL_NEGATIVE_BINOMIAL;
    return cfg->model_type == HFG_MODEL_GAUSSIAN ? MODEL_GAUSSIAN : MODEL_TRUNC_EXP_GAUSSIAN;
}

/* reference model with the reference's own initial values */
static HMM *make_model(const hfg_config *cfg, const int32_t *regionCov, int windowLen, int startOnly,
                       const double *alpha) {
    CoverageHeader *h = make_header(cfg, regionCov, startOnly);
    MatrixDouble *a = make_alpha(alpha);
    HMM *model = createModel(model_type(cfg), cfg->n_comps[HFG_STATE_COL], h, a, cfg->max_high_mapq_ratio,
                             cfg->min_high_mapq_ratio, windowLen, 0.0);
    MatrixDouble_destruct(a);
    free(h->regionCoverages);
    free(h);
    return model;
}

static void params_to_model(const hfg_config *cfg, const hfg_region_params *params, HMM *model) {
    for (int r = 0; r < cfg->n_regions; r++) {
        EmissionDistSeries *eds = model->emissionDistSeriesPerRegion[r];
        for (int s = 0; s < NS; s++) {
            EmissionDist *ed = eds->emissionDists[s];
            if (ed->distType == DIST_TRUNC_EXPONENTIAL) {
                TruncExponential *te = ed->dist;
                te->lambda = params[r].lambda;
                te->truncPoint = params[r].trunc_point;
            } else if (ed->distType == DIST_NEGATIVE_BINOMIAL) {
                NegativeBinomial *nb = ed->dist;
                for (int c = 0; c < nb->numberOfComps; c++) {
                    nb->theta[c] = params[r].mean[s][c];
                    nb->lambda[c] = params[r].var[s][c];
                    nb->weights[c] = params[r].weight[s][c];
                }
                NegativeBinomial_fillDigammaTable(nb); /* as after every parameter update (hmm_utils.c:1893-1899) */
            } else {
                Gaussian *g = ed->dist;
                for (int c = 0; c < g->numberOfComps; c++) {
                    g->mean[c] = params[r].mean[s][c];
                    g->var[c] = params[r].var[s][c];
                    g->weights[c] = params[r].weight[s][c];
                }
            }
        }
        Transition *t = model->transitionPerRegion[r];
        for (int i = 0; i < NS + 1; i++)
            for (int j = 0; j < NS + 1; j++) t->matrix->data[i][j] = params[r].trans[i][j];
    }
}

static void model_to_params(const hfg_config *cfg, HMM *model, hfg_region_params *params) {
    for (int r = 0; r < cfg->n_regions; r++) {
        memset(&params[r], 0, sizeof(hfg_region_params));
        EmissionDistSeries *eds = model->emissionDistSeriesPerRegion[r];
        for (int s = 0; s < NS; s++) {
            EmissionDist *ed = eds->emissionDists[s];
            if (ed->distType == DIST_TRUNC_EXPONENTIAL) {
                TruncExponential *te = ed->dist;
                params[r].lambda = te->lambda;
                params[r].trunc_point = te->truncPoint;
            } else if (ed->distType == DIST_NEGATIVE_BINOMIAL) {
                NegativeBinomial *nb = ed->dist;
                for (int c = 0; c < nb->numberOfComps; c++) {
                    params[r].mean[s][c] = nb->theta[c];
                    params[r].var[s][c] = nb->lambda[c];
                    params[r].weight[s][c] = nb->weights[c];
                }
            } else {
                Gaussian *g = ed->dist;
                for (int c = 0; c < g->numberOfComps; c++) {
                    params[r].mean[s][c] = g->mean[c];
                    params[r].var[s][c] = g->var[c];
                    params[r].weight[s][c] = g->weights[c];
                }
            }
        }
        Transition *t = model->transitionPerRegion[r];
        for (int i = 0; i < NS + 1; i++)
            for (int j = 0; j < NS + 1; j++) params[r].trans[i][j] = t->matrix->data[i][j];
    }
}

static void model_to_stats(const hfg_config *cfg, HMM *model, hfg_region_stats *stats) {
    for (int r = 0; r < cfg->n_regions; r++) {
        memset(&stats[r], 0, sizeof(hfg_region_stats));
        EmissionDistSeries *eds = model->emissionDistSeriesPerRegion[r];
        for (int s = 0; s < NS; s++) {
            EmissionDist *ed = eds->emissionDists[s];
            if (ed->distType == DIST_TRUNC_EXPONENTIAL) {
                TruncExponential *te = ed->dist;
                stats[r].lambda_num = te->lambdaEstimator->numeratorPerComp[0];
                stats[r].lambda_den = te->lambdaEstimator->denominatorPerComp[0];
            } else if (ed->distType == DIST_NEGATIVE_BINOMIAL) {
                NegativeBinomial *nb = ed->dist; /* theta -> mean_*, lambda -> var_* */
                for (int c = 0; c < nb->numberOfComps; c++) {
                    stats[r].mean_num[s][c] = nb->thetaEstimator->numeratorPerComp[c];
                    stats[r].mean_den[s][c] = nb->thetaEstimator->denominatorPerComp[c];
                    stats[r].var_num[s][c] = nb->lambdaEstimator->numeratorPerComp[c];
                    stats[r].var_den[s][c] = nb->lambdaEstimator->denominatorPerComp[c];
                    stats[r].weight_num[s][c] = nb->weightsEstimator->numeratorPerComp[c];
                    stats[r].weight_den[s][c] = nb->weightsEstimator->denominatorPerComp[c];
                }
            } else {
                Gaussian *g = ed->dist;
                for (int c = 0; c < g->numberOfComps; c++) {
                    stats[r].mean_num[s][c] = g->meanEstimator->numeratorPerComp[c];
                    stats[r].mean_den[s][c] = g->meanEstimator->denominatorPerComp[c];
                    stats[r].var_num[s][c] = g->varEstimator->numeratorPerComp[c];
                    stats[r].var_den[s][c] = g->varEstimator->denominatorPerComp[c];
                    stats[r].weight_num[s][c] = g->weightsEstimator->numeratorPerComp[c];
                    stats[r].weight_den[s][c] = g->weightsEstimator->denominatorPerComp[c];
                }
            }
        }
        MatrixDouble *cm = model->transitionPerRegion[r]->transitionCountData->countMatrix;
        for (int i = 0; i < NS; i++)
            for (int j = 0; j < NS; j++) stats[r].trans_count[i][j] = cm->data[i][j];
    }
}

static void stats_to_model(const hfg_config *cfg, const hfg_region_stats *stats, HMM *model) {
    for (int r = 0; r < cfg->n_regions; r++) {
        EmissionDistSeries *eds = model->emissionDistSeriesPerRegion[r];
        for (int s = 0; s < NS; s++) {
            EmissionDist *ed = eds->emissionDists[s];
            if (ed->distType == DIST_TRUNC_EXPONENTIAL) {
                TruncExponential *te = ed->dist;
                te->lambdaEstimator->numeratorPerComp[0] = stats[r].lambda_num;
                te->lambdaEstimator->denominatorPerComp[0] = stats[r].lambda_den;
            } else if (ed->distType == DIST_NEGATIVE_BINOMIAL) {
                NegativeBinomial *nb = ed->dist;
                for (int c = 0; c < nb->numberOfComps; c++) {
                    nb->thetaEstimator->numeratorPerComp[c] = stats[r].mean_num[s][c];
                    nb->thetaEstimator->denominatorPerComp[c] = stats[r].mean_den[s][c];
                    nb->lambdaEstimator->numeratorPerComp[c] = stats[r].var_num[s][c];
                    nb->lambdaEstimator->denominatorPerComp[c] = stats[r].var_den[s][c];
                    nb->weightsEstimator->numeratorPerComp[c] = stats[r].weight_num[s][c];
                    nb->weightsEstimator->denominatorPerComp[c] = stats[r].weight_den[s][c];
                }
            } else {
                Gaussian *g = ed->dist;
                for (int c = 0; c < g->numberOfComps; c++) {
                    g->meanEstimator->numeratorPerComp[c] = stats[r].mean_num[s][c];
                    g->meanEstimator->denominatorPerComp[c] = stats[r].mean_den[s][c];
                    g->varEstimator->numeratorPerComp[c] = stats[r].var_num[s][c];
                    g->varEstimator->denominatorPerComp[c] = stats[r].var_den[s][c];
                    g->weightsEstimator->numeratorPerComp[c] = stats[r].weight_num[s][c];
                    g->weightsEstimator->denominatorPerComp[c] = stats[r].weight_den[s][c];
                }
            }
        }
        MatrixDouble *cm = model->transitionPerRegion[r]->transitionCountData->countMatrix;
        for (int i = 0; i < NS; i++)
            for (int j = 0; j < NS; j++) cm->data[i][j] = stats[r].trans_count[i][j];
    }
}

typedef struct RefData {
    stList *chunks; /* Chunk* */
    stList *ems;    /* EM* */
    HMM *model;
} RefData;

static RefData *build(const hfg_config *cfg, int n_chunks, const hfg_chunk_desc *cd, const uint16_t *cov,
                      const uint16_t *mapq, const uint16_t *clip, const uint8_t *region, const double *alpha,
                      const hfg_region_params *params) {
    RefData *d = calloc(1, sizeof(RefData));
    d->model = make_model(cfg, NULL, cd[0].window_len, 0, alpha);
    params_to_model(cfg, params, d->model);
    HMM_resetEstimators(d->model);
    d->chunks = stList_construct3(0, NULL);
    d->ems = stList_construct3(0, NULL);
    for (int c = 0; c < n_chunks; c++) {
        Chunk *chunk = Chunk_constructWithAllocatedSeq(20000000, cd[c].window_len, cd[c].n_windows);
        strcpy(chunk->ctg, "ctg");
        chunk->ctgLen = cd[c].ctg_len;
        chunk->s = cd[c].s;
        chunk->e = cd[c].e;
        chunk->coverageInfoSeqLen = cd[c].n_windows;
        for (int i = 0; i < cd[c].n_windows; i++) {
            int64_t k = cd[c].offset + i;
            CoverageInfo *ci = chunk->coverageInfoSeq[i];
            ci->coverage = cov[k];
            ci->coverage_high_mapq = mapq[k];
            ci->coverage_high_clip = clip[k];
            ci->annotation_flag = 0ULL;
            CoverageInfo_setRegionIndex(ci, region[k]);
            CoverageInfo_addInferenceData(ci, -1, -1);
        }
        stList_append(d->chunks, chunk);
        EM *em = EM_construct(chunk->coverageInfoSeq, chunk->coverageInfoSeqLen, d->model, chunk, cfg->mean_read_length);
        if (cfg->adjust_contig_ends) EM_setMinReadFractionAtEnds(em, cfg->min_read_fraction_at_ends);
        stList_append(d->ems, em);
    }
    return d;
}

static void teardown(RefData *d) {
    for (int c = 0; c < stList_length(d->ems); c++) {
        EM *em = stList_get(d->ems, c);
        EM_destruct(em);
        free(em);
        Chunk_destruct(stList_get(d->chunks, c));
    }
    stList_destruct(d->ems);
    stList_destruct(d->chunks);
    HMM_destruct(d->model);
    free(d->model);
    free(d);
}

int ref_estep(const hfg_config *cfg, int n_chunks, const hfg_chunk_desc *cd, const uint16_t *cov,
              const uint16_t *mapq, const uint16_t *clip, const uint8_t *region, const double *alpha,
              const hfg_region_params *params, hfg_region_stats *stats, double *loglik, double *chunk_logliks,
              int8_t *labels, double *posteriors, double *fwd, double *bwd, double *scales_out, int forward_only,
              int threads) {
    RefData *d = build(cfg, n_chunks, cd, cov, mapq, clip, region, alpha, params);
    if (forward_only) EM_runForwardForList(d->ems, d->model, threads);
    else EM_runOneIterationForList(d->ems, d->model, threads);
    if (loglik) *loglik = d->model->loglikelihood;
    if (stats && !forward_only) model_to_stats(cfg, d->model, stats);
    for (int c = 0; c < n_chunks; c++) {
        EM *em = stList_get(d->ems, c);
        if (chunk_logliks) chunk_logliks[c] = em->loglikelihood;
        for (int i = 0; i < cd[c].n_windows; i++) {
            int64_t k = cd[c].offset + i;
            if (fwd) memcpy(fwd + k * NS, em->f[i], sizeof(double) * NS);
            if (scales_out) scales_out[k] = em->scales[i];
            if (forward_only) continue;
            if (bwd) memcpy(bwd + k * NS, em->b[i], sizeof(double) * NS);
            if (labels) labels[k] = ((Inference *) em->coverageInfoSeq[i]->data)->prediction;
            if (posteriors) {
                double *p = EM_getPosterior(em, i);
                memcpy(posteriors + k * NS, p, sizeof(double) * NS);
                free(p);
            }
        }
    }
    teardown(d);
    return 0;
}

int ref_model_init(const hfg_config *cfg, const int32_t *region_coverages, int window_len, int start_only_mode,
                   hfg_region_params *params) {
    HMM *model = make_model(cfg, region_coverages, window_len, start_only_mode, NULL);
    model_to_params(cfg, model, params);
    HMM_destruct(model);
    free(model);
    return 0;
}

int ref_mstep(const hfg_config *cfg, hfg_region_params *params, const hfg_region_stats *stats, double tol,
              int *converged_out) {
    HMM *model = make_model(cfg, NULL, 4000, 0, NULL);
    params_to_model(cfg, params, model);
    stats_to_model(cfg, stats, model);
    *converged_out = HMM_estimateParameters(model, tol) ? 1 : 0;
    model_to_params(cfg, model, params);
    HMM_destruct(model);
    free(model);
    return 0;
}

/* the EM loop of runHMMFlagger (src/hmm_flagger.c:337-467) driven with the reference's own functions */
int ref_run_em(const hfg_config *cfg, int n_chunks, const hfg_chunk_desc *cd, const uint16_t *cov,
               const uint16_t *mapq, const uint16_t *clip, const uint8_t *region, const double *alpha,
               hfg_region_params *params, int max_iterations, double tol, double *logliks, int *n_esteps,
               int8_t *labels, int threads, double *estep_seconds) {
    RefData *d = build(cfg, n_chunks, cd, cov, mapq, clip, region, alpha, params);
    int iter = 1, k = 0;
    bool converged = false;
    double secs = 0.0;
    struct timespec t0, t1;
    while (iter <= max_iterations && !converged) {
        clock_gettime(CLOCK_MONOTONIC, &t0);
        EM_runOneIterationForList(d->ems, d->model, threads);
        clock_gettime(CLOCK_MONOTONIC, &t1);
        secs += (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
        logliks[k++] = d->model->loglikelihood;
        converged = HMM_estimateParameters(d->model, tol);
        HMM_resetEstimators(d->model);
        iter++;
    }
    clock_gettime(CLOCK_MONOTONIC, &t0);
    EM_runOneIterationForList(d->ems, d->model, threads);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    secs += (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
    logliks[k++] = d->model->loglikelihood;
    *n_esteps = k;
    if (estep_seconds) *estep_seconds = secs;
    model_to_params(cfg, d->model, params);
    if (labels) {
        for (int c = 0; c < n_chunks; c++) {
            EM *em = stList_get(d->ems, c);
            for (int i = 0; i < cd[c].n_windows; i++)
                labels[cd[c].offset + i] = ((Inference *) em->coverageInfoSeq[i]->data)->prediction;
        }
    }
    teardown(d);
    return 0;
}

/* persistent handle for benchmarking: build the reference structs once, then time whole EM iterations
 * (EM_runOneIterationForList + HMM_estimateParameters + HMM_resetEstimators, src/hmm_flagger.c:344,419,425) */
void *ref_open(const hfg_config *cfg, int n_chunks, const hfg_chunk_desc *cd, const uint16_t *cov,
               const uint16_t *mapq, const uint16_t *clip, const uint8_t *region, const double *alpha,
               const hfg_region_params *params) {
    return build(cfg, n_chunks, cd, cov, mapq, clip, region, alpha, params);
}

double ref_step(void *handle, int threads, double tol, double *loglik) {
    RefData *d = handle;
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    EM_runOneIterationForList(d->ems, d->model, threads);
    if (loglik) *loglik = d->model->loglikelihood;
    HMM_estimateParameters(d->model, tol);
    HMM_resetEstimators(d->model);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}

void ref_close(void *handle) { teardown(handle); }

/* getBestNumberOfCollapsedComps (src/hmm_flagger.c:105-111) on a one-window stand-in chunk that carries the maximum
 * coverage, followed by the clamp that main() applies (src/hmm_flagger.c:1012-1013). */
int ref_best_num_collapsed_comps(int max_coverage, const int32_t *region_coverages, int n_regions) {
    ChunksCreator cc;
    memset(&cc, 0, sizeof(cc));
    CoverageHeader header;
    memset(&header, 0, sizeof(header));
    header.numberOfRegions = n_regions;
    header.regionCoverages = malloc(sizeof(int) * n_regions);
    for (int r = 0; r < n_regions; r++) header.regionCoverages[r] = region_coverages[r];
    cc.header = &header;
    cc.chunks = stList_construct3(0, NULL);
    Chunk *chunk = Chunk_constructWithAllocatedSeq(20000000, 4000, 1);
    chunk->coverageInfoSeqLen = 1;
    chunk->coverageInfoSeq[0]->coverage = (u_int16_t) max_coverage;
    stList_append(cc.chunks, chunk);
    int k = getBestNumberOfCollapsedComps(&cc);
    k = k < 2 ? 2 : k;
    k = k > 10 ? 10 : k;
    Chunk_destruct(chunk);
    stList_destruct(cc.chunks);
    free(header.regionCoverages);
    return k;
}

/* ---- the reference's own .cov / .cov.gz chunk builder, flattened (checker for flagger_b200/csrc/hfg_cov_reader.c) ----
 * ChunksCreator_constructFromCov + ChunksCreator_parseChunks (chunk.c:141-236,486-547).  Note: the reference writes
 * `<path>.index` next to the input, so point it at a writable copy. */
void *ref_cov_open(const char *path, int chunkLen, int windowLen, int threads) {
    ChunksCreator *cc = ChunksCreator_constructFromCov((char *) path, NULL, chunkLen, threads, windowLen);
    if (ChunksCreator_parseChunks(cc) != 0) return NULL; /* chunks stay in file order, as in getChunksCreator (src/hmm_flagger.c:60-103) */
    return cc;
}

void ref_cov_counts(void *handle, int32_t *n_chunks, int64_t *n_windows, int32_t *header /* 8 ints */) {
    ChunksCreator *cc = handle;
    int64_t W = 0;
    for (int c = 0; c < stList_length(cc->chunks); c++) W += ((Chunk *) stList_get(cc->chunks, c))->coverageInfoSeqLen;
    *n_chunks = (int32_t) stList_length(cc->chunks);
    *n_windows = W;
    CoverageHeader *h = cc->header;
    header[0] = h->numberOfAnnotations;
    header[1] = h->numberOfRegions;
    header[2] = h->numberOfLabels;
    header[3] = h->isTruthAvailable;
    header[4] = h->isPredictionAvailable;
    header[5] = h->startOnlyMode;
    header[6] = h->averageAlignmentLength;
    header[7] = h->numberOfRegions > 0 ? h->regionCoverages[0] : 0;
}

void ref_cov_fill(void *handle, hfg_chunk_desc *chunks, char *names /* n_chunks x 200 */, uint16_t *cov, uint16_t *mapq,
                  uint16_t *clip, uint64_t *flags, int8_t *truth, int8_t *prediction, int32_t *region_coverages) {
    ChunksCreator *cc = handle;
    int64_t o = 0;
    for (int c = 0; c < stList_length(cc->chunks); c++) {
        Chunk *ch = stList_get(cc->chunks, c);
        memset(&chunks[c], 0, sizeof(hfg_chunk_desc));
        chunks[c].ctg_len = ch->ctgLen;
        chunks[c].s = ch->s;
        chunks[c].e = ch->e;
        chunks[c].window_len = ch->windowLen;
        chunks[c].n_windows = ch->coverageInfoSeqLen;
        chunks[c].offset = o;
        strncpy(names + (size_t) c * 200, ch->ctg, 199);
        for (int i = 0; i < ch->coverageInfoSeqLen; i++, o++) {
            CoverageInfo *ci = ch->coverageInfoSeq[i];
            cov[o] = ci->coverage;
            mapq[o] = ci->coverage_high_mapq;
            clip[o] = ci->coverage_high_clip;
            flags[o] = ci->annotation_flag;
            Inference *inf = ci->data;
            truth[o] = inf ? inf->truth : -1;
            prediction[o] = inf ? inf->prediction : -1;
        }
    }
    for (int r = 0; r < cc->header->numberOfRegions; r++) region_coverages[r] = cc->header->regionCoverages[r];
}

void ref_cov_close(void *handle) { ChunksCreator_destruct(handle); }

/* ---- SQUAREM host arithmetic (checker for hfg_squarem_* / hfg_params_feasible) ----
 * The reference's own SquareAccelerator (hmm.c:820-1098) fed with three parameter sets: computeRates +
 * computeValuesForModelPrime, then `n_shrinks` x shrinkAlphaAndRecomputeModelPrime(margin).  Returns HMM_isFeasible of the
 * resulting model. */
/* defined in hmm.c but not declared in hmm.h */
HMM *SquareAccelerator_shrinkAlphaAndRecomputeModelPrime(SquareAccelerator *accelerator, double alphaMargin);

int ref_squarem(const hfg_config *cfg, const hfg_region_params *p0, const hfg_region_params *p1,
                const hfg_region_params *p2, int n_shrinks, double margin, hfg_region_params *prime, double *alpha_rate) {
    HMM *m0 = make_model(cfg, NULL, 4000, 0, NULL), *m1 = make_model(cfg, NULL, 4000, 0, NULL),
        *m2 = make_model(cfg, NULL, 4000, 0, NULL);
    params_to_model(cfg, p0, m0);
    params_to_model(cfg, p1, m1);
    params_to_model(cfg, p2, m2);
    SquareAccelerator *acc = SquareAccelerator_construct();
    SquareAccelerator_setModel0(acc, m0);
    SquareAccelerator_setModel1(acc, m1);
    SquareAccelerator_setModel2(acc, m2);
    SquareAccelerator_computeRates(acc);
    HMM *mp = SquareAccelerator_computeValuesForModelPrime(acc);
    for (int i = 0; i < n_shrinks; i++) mp = SquareAccelerator_shrinkAlphaAndRecomputeModelPrime(acc, margin);
    model_to_params(cfg, mp, prime);
    *alpha_rate = acc->alphaRate;
    return HMM_isFeasible(mp) ? 1 : 0;
}

/* the accelerated EM loop of runHMMFlagger (src/hmm_flagger.c:337-467 with acceleration == true) driven with the reference's
 * own functions; logliks[k] = what loglikelihood.tsv holds, alpha_rates[k] = the accepted SQUAREM step */
int ref_run_em_accelerated(const hfg_config *cfg, int n_chunks, const hfg_chunk_desc *cd, const uint16_t *cov,
                           const uint16_t *mapq, const uint16_t *clip, const uint8_t *region, const double *alpha,
                           hfg_region_params *params, int max_iterations, double tol, double *logliks,
                           double *alpha_rates, int *n_outer, int8_t *labels, int threads) {
    RefData *d = build(cfg, n_chunks, cd, cov, mapq, clip, region, alpha, params);
    HMM *model = d->model;
    int iter = 1, k = 0;
    bool converged = false;
    while (iter <= max_iterations && !converged) {
        EM_runOneIterationForList(d->ems, model, threads);
        logliks[k] = model->loglikelihood;
        SquareAccelerator *accelerator = SquareAccelerator_construct();
        SquareAccelerator_setModel0(accelerator, model);
        HMM_estimateParameters(model, tol);
        SquareAccelerator_setModel1(accelerator, model);
        HMM_resetEstimators(model);
        EM_runOneIterationForList(d->ems, model, threads);
        HMM_estimateParameters(model, tol);
        SquareAccelerator_setModel2(accelerator, model);
        HMM *modelPrime = SquareAccelerator_getModelPrime(accelerator, d->ems, threads);
        if (alpha_rates) alpha_rates[k] = accelerator->alphaRate;
        k++;
        HMM_resetEstimators(modelPrime);
        EM_runOneIterationForList(d->ems, modelPrime, threads);
        HMM_destruct(model);
        model = HMM_copy(modelPrime);
        d->model = model;
        for (int c = 0; c < stList_length(d->ems); c++) EM_renewParametersAndEstimatorsFromModel(stList_get(d->ems, c), model);
        SquareAccelerator_destruct(accelerator);
        converged = HMM_estimateParameters(model, tol);
        HMM_resetEstimators(model);
        iter++;
    }
    EM_runOneIterationForList(d->ems, model, threads);
    logliks[k] = model->loglikelihood;
    *n_outer = k;
    model_to_params(cfg, model, params);
    if (labels) {
        for (int c = 0; c < n_chunks; c++) {
            EM *em = stList_get(d->ems, c);
            for (int i = 0; i < cd[c].n_windows; i++)
                labels[cd[c].offset + i] = ((Inference *) em->coverageInfoSeq[i]->data)->prediction;
        }
    }
    teardown(d);
    return 0;
}
