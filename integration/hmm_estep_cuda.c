/*
 * integration/hmm_estep_cuda.c -- the reference-side binding of libhfg.
 *
 * Drop this file into programs/submodules/hmm/ of mobinasri/flagger v1.2.0 (or link it ahead of hmm.o with the four
 * symbols below weakened) and `hmm_flagger` runs its E-step on the GPU.  It is written against
 * the reference's OWN headers and keeps the reference's signatures:
 *
 *     void    EM_runOneIterationForList(stList *emList, HMM *model, int threads);   hmm.h:109, hmm.c:739-780
 *     void    EM_runForwardForList   (stList *emList, HMM *model, int threads);    hmm.h:113, hmm.c:790-816
 *     double *EM_getPosterior        (EM *em, int pos);                            hmm.h:99,  hmm.c:671-685
 *     int     EM_getMostProbableState(EM *em, int pos);                            hmm.h:101, hmm.c:687-692
 *
 * Everything else of the binary -- CLI, .cov/.cov.gz/.bin parsing, window building, model construction, M-step,
 * SQUAREM, summary tables, BED/TSV writers -- stays the reference's code.  `threads` is ignored: the chunks of the list
 * are resident on the GPU after the first call.  Post-conditions honoured:
 *   1. model->loglikelihood = sum over chunks of sum_i log(scale_i)
 *   2. the model's ParameterEstimator / TransitionCountData arrays are INCREMENTED by this call's statistics
 *      (EM_updateModelEstimators, hmm.c:548-560), so HMM_estimateParameters / HMM_resetEstimators work unchanged
 *   3. every window's Inference.prediction is set (hmm.c:730-736)
 *   4. EM_getPosterior serves --writePosteriorProbs from the library's posterior buffer
 * Fatal conditions keep the reference's behaviour: message on stderr and exit(EXIT_FAILURE) (hmm.c:412-415).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "hmm.h"
#include "hmm_utils.h"
#include "chunk.h"
#include "ptBlock.h"
#include "hfg.h"

typedef struct Binding {
    stList *emList;   /* identity of the chunk set that is resident */
    hfg_ctx *ctx;
    hfg_config cfg;
    int nChunks;
    int64_t nWindows;
    EM **ems;         /* list order */
    int64_t *offsets; /* first window of each chunk in the flat arrays */
    int8_t *labels;
    hfg_region_params *params;
    hfg_region_stats *stats;
    double *posteriors; /* nWindows x 4, fetched lazily after an E-step */
    int posteriorsValid;
} Binding;

static Binding g_bind;

static void die(const char *what, hfg_ctx *ctx) {
    fprintf(stderr, "[hmm_estep_cuda] %s: %s\n", what, hfg_last_error(ctx));
    exit(EXIT_FAILURE);
}

static int state_is_gaussian(HMM *model, int s) {
    return model->emissionDistSeriesPerRegion[0]->emissionDists[s]->distType == DIST_GAUSSIAN;
}

/* HMM -> flat parameter block (include/hfg.h: hfg_region_params) */
static void flatten_params(HMM *model, hfg_region_params *out) {
    const int N = model->numberOfStates;
    for (int r = 0; r < model->numberOfRegions; r++) {
        hfg_region_params *p = &out[r];
        memset(p, 0, sizeof(*p));
        EmissionDistSeries *eds = model->emissionDistSeriesPerRegion[r];
        for (int s = 0; s < N; s++) {
            EmissionDist *ed = eds->emissionDists[s];
            if (ed->distType == DIST_TRUNC_EXPONENTIAL) {
                TruncExponential *te = (TruncExponential *) ed->dist;
                p->lambda = te->lambda;
                p->trunc_point = te->truncPoint;
            } else if (ed->distType == DIST_NEGATIVE_BINOMIAL) {
                /* theta / lambda travel in the mean / var slots (include/hfg.h, enum hfg_model_type) */
                NegativeBinomial *nb = (NegativeBinomial *) ed->dist;
                for (int c = 0; c < nb->numberOfComps; c++) {
                    p->mean[s][c] = nb->theta[c];
                    p->var[s][c] = nb->lambda[c];
                    p->weight[s][c] = nb->weights[c];
                }
            } else {
                Gaussian *g = (Gaussian *) ed->dist;
                for (int c = 0; c < g->numberOfComps; c++) {
                    p->mean[s][c] = g->mean[c];
                    p->var[s][c] = g->var[c];
                    p->weight[s][c] = g->weights[c];
                }
            }
        }
        MatrixDouble *m = model->transitionPerRegion[r]->matrix;
        for (int i = 0; i <= N; i++)
            for (int j = 0; j <= N; j++) p->trans[i][j] = m->data[i][j];
    }
}

/* flat statistics -> += into the model's estimators (what EM_updateModelEstimators does chunk by chunk) */
static void add_stats(HMM *model, const hfg_region_stats *stats) {
    const int N = model->numberOfStates;
    for (int r = 0; r < model->numberOfRegions; r++) {
        const hfg_region_stats *st = &stats[r];
        EmissionDistSeries *eds = model->emissionDistSeriesPerRegion[r];
        for (int s = 0; s < N; s++) {
            EmissionDist *ed = eds->emissionDists[s];
            if (ed->distType == DIST_TRUNC_EXPONENTIAL) {
                TruncExponential *te = (TruncExponential *) ed->dist;
                te->lambdaEstimator->numeratorPerComp[0] += st->lambda_num;
                te->lambdaEstimator->denominatorPerComp[0] += st->lambda_den;
            } else if (ed->distType == DIST_NEGATIVE_BINOMIAL) {
                NegativeBinomial *nb = (NegativeBinomial *) ed->dist;
                for (int c = 0; c < nb->numberOfComps; c++) {
                    nb->thetaEstimator->numeratorPerComp[c] += st->mean_num[s][c];
                    nb->thetaEstimator->denominatorPerComp[c] += st->mean_den[s][c];
                    nb->lambdaEstimator->numeratorPerComp[c] += st->var_num[s][c];
                    nb->lambdaEstimator->denominatorPerComp[c] += st->var_den[s][c];
                    nb->weightsEstimator->numeratorPerComp[c] += st->weight_num[s][c];
                    nb->weightsEstimator->denominatorPerComp[c] += st->weight_den[s][c];
                }
            } else {
                Gaussian *g = (Gaussian *) ed->dist;
                for (int c = 0; c < g->numberOfComps; c++) {
                    g->meanEstimator->numeratorPerComp[c] += st->mean_num[s][c];
                    g->meanEstimator->denominatorPerComp[c] += st->mean_den[s][c];
                    g->varEstimator->numeratorPerComp[c] += st->var_num[s][c];
                    g->varEstimator->denominatorPerComp[c] += st->var_den[s][c];
                    g->weightsEstimator->numeratorPerComp[c] += st->weight_num[s][c];
                    g->weightsEstimator->denominatorPerComp[c] += st->weight_den[s][c];
                }
            }
        }
        MatrixDouble *cm = model->transitionPerRegion[r]->transitionCountData->countMatrix;
        for (int i = 0; i < N; i++)
            for (int j = 0; j < N; j++) cm->data[i][j] += st->trans_count[i][j];
    }
}

/* first call for a list: describe the run to the library and upload every chunk once */
static void bind(stList *emList, HMM *model) {
    Binding *b = &g_bind;
    if (b->ctx) {
        hfg_destroy(b->ctx);
        free(b->ems); free(b->offsets); hfg_host_free(b->labels); free(b->params); free(b->stats); free(b->posteriors);
        memset(b, 0, sizeof(*b));
    }
    if (model->numberOfStates != HFG_NUM_STATES) {
        fprintf(stderr, "[hmm_estep_cuda] only the 4-state models run on the GPU path\n");
        exit(EXIT_FAILURE);
    }
    b->emList = emList;
    b->nChunks = (int) stList_length(emList);
    b->ems = malloc(sizeof(EM *) * b->nChunks);
    b->offsets = malloc(sizeof(int64_t) * (b->nChunks + 1));
    hfg_chunk_desc *desc = calloc(b->nChunks, sizeof(hfg_chunk_desc));
    int64_t W = 0;
    for (int c = 0; c < b->nChunks; c++) {
        EM *em = stList_get(emList, c);
        b->ems[c] = em;
        b->offsets[c] = W;
        desc[c].ctg_len = em->chunk->ctgLen;
        desc[c].s = em->chunk->s;
        desc[c].e = em->chunk->e;
        desc[c].window_len = em->chunk->windowLen;
        desc[c].n_windows = em->seqLen;
        desc[c].offset = W;
        W += em->seqLen;
    }
    b->offsets[b->nChunks] = W;
    b->nWindows = W;
    uint16_t *cov = malloc(sizeof(uint16_t) * W), *mq = malloc(sizeof(uint16_t) * W), *cl = malloc(sizeof(uint16_t) * W);
    uint8_t *reg = malloc(W);
    for (int c = 0; c < b->nChunks; c++) {
        EM *em = b->ems[c];
        for (int i = 0; i < em->seqLen; i++) {
            CoverageInfo *ci = em->coverageInfoSeq[i];
            const int64_t k = b->offsets[c] + i;
            cov[k] = ci->coverage;
            mq[k] = ci->coverage_high_mapq;
            cl[k] = ci->coverage_high_clip;
            reg[k] = (uint8_t) CoverageInfo_getRegionIndex(ci);
        }
    }
    EM *em0 = b->ems[0];
    TransitionRequirements *req = model->transitionPerRegion[0]->requirements;
    memset(&b->cfg, 0, sizeof(b->cfg));
    b->cfg.model_type = model->modelType == MODEL_GAUSSIAN ? HFG_MODEL_GAUSSIAN
                        : model->modelType == MODEL_NEGATIVE_BINOMIAL ? HFG_MODEL_NEGATIVE_BINOMIAL
                                                                      : HFG_MODEL_TRUNC_EXP_GAUSSIAN;
    b->cfg.n_regions = model->numberOfRegions;
    for (int s = 0; s < HFG_NUM_STATES; s++)
        b->cfg.n_comps[s] = EmissionDistSeries_getNumberOfComps(model->emissionDistSeriesPerRegion[0], s);
    b->cfg.adjust_contig_ends = em0->adjustContigEnds ? 1 : 0;
    b->cfg.mean_read_length = em0->meanReadLength;
    b->cfg.min_read_fraction_at_ends = em0->minReadFractionAtEnds;
    b->cfg.max_high_mapq_ratio = req->maxHighMapqRatio;
    b->cfg.min_high_mapq_ratio = req->minHighMapqRatio;
    b->cfg.min_highly_clipped_ratio = req->minHighlyClippedRatio;
    b->cfg.device = getenv("HFG_DEVICE") ? atoi(getenv("HFG_DEVICE")) : 0;
    if (hfg_create(&b->ctx, &b->cfg) != HFG_OK) die("hfg_create", NULL);
    if (hfg_set_chunks(b->ctx, b->nChunks, desc, cov, mq, cl, reg) != HFG_OK) die("hfg_set_chunks", b->ctx);
    free(desc); free(cov); free(mq); free(cl); free(reg);
    b->labels = hfg_host_alloc(W); /* page-locked: the device writes the labels into it directly */
    b->params = malloc(sizeof(hfg_region_params) * model->numberOfRegions);
    b->stats = malloc(sizeof(hfg_region_stats) * model->numberOfRegions);
    b->posteriors = NULL;
    fprintf(stderr, "[hmm_estep_cuda] %d chunks / %lld windows resident on GPU %d\n", b->nChunks, (long long) W,
            b->cfg.device);
}

static void flatten_alpha(HMM *model, double *alpha) {
    for (int i = 0; i < HFG_NUM_STATES; i++)
        for (int j = 0; j < HFG_NUM_STATES; j++) alpha[i * HFG_NUM_STATES + j] = model->alpha->data[i][j];
}

static int still_bound(stList *emList);

void EM_runOneIterationForList(stList *emList, HMM *model, int threads) {
    (void) threads;
    Binding *b = &g_bind;
    if (!still_bound(emList)) bind(emList, model);
    double alpha[HFG_NUM_STATES * HFG_NUM_STATES], loglik = 0.0;
    flatten_alpha(model, alpha);
    flatten_params(model, b->params);
    const int rc = hfg_em_iteration(b->ctx, alpha, b->params, b->stats, &loglik, b->labels);
    if (rc == HFG_ERR_SCALE_UNDERFLOW || rc == HFG_ERR_NAN) { /* the reference's two fatal conditions */
        fprintf(stderr, "%s\n", hfg_last_error(b->ctx));
        exit(EXIT_FAILURE);
    }
    if (rc != HFG_OK) die("hfg_em_iteration", b->ctx);
    model->loglikelihood = loglik;
    add_stats(model, b->stats);
    for (int c = 0; c < b->nChunks; c++) {
        EM *em = b->ems[c];
        em->model = model; /* EM_renewParametersAndEstimatorsFromModel keeps em->model current (hmm.c:297) */
        const int8_t *lab = b->labels + b->offsets[c];
        for (int i = 0; i < em->seqLen; i++) {
            CoverageInfo *ci = em->coverageInfoSeq[i];
            if (ci->data != NULL) ((Inference *) ci->data)->prediction = lab[i];
        }
    }
    b->posteriorsValid = 0;
}

void EM_runForwardForList(stList *emList, HMM *model, int threads) {
    (void) threads;
    Binding *b = &g_bind;
    if (!still_bound(emList)) bind(emList, model);
    double alpha[HFG_NUM_STATES * HFG_NUM_STATES], loglik = 0.0;
    flatten_alpha(model, alpha);
    flatten_params(model, b->params);
    const int rc = hfg_forward_only(b->ctx, alpha, b->params, &loglik);
    if (rc == HFG_ERR_SCALE_UNDERFLOW || rc == HFG_ERR_NAN) {
        fprintf(stderr, "%s\n", hfg_last_error(b->ctx));
        exit(EXIT_FAILURE);
    }
    if (rc != HFG_OK) die("hfg_forward_only", b->ctx);
    model->loglikelihood = loglik;
    for (int c = 0; c < b->nChunks; c++) b->ems[c]->model = model;
}

/* global window of position `pos` of a chunk.  Callers walk a chunk window by window and the chunks in list order, so the
 * chunk found last time (or the one after it) is tried first: O(1) per call instead of a scan over all chunks */
static int64_t window_index(EM *em, int pos) {
    Binding *b = &g_bind;
    static int last = 0;
    if (last >= b->nChunks) last = 0;
    if (b->ems[last] == em) return b->offsets[last] + pos;
    if (last + 1 < b->nChunks && b->ems[last + 1] == em) return b->offsets[++last] + pos;
    for (int c = 0; c < b->nChunks; c++)
        if (b->ems[c] == em) {
            last = c;
            return b->offsets[c] + pos;
        }
    fprintf(stderr, "[hmm_estep_cuda] EM_getPosterior on an EM that is not resident\n");
    exit(EXIT_FAILURE);
}

/* Is the resident copy still the list the caller passes?  The address of the stList alone is not enough: a caller that
 * runs several jobs in one process (the alpha tuner) may rebuild the list at the same address. */
static int still_bound(stList *emList) {
    Binding *b = &g_bind;
    if (b->ctx == NULL || b->emList != emList || stList_length(emList) != b->nChunks || b->nChunks < 1) return 0;
    EM *first = stList_get(emList, 0), *last = stList_get(emList, b->nChunks - 1);
    return first == b->ems[0] && last == b->ems[b->nChunks - 1] &&
           b->offsets[b->nChunks - 1] + last->seqLen == b->nWindows;
}

double *EM_getPosterior(EM *em, int pos) {
    Binding *b = &g_bind;
    if (!b->posteriorsValid) {
        if (!b->posteriors) b->posteriors = malloc(sizeof(double) * 4 * b->nWindows);
        if (hfg_get_posteriors(b->ctx, b->posteriors) != HFG_OK) die("hfg_get_posteriors", b->ctx);
        b->posteriorsValid = 1;
    }
    double *out = malloc(sizeof(double) * HFG_NUM_STATES); /* caller frees, as in the reference */
    memcpy(out, b->posteriors + 4 * window_index(em, pos), sizeof(double) * HFG_NUM_STATES);
    return out;
}

int EM_getMostProbableState(EM *em, int pos) { return g_bind.labels[window_index(em, pos)]; }
