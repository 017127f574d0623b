/*
 * oracle/hmm_oracle.c -- TEST INFRASTRUCTURE ONLY.  NOT PART OF THE PRODUCT.
 *
 * Plain-C, single-threaded restatement of the HMM-Flagger v1.2.0 E-step / M-step
 * (mobinasri/flagger, programs/submodules/{hmm,hmm_utils}) used as the parity checker
 * for the CUDA path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load this file's library; the product (flagger_b200/)
 * never does.
 *
 * Pinning: the reference ships NO test, golden vector or known-answer fixture for this
 * arithmetic (programs/Makefile:38-59 lists its tests; none touches hmm.c/hmm_utils.c).
 * This restatement is instead pinned against the UNMODIFIED reference sources compiled
 * into oracle/_ref/ (see oracle/Makefile, tests/test_oracle_vs_reference.py) and against
 * fixtures generated from that build (tests/golden/, tests/golden/make_golden.py).
 *
 * Every function cites the reference file:line it restates (paths under programs/).
 * The floating-point operation ORDER of the reference is kept on purpose, so that with
 * -ffp-contract=off this file reproduces the reference bit for bit.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#include "../include/hfg.h" /* plain data layouts shared with the boundary (no code) */
#include "hmm_oracle.h"

#define NS HFG_NUM_STATES
#define ORC_PI 3.14159          /* submodules/common/common.h:15 */
#define ORC_TERM 1e-4           /* Transition.terminationProb, hmm_utils.c:2112 */
#define ORC_MIN_COUNT 10        /* MIN_COUNT_FOR_PARAMETER_UPDATE, hmm_utils.h:11 */
#define ORC_MAX_COV 250         /* MAX_COVERAGE_VALUE, hmm_utils.h:15 */
#define ORC_TRUNC_FRACTION 0.25 /* EXP_TRUNC_POINT_COV_FRACTION, hmm_utils.h:12 */
#define ORC_ERR_COEF 0.1        /* ERR_COMP_BINDING_COEF, hmm_utils.h:14 */
#define ORC_PSEUDO 0.001        /* TRANSITION_PSEUDO_COUNT_VALUE, hmm.c:16 */
#define ORC_DIAG 0.99           /* TRANSITION_INITIAL_DIAG_PROB, hmm.c:15 */

/* submodules/common/common.c:142-148: min/max take and return int, so double arguments are truncated */
static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a < b ? b : a; }

/* EM_computeAdjustmentBeta, hmm.c:301-316 */
double orc_beta(const hfg_config *cfg, const hfg_chunk_desc *ch, int i) {
    if (!cfg->adjust_contig_ends) return 1.0;
    double minFrac = cfg->min_read_fraction_at_ends;
    int mid = imin((int) (ch->s + (double) ch->window_len * (i + 0.5)),
                   (int) ((ch->s + (double) ch->window_len * i + ch->e) / 2));
    int Lr = cfg->mean_read_length;
    int l = imax(mid - Lr + 1, (int) (-(1 - minFrac) * Lr));
    int u = imin(mid, (int) (ch->ctg_len - minFrac * Lr));
    double beta = (double) (u - l) / Lr;
    if (beta <= 0.25) return 0.25;
    return beta;
}

/* true for the states whose parameters live in the mixture slots mean/var/weight of hfg_region_params: the Gaussians, and
 * every state of the negative-binomial model (theta in .mean, lambda in .var, see hmm_oracle.h) */
static int state_is_gaussian(const hfg_config *cfg, int s) {
    return !(cfg->model_type == HFG_MODEL_TRUNC_EXP_GAUSSIAN && s == HFG_STATE_ERR);
}

static int is_nb(const hfg_config *cfg) { return cfg->model_type == ORC_MODEL_NEGATIVE_BINOMIAL; }

/* ---- negative-binomial model (MODEL_NEGATIVE_BINOMIAL; hmm_utils.c:320-640) ------------------------------------- */

#include "digamma_coef.h" /* generated: tools/gen_digamma_coef.py */

/* digamma in long double, restating the routine the reference vendors under submodules/digamma (digamma.c:36-116,
 * R. J. Mathar 2005): reflection below 0, psi(x) = psi(1+x) - 1/x below 1, the duplication formula above 3, exact values at
 * 1, 2, 3 and the Chebyshev series in T_n(x - 2) on (1, 3).  Same operations in the same order, so the long-double results
 * agree bit for bit (tests/test_oracle_nb.py). */
long double orc_digammal(long double x) {
    const long double pi = 3.14159265358979323846264338327950288L;
    const long double euler = 0.577215664901532860606512090082402431L;
    const long double ln2 = 0.693147180559945309417232121458176568L;
    if (x < 0.0L) return orc_digammal(1.0L - x) + pi / tanl(pi * (1.0L - x));
    if (x < 1.0L) return orc_digammal(1.0L + x) - 1.0L / x;
    if (x == 1.0L) return -euler;
    if (x == 2.0L) return 1.0L - euler;
    if (x == 3.0L) return 1.5L - euler;
    if (x > 3.0L) return 0.5L * (orc_digammal(x / 2.0L) + orc_digammal((x + 1.0L) / 2.0L)) + ln2;
    long double t = x - 2.0L;
    long double prev = 1.0L, cur = t;
    long double acc = orc_digamma_coef[0] + orc_digamma_coef[1] * cur;
    for (int n = 2; n < ORC_DIGAMMA_NCOEF; n++) {
        long double next = 2.0L * t * cur - prev; /* T_{n} = 2 t T_{n-1} - T_{n-2} */
        acc += orc_digamma_coef[n] * next;
        prev = cur;
        cur = next;
    }
    return acc;
}

/* NegativeBinomial_getR, hmm_utils.c:455-458 */
static double nb_r(double theta, double lambda) { return -1 * lambda / log(theta); }

/* NegativeBinomial_getComponentProbs, hmm_utils.c:494-515.  Returns 1 if a pmf was NaN (the reference exits). */
static int nb_comp_probs(const hfg_region_params *p, int s, int ncomp, uint8_t x, double *probs) {
    int nan = 0;
    for (int c = 0; c < ncomp; c++) {
        double theta = p->mean[s][c];
        double r = nb_r(theta, p->var[s][c]);
        double w = p->weight[s][c];
        probs[c] = w * exp(lgamma(r + x) - lgamma(r) - lgamma(x + 1) + r * log(theta) + (double) x * log(1 - theta));
        if (probs[c] != probs[c]) nan = 1;
        if (probs[c] < 1e-40) probs[c] = 1e-40;
    }
    return nan;
}

/* NegativeBinomial_fillDigammaTable, hmm_utils.c:392-406: table[x] = digamma(r + x), x = 0..MAX_COVERAGE_VALUE */
static void nb_digamma_table(double theta, double lambda, double *table) {
    double r = nb_r(theta, lambda);
    table[0] = (double) orc_digammal(r);
    for (int x = 1; x <= ORC_MAX_COV; x++) table[x] = table[x - 1] + 1.0 / (r + x - 1);
}

/* NegativeBinomial_updateEstimator, hmm_utils.c:536-563, for one (state, x) cell of the count histogram.
 * theta statistics go to mean_*, lambda statistics to var_* (hmm_oracle.h). */
static int nb_update_estimator(const hfg_region_params *p, int s, int ncomp, uint8_t x, double count,
                               hfg_region_stats *st, const double *dig /* [ncomp][ORC_MAX_COV + 1] */) {
    double probs[HFG_MAX_COMPS];
    int nan = nb_comp_probs(p, s, ncomp, x, probs);
    double tot = 0.0;
    for (int c = 0; c < ncomp; c++) tot += probs[c];
    for (int c = 0; c < ncomp; c++) {
        double theta = p->mean[s][c];
        double r = nb_r(theta, p->var[s][c]);
        double bt = -1 * theta / (1 - theta) - 1 / log(theta);
        double w = count * probs[c] / tot;
        const double *table = dig + (size_t) c * (ORC_MAX_COV + 1);
        double delta = r * (table[x] - table[0]);
        st->var_num[s][c] += w * delta; /* lambdaEstimator */
        st->var_den[s][c] += w;
        st->mean_num[s][c] += w * delta * bt; /* thetaEstimator */
        st->mean_den[s][c] += w * delta * bt + w * (x - delta);
        st->weight_num[s][c] += w; /* ParameterEstimator_incrementDenominatorForAllComps, hmm_utils.c:66-74 */
        for (int k = 0; k < ncomp; k++) st->weight_den[s][k] += w;
    }
    return nan;
}

/* Gaussian_getComponentProbs, hmm_utils.c:768-793.  Returns 1 if a pdf was NaN (the reference exits). */
static int gaussian_comp_probs(const hfg_region_params *p, int s, int ncomp, uint8_t x, uint8_t preX,
                               double alpha, double beta, double *probs) {
    int nan = 0;
    for (int c = 0; c < ncomp; c++) {
        double mean = (1 - alpha) * p->mean[s][c] + alpha * preX;
        mean *= beta;
        double var = p->var[s][c];
        var *= beta;
        double w = p->weight[s][c];
        double d = x - mean;
        probs[c] = w / (sqrt(var * 2 * ORC_PI)) * exp(-0.5 * (d * d) / var); /* pow(d,2) == d*d */
        if (probs[c] != probs[c]) nan = 1;
        if (probs[c] < 1e-40) probs[c] = 1e-40;
    }
    return nan;
}

/* TruncExponential_getProb, hmm_utils.c:941-947 */
static double trunc_exp_prob(const hfg_region_params *p, uint8_t x, double beta) {
    double lam = p->lambda / beta;
    double b = beta * p->trunc_point;
    if (p->trunc_point < x) return 0.0;
    return lam * exp(-lam * x) / (1 - exp(-lam * b));
}

/* EmissionDist_getProb, hmm_utils.c:1409-1417 (+ Gaussian_getProb :753-758: sum in component order) */
static double emission(const hfg_config *cfg, const hfg_region_params *p, int s, uint8_t x, uint8_t preX,
                       double alpha, double beta, int *nan) {
    if (!state_is_gaussian(cfg, s)) return trunc_exp_prob(p, x, beta);
    double probs[HFG_MAX_COMPS];
    if (is_nb(cfg)) /* NegativeBinomial_getProb, hmm_utils.c:479-484: no dependence on preX, alpha or beta */
        *nan |= nb_comp_probs(p, s, cfg->n_comps[s], x, probs);
    else
        *nan |= gaussian_comp_probs(p, s, cfg->n_comps[s], x, preX, alpha, beta, probs);
    double tot = 0.0;
    for (int c = 0; c < cfg->n_comps[s]; c++) tot += probs[c];
    return tot;
}

/* ValidityFunction_check{DupByMapq,ColByMapq,MsjByClipping}, hmm_utils.c:2229-2254; state index 4 is the END column
 * (tested through the STATE_MSJ == 4 rule because excludeMisjoin shifts the matrix, hmm_utils.c:2278-2292). */
static int state_valid(const hfg_config *cfg, int s, uint16_t cov, uint16_t mapq, uint16_t clip) {
    double highMapqRatio = (double) mapq / (0.1 + cov);
    double clipRatio = (double) clip / (0.1 + cov);
    if (s == HFG_STATE_DUP && highMapqRatio > cfg->max_high_mapq_ratio) return 0;
    if (s == HFG_STATE_COL && highMapqRatio < cfg->min_high_mapq_ratio) return 0;
    if (s == 4 && clipRatio < cfg->min_highly_clipped_ratio) return 0;
    return 1;
}

/* Transition_getProbConditional, hmm_utils.c:2278-2292 */
static double trans_cond(const hfg_config *cfg, const hfg_region_params *p, int pre, int s, uint16_t cov,
                         uint16_t mapq, uint16_t clip) {
    double tot = 0.0;
    for (int k = 0; k < NS + 1; k++)
        if (state_valid(cfg, k, cov, mapq, clip)) tot += p->trans[pre][k];
    double prob = p->trans[pre][s];
    return state_valid(cfg, s, cov, mapq, clip) ? prob / tot : 0.0;
}

/* One chunk: EM_runForward (hmm.c:423-434), EM_runBackward (:535-545), EM_updateEstimators (:638-650),
 * label loop (:730-736).  f,b: L x 4; scales: L.  Returns hfg_status. */
static int run_chunk(const hfg_config *cfg, const hfg_chunk_desc *ch, const uint16_t *cov, const uint16_t *mapq,
                     const uint16_t *clip, const uint8_t *region, const double *alpha,
                     const hfg_region_params *params, hfg_region_stats *stats, double *loglik_out, int8_t *labels,
                     double *post, double *f, double *b, double *scales, int forward_only,
                     double *nb_counts /* [R][NS][ORC_MAX_COV], zeroed; NULL unless negative binomial */,
                     const double *nb_dig /* [R][NS][HFG_MAX_COMPS][ORC_MAX_COV + 1] */) {
    const int L = ch->n_windows;
    int nan = 0;
    double loglik = 0.0;
    /* ---- forward ---- */
    for (int i = 0; i < L; i++) {
        double beta = orc_beta(cfg, ch, i);
        int r = region[i];
        uint8_t x = (uint8_t) cov[i];
        double scale = 0.0;
        if (i == 0) { /* EM_fillFirstColumnForward, hmm.c:333-364: preX = 0, alpha = 0, start prob, no validity mask */
            for (int s = 0; s < NS; s++) {
                double eProb = emission(cfg, &params[r], s, x, 0, 0.0, beta, &nan);
                double tProb = params[r].trans[NS][s];
                f[s] = eProb * tProb;
                scale += f[s];
            }
        } else { /* EM_fillOneColumnForward, hmm.c:366-420 */
            int preR = region[i - 1];
            uint8_t preX = (uint8_t) cov[i - 1];
            for (int s = 0; s < NS; s++) {
                double acc = 0.0;
                for (int pre = 0; pre < NS; pre++) {
                    double a = alpha[pre * NS + s];
                    double eProb = emission(cfg, &params[r], s, x, preX, a, beta, &nan);
                    double tProb = (r != preR) ? 1.0 / (NS + 1)
                                               : trans_cond(cfg, &params[r], pre, s, cov[i], mapq[i], clip[i]);
                    acc += (f[(i - 1) * NS + pre] * tProb * eProb);
                }
                f[i * NS + s] = acc;
                scale += acc;
            }
            if (scale < 1e-50) return HFG_ERR_SCALE_UNDERFLOW;
        }
        scales[i] = scale;
        for (int s = 0; s < NS; s++) f[i * NS + s] /= scale;
        loglik += log(scale);
    }
    *loglik_out = loglik;
    if (nan) return HFG_ERR_NAN;
    if (forward_only) return HFG_OK;

    /* ---- backward ---- */
    for (int s = 0; s < NS; s++) { /* EM_fillLastColumnBackward, hmm.c:452-467 */
        int r = region[L - 1];
        b[(L - 1) * NS + s] = params[r].trans[s][NS];
    }
    for (int s = 0; s < NS; s++) b[(L - 1) * NS + s] /= scales[L - 1];
    for (int i = L - 2; i >= 0; i--) { /* EM_fillOneColumnBackward, hmm.c:470-529 */
        double beta = orc_beta(cfg, ch, i + 1);
        int r = region[i + 1], preR = region[i];
        uint8_t x = (uint8_t) cov[i + 1], preX = (uint8_t) cov[i];
        double acc[NS] = {0.0, 0.0, 0.0, 0.0};
        for (int s = 0; s < NS; s++) {
            for (int pre = 0; pre < NS; pre++) {
                double a = alpha[pre * NS + s];
                double eProb = emission(cfg, &params[r], s, x, preX, a, beta, &nan);
                double tProb = (r != preR) ? 1.0 / (NS + 1)
                                           : trans_cond(cfg, &params[r], pre, s, cov[i + 1], mapq[i + 1], clip[i + 1]);
                acc[pre] += tProb * eProb * b[(i + 1) * NS + s];
            }
        }
        if (scales[i] < 1e-50) return HFG_ERR_SCALE_UNDERFLOW;
        for (int s = 0; s < NS; s++) b[i * NS + s] = acc[s] / scales[i];
    }

    /* ---- statistics: pairs i -> i+1 for i = 1 .. L-2 (EM_updateEstimators, hmm.c:638-650; :563-636) ---- */
    for (int i = 1; i < L - 1; i++) {
        double beta = orc_beta(cfg, ch, i + 1);
        int r = region[i + 1], preR = region[i];
        uint8_t x = (uint8_t) cov[i + 1], preX = (uint8_t) cov[i];
        const hfg_region_params *p = &params[r];
        hfg_region_stats *st = &stats[r];
        for (int s = 0; s < NS; s++) {
            for (int pre = 0; pre < NS; pre++) {
                double a = alpha[pre * NS + s];
                double eProb = emission(cfg, p, s, x, preX, a, beta, &nan);
                double tProb = (r != preR) ? 1.0 / (NS + 1)
                                           : trans_cond(cfg, p, pre, s, cov[i + 1], mapq[i + 1], clip[i + 1]);
                double count = f[i * NS + pre] * tProb * eProb * b[(i + 1) * NS + s];
                double adj = count / ORC_TERM;
                if (is_nb(cfg)) { /* hmm.c:615-617 -> CountData_increment, count_data.c:56-64: histogram of x per state,
                                     values past the last bin (x = 250) land in bin 249 */
                    int bin = x < ORC_MAX_COV ? x : ORC_MAX_COV - 1;
                    nb_counts[((size_t) r * NS + s) * ORC_MAX_COV + bin] += adj;
                } else if (!state_is_gaussian(cfg, s)) { /* TruncExponential_updateEstimator, hmm_utils.c:1027-1034 */
                    st->lambda_num += adj * x;
                    st->lambda_den += adj;
                } else { /* Gaussian_updateEstimator, hmm_utils.c:812-839 */
                    int nc = cfg->n_comps[s];
                    double x_adj = (x - a * preX) / (1.0 - a);
                    double probs[HFG_MAX_COMPS];
                    nan |= gaussian_comp_probs(p, s, nc, x, preX, a, beta, probs);
                    double tot = 0.0;
                    for (int c = 0; c < nc; c++) tot += probs[c];
                    for (int c = 0; c < nc; c++) {
                        double w = adj * probs[c] / tot;
                        st->mean_num[s][c] += w * x_adj;
                        st->mean_den[s][c] += w;
                        double z = (x_adj - p->mean[s][c]) * (1.0 - a);
                        st->var_num[s][c] += w * z * z;
                        st->var_den[s][c] += w;
                        st->weight_num[s][c] += w; /* ParameterEstimator_incrementDenominatorForAllComps, hmm_utils.c:66-74 */
                        for (int k = 0; k < nc; k++) st->weight_den[s][k] += w;
                    }
                }
                st->trans_count[pre][s] += adj; /* TransitionCountData_increment, hmm_utils.c:2010-2015 */
            }
        }
    }
    if (is_nb(cfg)) { /* EM_updateEstimators tail, hmm.c:643-649 -> EmissionDistSeries_updateAllEstimatorsUsingCountData,
                         hmm_utils.c:1662-1672: region, state, x ascending, cells with a positive count only */
        for (int r = 0; r < cfg->n_regions; r++)
            for (int s = 0; s < NS; s++)
                for (int x = 0; x < ORC_MAX_COV; x++) {
                    double count = nb_counts[((size_t) r * NS + s) * ORC_MAX_COV + x];
                    if (0 < count)
                        nan |= nb_update_estimator(&params[r], s, cfg->n_comps[s], (uint8_t) x, count, &stats[r],
                                                   nb_dig + ((size_t) r * NS + s) * HFG_MAX_COMPS * (ORC_MAX_COV + 1));
                }
    }
    if (nan) return HFG_ERR_NAN;

    /* ---- decode: EM_getPosterior / EM_getMostProbableState (hmm.c:671-692), first max (common.c:292-303) ---- */
    for (int i = 0; i < L; i++) {
        double g[NS], total = 0.0;
        for (int s = 0; s < NS; s++) {
            g[s] = f[i * NS + s] * b[i * NS + s] * scales[i];
            total += g[s];
        }
        for (int s = 0; s < NS; s++) g[s] /= total;
        int best = 0;
        for (int s = 1; s < NS; s++)
            if (g[best] < g[s]) best = s;
        if (labels) labels[i] = (int8_t) best;
        if (post)
            for (int s = 0; s < NS; s++) post[i * NS + s] = g[s];
    }
    return HFG_OK;
}

/* Test hook: when set, orc_estep adds every chunk's negative-binomial count histogram ([R][4][ORC_MAX_COV], zeroed by the
 * caller) into it, in list order -- the input the product's hfg_nb_stats_from_histogram is checked with. */
static double *g_nb_histogram_out = NULL;
void orc_nb_histogram_out(double *out) { g_nb_histogram_out = out; }

/* EM_runOneIterationForList / EM_runForwardForList, hmm.c:739-816: chunks in list order, per-chunk private
 * accumulators merged in list order (EM_updateModelEstimators, hmm.c:548-560).  Optional outputs may be NULL. */
int orc_estep(const hfg_config *cfg, int n_chunks, const hfg_chunk_desc *chunks, const uint16_t *cov,
              const uint16_t *mapq, const uint16_t *clip, const uint8_t *region, const double *alpha,
              const hfg_region_params *params, hfg_region_stats *stats, double *loglik, double *chunk_logliks,
              int8_t *labels, double *posteriors, double *fwd, double *bwd, double *scales_out, int forward_only) {
    const int R = cfg->n_regions;
    double total = 0.0;
    int status = HFG_OK;
    if (stats) memset(stats, 0, (size_t) R * sizeof(hfg_region_stats));
    hfg_region_stats *priv = calloc((size_t) R, sizeof(hfg_region_stats));
    int maxL = 1;
    for (int c = 0; c < n_chunks; c++)
        if (chunks[c].n_windows > maxL) maxL = chunks[c].n_windows;
    double *f = malloc(sizeof(double) * NS * (size_t) maxL);
    double *b = malloc(sizeof(double) * NS * (size_t) maxL);
    double *sc = malloc(sizeof(double) * (size_t) maxL);
    double *nb_counts = NULL, *nb_dig = NULL;
    const size_t n_counts = (size_t) R * NS * ORC_MAX_COV;
    if (is_nb(cfg) && !forward_only) {
        nb_counts = malloc(sizeof(double) * n_counts);
        nb_dig = calloc((size_t) R * NS * HFG_MAX_COMPS * (ORC_MAX_COV + 1), sizeof(double));
        for (int r = 0; r < R; r++) /* every chunk's copy of the model refills the same table (hmm_utils.c:377-390) */
            for (int s = 0; s < NS; s++)
                for (int k = 0; k < cfg->n_comps[s]; k++)
                    nb_digamma_table(params[r].mean[s][k], params[r].var[s][k],
                                     nb_dig + (((size_t) r * NS + s) * HFG_MAX_COMPS + k) * (ORC_MAX_COV + 1));
    }
    for (int c = 0; c < n_chunks && status == HFG_OK; c++) {
        const hfg_chunk_desc *ch = &chunks[c];
        const int64_t o = ch->offset;
        double ll = 0.0;
        memset(priv, 0, (size_t) R * sizeof(hfg_region_stats));
        if (nb_counts) memset(nb_counts, 0, sizeof(double) * n_counts);
        status = run_chunk(cfg, ch, cov + o, mapq + o, clip + o, region + o, alpha, params, priv, &ll,
                           labels ? labels + o : NULL, posteriors ? posteriors + o * NS : NULL, f, b, sc,
                           forward_only, nb_counts, nb_dig);
        if (status != HFG_OK) break;
        if (nb_counts && g_nb_histogram_out)
            for (size_t k = 0; k < n_counts; k++) g_nb_histogram_out[k] += nb_counts[k];
        total += ll;
        if (chunk_logliks) chunk_logliks[c] = ll;
        if (fwd) memcpy(fwd + o * NS, f, sizeof(double) * NS * (size_t) ch->n_windows);
        if (scales_out) memcpy(scales_out + o, sc, sizeof(double) * (size_t) ch->n_windows);
        if (bwd && !forward_only) memcpy(bwd + o * NS, b, sizeof(double) * NS * (size_t) ch->n_windows);
        if (stats && !forward_only) {
            const double *src = (const double *) priv;
            double *dst = (double *) stats;
            size_t n = (size_t) R * sizeof(hfg_region_stats) / sizeof(double);
            for (size_t k = 0; k < n; k++) dst[k] += src[k];
        }
    }
    free(priv);
    free(f);
    free(b);
    free(sc);
    free(nb_counts);
    free(nb_dig);
    if (loglik) *loglik = total;
    return status;
}

/* src/hmm_flagger.c:105-111,1008-1013 */
int orc_best_num_collapsed_comps(int max_coverage, const int32_t *region_coverages, int n_regions) {
    int mn = region_coverages[0];
    for (int i = 1; i < n_regions; i++)
        if (region_coverages[i] < mn) mn = region_coverages[i];
    int k = max_coverage / mn + 1;
    if (k < 2) k = 2;
    if (k > 10) k = 10;
    return k;
}

/* createModel (src/hmm_flagger.c:164-237) + HMM_construct (hmm.c:22-77) + EmissionDistSeries_constructForModel
 * (hmm_utils.c:1605-1652) + Transition_constructSymmetricBiased (hmm_utils.c:2109-2128); initialRandomDev = 0
 * (getRandomNumber(1,1) evaluates to 1.0). */
int orc_model_init(const hfg_config *cfg, const int32_t *region_coverages, int window_len, int start_only_mode,
                   hfg_region_params *params) {
    double medianCoverage = region_coverages[0];
    if (start_only_mode) medianCoverage *= (double) window_len / cfg->mean_read_length;
    double means[NS][HFG_MAX_COMPS];
    memset(means, 0, sizeof(means));
    means[HFG_STATE_ERR][0] = medianCoverage * ORC_ERR_COEF * 1.0;
    means[HFG_STATE_DUP][0] = medianCoverage * 0.5 * 1.0;
    means[HFG_STATE_HAP][0] = medianCoverage * 1.0 * 1.0;
    for (int i = 0; i < cfg->n_comps[HFG_STATE_COL]; i++)
        means[HFG_STATE_COL][i] = means[HFG_STATE_HAP][0] * (i + 2) * 1.0;
    for (int r = 0; r < cfg->n_regions; r++) {
        hfg_region_params *p = &params[r];
        memset(p, 0, sizeof(*p));
        double scale = (double) region_coverages[r] / medianCoverage;
        for (int s = 0; s < NS; s++) {
            if (!state_is_gaussian(cfg, s)) continue; /* Err is a TruncExponential in the default model */
            for (int c = 0; c < cfg->n_comps[s]; c++) {
                p->mean[s][c] = means[s][c] * scale;   /* Double_multiply2DArray, hmm.c:43-47 */
                p->var[s][c] = p->mean[s][c] * 1.0;    /* Gaussian_constructByMean(mean, 1.0, n), hmm_utils.c:733-741 */
                p->weight[s][c] = 1.0 / cfg->n_comps[s]; /* hmm_utils.c:667 */
                if (is_nb(cfg)) { /* NegativeBinomial_constructByMean(mean, 1.5, n), hmm_utils.c:335-342,433-453,1635-1639 */
                    double mean = p->mean[s][c], var = mean * 1.5;
                    double theta = mean / var;
                    double r = pow(mean, 2) / (var - mean);
                    p->mean[s][c] = theta;
                    p->var[s][c] = -1 * r * log(theta);
                }
            }
        }
        if (cfg->model_type == HFG_MODEL_TRUNC_EXP_GAUSSIAN) {
            p->lambda = 1.0;                                        /* hmm_utils.c:1619 */
            p->trunc_point = p->mean[HFG_STATE_HAP][0] * ORC_TRUNC_FRACTION;
        }
        double off = (1.0 - ORC_DIAG) / (NS - 1) * (1.0 - ORC_TERM);
        double diag = ORC_DIAG * (1.0 - ORC_TERM);
        for (int i = 0; i < NS + 1; i++)
            for (int j = 0; j < NS + 1; j++) p->trans[i][j] = (i == j) ? diag : off;
        for (int s = 0; s < NS; s++) {
            p->trans[NS][s] = 1.0 / NS;
            p->trans[s][NS] = ORC_TERM;
        }
        p->trans[NS][NS] = 0.0;
    }
    return HFG_OK;
}

/* TruncExponential_getLogLikelihoodByParams, hmm_utils.c:949-956 */
static double trunc_ll(double lam, double b, double num, double denom) {
    return denom * log(lam) - denom * log(1.0 - exp(-lam * b)) - num * lam;
}

/* TruncExponential_estimateLambda, hmm_utils.c:969-1011 (golden-section search) */
static double estimate_lambda(double truncPoint, double num, double denom, double tol) {
    double a = 0.0, b = truncPoint;
    double invphi = (sqrt(5.0) - 1.0) / 2.0;
    double invphi2 = (3.0 - sqrt(5.0)) / 2.0;
    double h = b - a;
    if (h <= tol) return (b + a) / 2.0;
    int n = (int) ceil(log(tol / h) / log(invphi));
    double c = a + invphi2 * h;
    double d = a + invphi * h;
    double yc = trunc_ll(c, truncPoint, num, denom);
    double yd = trunc_ll(d, truncPoint, num, denom);
    for (int k = 0; k < n - 1; k++) {
        if (yc > yd) {
            b = d; d = c; yd = yc;
            h = invphi * h;
            c = a + invphi2 * h;
            yc = trunc_ll(c, truncPoint, num, denom);
        } else {
            a = c; c = d; yc = yd;
            h = invphi * h;
            d = a + invphi * h;
            yd = trunc_ll(d, truncPoint, num, denom);
        }
    }
    if (yc > yd) return (a + d) / 2.0;
    return (c + b) / 2.0;
}

/* binding coefficients, hmm_utils.c:191-238 (Gaussian) ; weight and lambda are unbound (coef 0) */
static double bind_coef(int s, int c) {
    switch (s) {
        case HFG_STATE_ERR: return ORC_ERR_COEF;
        case HFG_STATE_DUP: return 0.5;
        case HFG_STATE_HAP: return 1.0;
        default: return 2.0 + 1.0 * c; /* ParameterBinding_constructSequenceByStep */
    }
}

/* type 0 = Gaussian mean / NB theta, type 1 = Gaussian var / NB lambda.  The NB theta is bound with coefficient 1 in every
 * state and component (ParameterBinding_getDefault1DArrayForNegativeBinomial, hmm_utils.c:240-290); everything else uses
 * the 0.1 / 0.5 / 1 / 2+c ladder. */
static double bind_factor(const hfg_config *cfg, int type, int s, int c) {
    if (is_nb(cfg) && type == 0) return 1.0;
    return bind_coef(s, c);
}

/* Gaussian_updateParameter / TruncExponential_updateParameter convergence rule, hmm_utils.c:855-858,1051-1053 */
static int conv_emission(double oldv, double newv, double tol) {
    double diffRatio = 1.0e-4 < oldv ? fabs(newv / oldv - 1.0) : 0.0;
    return diffRatio < tol;
}

/* HMM_estimateParameters, hmm.c:120-127 -> EmissionDistSeries_estimateParameters (hmm_utils.c:1860-1903)
 * and Transition_estimateTransitionMatrix (hmm_utils.c:2185-2219). */
int orc_mstep(const hfg_config *cfg, hfg_region_params *params, const hfg_region_stats *stats, double tol,
              int *converged_out) {
    int converged = 1;
    for (int r = 0; r < cfg->n_regions; r++) {
        hfg_region_params *p = &params[r];
        const hfg_region_stats *st = &stats[r];
        /* MEAN then VAR: bound estimator pooled over all Gaussian states (hmm_utils.c:1791-1858) */
        for (int type = 0; type < 2; type++) {
            double bnum = 0.0, bden = 0.0;
            for (int s = 0; s < NS; s++) {
                if (!state_is_gaussian(cfg, s)) continue;
                for (int c = 0; c < cfg->n_comps[s]; c++) {
                    double factor = bind_factor(cfg, type, s, c);
                    double num = type == 0 ? st->mean_num[s][c] : st->var_num[s][c];
                    double den = type == 0 ? st->mean_den[s][c] : st->var_den[s][c];
                    bnum += num / factor;
                    bden += den;
                }
            }
            double est = bden == 0 ? 0.0 : bnum / bden; /* ParameterEstimator_getEstimation, hmm_utils.c:76-92 */
            for (int s = 0; s < NS; s++) {
                if (!state_is_gaussian(cfg, s)) continue;
                for (int c = 0; c < cfg->n_comps[s]; c++) {
                    double value = est * bind_factor(cfg, type, s, c);
                    if (ORC_MIN_COUNT < bden) {
                        double *dst = type == 0 ? &p->mean[s][c] : &p->var[s][c];
                        converged &= conv_emission(*dst, value, tol);
                        *dst = value;
                    }
                }
            }
        }
        /* WEIGHT: unbound, own estimator */
        for (int s = 0; s < NS; s++) {
            if (!state_is_gaussian(cfg, s)) continue;
            for (int c = 0; c < cfg->n_comps[s]; c++) {
                double den = st->weight_den[s][c];
                double value = den == 0 ? 0.0 : st->weight_num[s][c] / den;
                if (ORC_MIN_COUNT < den) {
                    converged &= conv_emission(p->weight[s][c], value, tol);
                    p->weight[s][c] = value;
                }
            }
        }
        if (cfg->model_type == HFG_MODEL_TRUNC_EXP_GAUSSIAN) {
            /* lambda: golden-section on the CURRENT truncPoint (hmm_utils.c:76-92,969-1011) */
            double den = st->lambda_den;
            if (den != 0 && ORC_MIN_COUNT < den) {
                double value = estimate_lambda(p->trunc_point, st->lambda_num, den, 1e-6);
                converged &= conv_emission(p->lambda, value, tol);
                p->lambda = value;
            }
            p->trunc_point = p->mean[HFG_STATE_HAP][0] * ORC_TRUNC_FRACTION; /* hmm_utils.c:1878-1882 */
        }
        /* transitions, hmm_utils.c:2185-2219 */
        for (int i1 = 0; i1 < NS; i1++) {
            double rowSum = 0.0;
            for (int i2 = 0; i2 < NS; i2++) rowSum += st->trans_count[i1][i2] + ORC_PSEUDO;
            for (int i2 = 0; i2 < NS; i2++) {
                double oldValue = p->trans[i1][i2];
                double newValue = (st->trans_count[i1][i2] + ORC_PSEUDO) / rowSum * (1.0 - ORC_TERM);
                p->trans[i1][i2] = newValue;
                double diffRatio = 1.0e-6 < oldValue ? fabs(newValue / oldValue - 1.0) : 0.0;
                converged &= diffRatio < tol;
            }
        }
        for (int i1 = 0; i1 < NS; i1++) p->trans[i1][NS] = ORC_TERM;
        for (int i2 = 0; i2 < NS; i2++) p->trans[NS][i2] = 1.0 / NS;
        p->trans[NS][NS] = 0.0;
    }
    *converged_out = converged;
    return HFG_OK;
}

/* runHMMFlagger EM loop without file outputs, src/hmm_flagger.c:337-467 */
int orc_run_em(const hfg_config *cfg, int n_chunks, const hfg_chunk_desc *chunks, const uint16_t *cov,
               const uint16_t *mapq, const uint16_t *clip, const uint8_t *region, const double *alpha,
               hfg_region_params *params, int max_iterations, double tol, double *logliks, int *n_esteps,
               int8_t *labels) {
    hfg_region_stats *stats = calloc((size_t) cfg->n_regions, sizeof(hfg_region_stats));
    int iter = 1, converged = 0, k = 0, status = HFG_OK;
    while (iter <= max_iterations && !converged) {
        status = orc_estep(cfg, n_chunks, chunks, cov, mapq, clip, region, alpha, params, stats, &logliks[k], NULL,
                           NULL, NULL, NULL, NULL, NULL, 0);
        if (status != HFG_OK) goto done;
        k++;
        orc_mstep(cfg, params, stats, tol, &converged);
        iter++;
    }
    status = orc_estep(cfg, n_chunks, chunks, cov, mapq, clip, region, alpha, params, stats, &logliks[k], NULL, labels,
                       NULL, NULL, NULL, NULL, 0);
    if (status == HFG_OK) k++;
done:
    *n_esteps = k;
    free(stats);
    return status;
}

/* ---- SQUAREM (--accelerate) restated: SquareAccelerator, hmm.c:820-1098 ---------------------------------------- */

/* The accelerated parameters of one region copied into / out of a vector, in the order the reference visits them
 * (EmissionDistSeriesParamIter_next, hmm_utils.c:1218-1245 over EmissionDistParamIter_next :1151-1195): per state, a
 * truncated exponential contributes its rate (TRUNC_EXP_LAMBDA, the only type the iterator reaches), a Gaussian
 * contributes mean, var, weight of comp 0, then of comp 1, ...; then rows 0..3 x columns 0..3 of the transition matrix
 * (hmm.c:977-992,1072-1090). */
static int accel_gather(const hfg_config *cfg, const hfg_region_params *p, double *vec) {
    int n = 0;
    for (int s = 0; s < NS; s++) {
        if (!state_is_gaussian(cfg, s)) {
            vec[n++] = p->lambda;
            continue;
        }
        for (int c = 0; c < cfg->n_comps[s]; c++) {
            vec[n++] = p->mean[s][c];
            vec[n++] = p->var[s][c];
            vec[n++] = p->weight[s][c];
        }
    }
    for (int i = 0; i < NS; i++)
        for (int j = 0; j < NS; j++) vec[n++] = p->trans[i][j];
    return n;
}

static void accel_scatter(const hfg_config *cfg, const double *vec, hfg_region_params *p) {
    int n = 0;
    for (int s = 0; s < NS; s++) {
        if (!state_is_gaussian(cfg, s)) {
            p->lambda = vec[n++];
            continue;
        }
        for (int c = 0; c < cfg->n_comps[s]; c++) {
            p->mean[s][c] = vec[n++];
            p->var[s][c] = vec[n++];
            p->weight[s][c] = vec[n++];
        }
    }
    for (int i = 0; i < NS; i++)
        for (int j = 0; j < NS; j++) p->trans[i][j] = vec[n++];
}

#define ACCEL_MAX (NS * HFG_MAX_COMPS * 3 + NS * NS)

int orc_feasible(const hfg_config *cfg, const hfg_region_params *params) { /* HMM_isFeasible, hmm.c:80-87 */
    int feasible = 1;
    for (int r = 0; r < cfg->n_regions; r++) {
        const hfg_region_params *p = &params[r];
        for (int s = 0; s < NS; s++) {
            if (!state_is_gaussian(cfg, s)) { /* hmm_utils.c:920-925 */
                if (!(0 < p->lambda)) feasible = 0;
                if (!(0 < p->trunc_point)) feasible = 0;
                continue;
            }
            for (int c = 0; c < cfg->n_comps[s]; c++) { /* hmm_utils.c:685-694; NB: hmm_utils.c:366-375 */
                if (is_nb(cfg) && !(p->mean[s][c] < 1)) feasible = 0;
                if (!(0 < p->mean[s][c])) feasible = 0;
                if (!(0 < p->var[s][c])) feasible = 0;
                if (!((0 <= p->weight[s][c]) && (p->weight[s][c] <= 1))) feasible = 0;
            }
        }
        for (int i = 0; i < NS; i++) /* hmm_utils.c:2130-2139 */
            for (int j = 0; j < NS; j++)
                if (p->trans[i][j] < 0 || 1 < p->trans[i][j]) feasible = 0;
    }
    return feasible;
}

static void accel_candidate(const hfg_config *cfg, const hfg_region_params *p0, const hfg_region_params *p1,
                            const hfg_region_params *p2, double rate, hfg_region_params *prime) {
    double v0[ACCEL_MAX], v1[ACCEL_MAX], v2[ACCEL_MAX], vp[ACCEL_MAX];
    for (int reg = 0; reg < cfg->n_regions; reg++) {
        const int n = accel_gather(cfg, &p0[reg], v0);
        accel_gather(cfg, &p1[reg], v1);
        accel_gather(cfg, &p2[reg], v2);
        for (int i = 0; i < n; i++) {
            const double r = v1[i] - v0[i];
            const double v = v2[i] - v1[i] - r;
            vp[i] = v0[i] - 2 * r * rate + v * pow(rate, 2); /* hmm.c:960 */
        }
        prime[reg] = p0[reg]; /* modelPrime starts as a copy of model 0 (hmm.c:853-858) */
        accel_scatter(cfg, vp, &prime[reg]);
        hfg_region_params *q = &prime[reg];
        for (int s = 0; s < NS; s++) { /* Gaussian_normalizeWeights, hmm_utils.c:675-683 */
            if (!state_is_gaussian(cfg, s)) continue;
            double sum = 0.0;
            for (int c = 0; c < cfg->n_comps[s]; c++) sum += q->weight[s][c];
            for (int c = 0; c < cfg->n_comps[s]; c++) q->weight[s][c] *= 1.0 / sum;
        }
        for (int i = 0; i < NS; i++) { /* Transition_normalizeTransitionRows, hmm_utils.c:2165-2183 */
            double rowSum = 0.0;
            for (int j = 0; j < NS; j++) rowSum += q->trans[i][j];
            for (int j = 0; j < NS; j++) q->trans[i][j] = q->trans[i][j] / rowSum * (1.0 - ORC_TERM);
        }
        for (int i = 0; i < NS; i++) q->trans[i][NS] = ORC_TERM;
        q->trans[NS][NS] = 0.0;
    }
}

static double accel_rate(const hfg_config *cfg, const hfg_region_params *p0, const hfg_region_params *p1,
                         const hfg_region_params *p2) { /* SquareAccelerator_computeRates, hmm.c:1000-1098 */
    double v0[ACCEL_MAX], v1[ACCEL_MAX], v2[ACCEL_MAX], num = 0.0, den = 0.0;
    for (int reg = 0; reg < cfg->n_regions; reg++) {
        const int n = accel_gather(cfg, &p0[reg], v0);
        accel_gather(cfg, &p1[reg], v1);
        accel_gather(cfg, &p2[reg], v2);
        for (int i = 0; i < n; i++) {
            const double r = v1[i] - v0[i];
            const double v = v2[i] - v1[i] - r;
            num += pow(r, 2);
            den += pow(v, 2);
        }
    }
    double rate = -1 * sqrt(num / den);
    if (rate > -1) rate = -1;
    return rate;
}

static void accel_shrink(const hfg_config *cfg, const hfg_region_params *p0, const hfg_region_params *p1,
                         const hfg_region_params *p2, double margin, double *rate, hfg_region_params *prime) {
    *rate = (*rate - 1) / 2; /* hmm.c:869-883 */
    if (*rate > (-1 - margin)) {
        *rate = -1.0;
        memcpy(prime, p0, sizeof(hfg_region_params) * (size_t) cfg->n_regions);
    } else {
        accel_candidate(cfg, p0, p1, p2, *rate, prime);
    }
}

/* computeRates + computeValuesForModelPrime + n_shrinks x shrinkAlphaAndRecomputeModelPrime; returns feasibility */
int orc_squarem(const hfg_config *cfg, const hfg_region_params *p0, const hfg_region_params *p1,
                const hfg_region_params *p2, int n_shrinks, double margin, hfg_region_params *prime, double *alpha_rate) {
    double rate = accel_rate(cfg, p0, p1, p2);
    accel_candidate(cfg, p0, p1, p2, rate, prime);
    for (int i = 0; i < n_shrinks; i++) accel_shrink(cfg, p0, p1, p2, margin, &rate, prime);
    *alpha_rate = rate;
    return orc_feasible(cfg, prime);
}

/* runHMMFlagger with acceleration == true (src/hmm_flagger.c:337-467) + SquareAccelerator_getModelPrime (hmm.c:885-918) */
int orc_run_em_accelerated(const hfg_config *cfg, int n_chunks, const hfg_chunk_desc *chunks, const uint16_t *cov,
                           const uint16_t *mapq, const uint16_t *clip, const uint8_t *region, const double *alpha,
                           hfg_region_params *params, int max_iterations, double tol, double *logliks,
                           double *alpha_rates, int *n_outer, int8_t *labels) {
    const int R = cfg->n_regions;
    hfg_region_stats *stats = calloc((size_t) R, sizeof(hfg_region_stats));
    hfg_region_params *p0 = calloc((size_t) R * 4, sizeof(hfg_region_params)), *p1 = p0 + R, *p2 = p1 + R, *pp = p2 + R;
    int iter = 1, converged = 0, k = 0, status = HFG_OK, ignored;
#define ESTEP(P, LL, FWD) orc_estep(cfg, n_chunks, chunks, cov, mapq, clip, region, alpha, (P), stats, (LL), NULL, NULL, \
                                    NULL, NULL, NULL, NULL, (FWD))
    while (iter <= max_iterations && !converged) {
        double ll0, llp;
        if ((status = ESTEP(params, &ll0, 0)) != HFG_OK) goto done;
        logliks[k] = ll0;
        memcpy(p0, params, sizeof(hfg_region_params) * (size_t) R);
        memcpy(p1, params, sizeof(hfg_region_params) * (size_t) R);
        orc_mstep(cfg, p1, stats, tol, &ignored);
        if ((status = ESTEP(p1, &llp, 0)) != HFG_OK) goto done;
        memcpy(p2, p1, sizeof(hfg_region_params) * (size_t) R);
        orc_mstep(cfg, p2, stats, tol, &ignored);
        double rate = accel_rate(cfg, p0, p1, p2);
        accel_candidate(cfg, p0, p1, p2, rate, pp);
        while (!orc_feasible(cfg, pp)) accel_shrink(cfg, p0, p1, p2, 1e-2, &rate, pp);
        if ((status = ESTEP(pp, &llp, 1)) != HFG_OK) goto done;
        while (llp < ll0) {
            accel_shrink(cfg, p0, p1, p2, 1e-2, &rate, pp);
            while (!orc_feasible(cfg, pp)) accel_shrink(cfg, p0, p1, p2, 1e-2, &rate, pp);
            if ((status = ESTEP(pp, &llp, 1)) != HFG_OK) goto done;
        }
        if (alpha_rates) alpha_rates[k] = rate;
        k++;
        if ((status = ESTEP(pp, &llp, 0)) != HFG_OK) goto done;
        memcpy(params, pp, sizeof(hfg_region_params) * (size_t) R);
        orc_mstep(cfg, params, stats, tol, &converged);
        iter++;
    }
    status = orc_estep(cfg, n_chunks, chunks, cov, mapq, clip, region, alpha, params, stats, &logliks[k], NULL, labels,
                       NULL, NULL, NULL, NULL, 0);
#undef ESTEP
done:
    *n_outer = k;
    free(stats);
    free(p0);
    return status;
}

/* persistent handle for benchmarking (mirrors ref_open/ref_step/ref_close of ref_harness.c): whole EM iterations
 * of the restatement, single-threaded */
typedef struct OrcRun {
    hfg_config cfg;
    int n_chunks;
    const hfg_chunk_desc *chunks;
    const uint16_t *cov, *mapq, *clip;
    const uint8_t *region;
    double alpha[16];
    hfg_region_params *params;
    hfg_region_stats *stats;
} OrcRun;

void *orc_open(const hfg_config *cfg, int n_chunks, const hfg_chunk_desc *cd, const uint16_t *cov,
               const uint16_t *mapq, const uint16_t *clip, const uint8_t *region, const double *alpha,
               const hfg_region_params *params) {
    OrcRun *r = calloc(1, sizeof(OrcRun));
    r->cfg = *cfg;
    r->n_chunks = n_chunks;
    r->chunks = cd;
    r->cov = cov;
    r->mapq = mapq;
    r->clip = clip;
    r->region = region;
    memcpy(r->alpha, alpha, sizeof(r->alpha));
    r->params = malloc(sizeof(hfg_region_params) * (size_t) cfg->n_regions);
    memcpy(r->params, params, sizeof(hfg_region_params) * (size_t) cfg->n_regions);
    r->stats = calloc((size_t) cfg->n_regions, sizeof(hfg_region_stats));
    return r;
}

#include <time.h>
double orc_step(void *handle, int threads, double tol, double *loglik) {
    (void) threads;
    OrcRun *r = handle;
    struct timespec t0, t1;
    int conv = 0;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    orc_estep(&r->cfg, r->n_chunks, r->chunks, r->cov, r->mapq, r->clip, r->region, r->alpha, r->params, r->stats,
              loglik, NULL, NULL, NULL, NULL, NULL, NULL, 0);
    orc_mstep(&r->cfg, r->params, r->stats, tol, &conv);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}

void orc_close(void *handle) {
    OrcRun *r = handle;
    free(r->params);
    free(r->stats);
    free(r);
}
