/*
 * hfg_host_model.c -- host mirrors (plain C) of the O(#parameters) functions on either side of the E-step seam:
 * initial model (createModel / HMM_construct) and the M-step (HMM_estimateParameters).  The reference keeps these
 * on the host and so do we; they exist here so that a whole EM run can be driven through include/hfg.h alone
 * (hfg_run_em, bench.py, the multi-GPU driver).  When libhfg is linked into the reference binary instead
 *  the reference's own versions are used and these are not called.
 */
#include <math.h>
#include <string.h>

#include "hfg_internal.h"

#define HFG_HD static inline
#include "hfg_mstep_inl.h"

#define INITIAL_DIAG_PROB 0.99     /* hmm.c:15 */

static int gaussian_state(const hfg_config *cfg, int s) {
    return !(cfg->model_type == HFG_MODEL_TRUNC_EXP_GAUSSIAN && s == HFG_STATE_ERR);
}

int hfg_best_num_collapsed_comps(int max_coverage, const int32_t *region_coverages, int n_regions) {
    /* getBestNumberOfCollapsedComps + clamp, src/hmm_flagger.c:105-111,1008-1013 */
    int lowest = region_coverages[0];
    for (int r = 1; r < n_regions; r++) lowest = region_coverages[r] < lowest ? region_coverages[r] : lowest;
    int k = max_coverage / lowest + 1;
    return k < 2 ? 2 : (k > 10 ? 10 : k);
}

int hfg_model_init(const hfg_config *cfg, const int32_t *region_coverages, int window_len, int start_only_mode,
                   hfg_region_params *params) {
    if (!cfg || !region_coverages || !params || cfg->n_regions < 1 || cfg->n_regions > HFG_MAX_REGIONS)
        return HFG_ERR_INVALID;
    for (int s = 0; s < HFG_NS; s++)
        if (cfg->n_comps[s] < 1 || cfg->n_comps[s] > HFG_MAX_COMPS) return HFG_ERR_INVALID;
    /* baseline = coverage of region 0 (src/hmm_flagger.c:186-195) */
    double baseline = region_coverages[0];
    if (start_only_mode) baseline *= (double) window_len / cfg->mean_read_length;
    const double hap = baseline * 1.0 * 1.0; /* "* getRandomNumber(1,1)" == 1.0 with --initialRandomDev 0 */
    const double off_diag = (1.0 - INITIAL_DIAG_PROB) / (HFG_NS - 1) * (1.0 - HFG_TERM_PROB);
    const double on_diag = INITIAL_DIAG_PROB * (1.0 - HFG_TERM_PROB);
    for (int r = 0; r < cfg->n_regions; r++) {
        hfg_region_params *p = &params[r];
        memset(p, 0, sizeof(*p));
        const double region_scale = (double) region_coverages[r] / baseline;
        for (int s = 0; s < HFG_NS; s++) {
            if (!gaussian_state(cfg, s)) continue;
            for (int c = 0; c < cfg->n_comps[s]; c++) {
                double m;
                if (s == HFG_STATE_ERR) m = baseline * 0.1 * 1.0;
                else if (s == HFG_STATE_DUP) m = baseline * 0.5 * 1.0;
                else if (s == HFG_STATE_HAP) m = hap;
                else m = hap * (c + 2) * 1.0;
                p->mean[s][c] = m * region_scale;          /* hmm.c:43-47 */
                p->var[s][c] = p->mean[s][c] * 1.0;        /* hmm_utils.c:733-741 with factor 1.0 (:1622,1630) */
                p->weight[s][c] = 1.0 / cfg->n_comps[s];   /* hmm_utils.c:667 */
                if (cfg->model_type == HFG_MODEL_NEGATIVE_BINOMIAL) /* same means, as (theta, lambda) (hfg_nb.c) */
                    hfg_nb_init_component(p->mean[s][c], &p->mean[s][c], &p->var[s][c]);
            }
        }
        if (cfg->model_type == HFG_MODEL_TRUNC_EXP_GAUSSIAN) {
            p->lambda = 1.0;                               /* hmm_utils.c:1619 */
            p->trunc_point = hap * region_scale * TRUNC_POINT_FRACTION;
        }
        for (int i = 0; i <= HFG_NS; i++)
            for (int j = 0; j <= HFG_NS; j++) p->trans[i][j] = i == j ? on_diag : off_diag; /* hmm_utils.c:2113-2116 */
        for (int s = 0; s < HFG_NS; s++) {
            p->trans[HFG_NS][s] = 1.0 / HFG_NS;
            p->trans[s][HFG_NS] = HFG_TERM_PROB;
        }
        p->trans[HFG_NS][HFG_NS] = 0.0;
    }
    return HFG_OK;
}

int hfg_mstep(const hfg_config *cfg, hfg_region_params *params, const hfg_region_stats *stats,
              double convergence_tol, int *converged) {
    if (!cfg || !params || !stats || !converged) return HFG_ERR_INVALID;
    int all_settled = 1;
    for (int r = 0; r < cfg->n_regions; r++)
        all_settled &= cfg->model_type == HFG_MODEL_NEGATIVE_BINOMIAL
                           ? hfg_nb_mstep_region(cfg->n_comps, &params[r], &stats[r], convergence_tol)
                           : hfg_mstep_region(cfg->model_type, cfg->n_comps, &params[r], &stats[r], convergence_tol);
    *converged = all_settled;
    return HFG_OK;
}

/* ---- SQUAREM (--accelerate): the host arithmetic of SquareAccelerator (hmm.c:820-1098) on flat parameters -------- */

/* HMM_isFeasible (hmm.c:80-87; hmm_utils.c:678-691,920-925,2130-2139).  The comparisons are written as the reference
 * writes them, so a NaN fails the emission checks but passes the transition check. */
int hfg_params_feasible(const hfg_config *cfg, const hfg_region_params *params) {
    int ok = 1;
    for (int r = 0; r < cfg->n_regions; r++) {
        const hfg_region_params *p = &params[r];
        for (int s = 0; s < HFG_NS; s++) {
            if (!gaussian_state(cfg, s)) {
                ok &= 0 < p->lambda;
                ok &= 0 < p->trunc_point;
            } else {
                for (int c = 0; c < cfg->n_comps[s]; c++) {
                    if (cfg->model_type == HFG_MODEL_NEGATIVE_BINOMIAL) ok &= p->mean[s][c] < 1; /* theta, hmm_utils.c:366-375 */
                    ok &= 0 < p->mean[s][c];
                    ok &= 0 < p->var[s][c];
                    ok &= (0 <= p->weight[s][c]) && (p->weight[s][c] <= 1);
                }
            }
        }
        for (int a = 0; a < HFG_NS; a++)
            for (int b = 0; b < HFG_NS; b++)
                if (p->trans[a][b] < 0 || 1 < p->trans[a][b]) ok = 0;
    }
    return ok;
}

/* the accelerated parameters in the reference's iteration order (EmissionDistSeriesParamIter, hmm_utils.c:1111-1245:
 * per state; truncated exponential = its rate only; Gaussian = component-major {mean, var, weight}), then the 4 x 4
 * transition block.  Returns the number of slots; slot[i] points into `p`. */
static int param_slots(const hfg_config *cfg, hfg_region_params *p, double **slot) {
    int n = 0;
    for (int s = 0; s < HFG_NS; s++) {
        if (!gaussian_state(cfg, s)) {
            slot[n++] = &p->lambda;
        } else {
            for (int c = 0; c < cfg->n_comps[s]; c++) {
                slot[n++] = &p->mean[s][c];
                slot[n++] = &p->var[s][c];
                slot[n++] = &p->weight[s][c];
            }
        }
    }
    for (int a = 0; a < HFG_NS; a++)
        for (int b = 0; b < HFG_NS; b++) slot[n++] = &p->trans[a][b];
    return n;
}

#define MAX_SLOTS (HFG_NS * HFG_MAX_COMPS * 3 + HFG_NS * HFG_NS)

/* SquareAccelerator_computeRates (hmm.c:1000-1098): alpha = -sqrt(sum r^2 / sum v^2), capped at -1, with
 * r = p1 - p0 and v = (p2 - p1) - r summed over all regions. */
double hfg_squarem_alpha_rate(const hfg_config *cfg, const hfg_region_params *p0, const hfg_region_params *p1,
                              const hfg_region_params *p2) {
    double num = 0.0, den = 0.0;
    double *s0[MAX_SLOTS], *s1[MAX_SLOTS], *s2[MAX_SLOTS];
    for (int reg = 0; reg < cfg->n_regions; reg++) {
        const int n = param_slots(cfg, (hfg_region_params *) &p0[reg], s0);
        param_slots(cfg, (hfg_region_params *) &p1[reg], s1);
        param_slots(cfg, (hfg_region_params *) &p2[reg], s2);
        for (int i = 0; i < n; i++) {
            const double r = *s1[i] - *s0[i];
            const double v = *s2[i] - *s1[i] - r;
            num += pow(r, 2);
            den += pow(v, 2);
        }
    }
    double rate = -1 * sqrt(num / den);
    if (rate > -1) rate = -1;
    return rate;
}

/* SquareAccelerator_computeValuesForModelPrime (hmm.c:921-997): p' = p0 - 2 r alpha + v alpha^2 per accelerated
 * parameter (everything else -- truncation point, start row -- stays p0's), then mixture weights and transition rows
 * renormalised (HMM_normalizeWeightsAndTransitionRows, hmm.c:89-94; hmm_utils.c:675-683,2165-2183). */
int hfg_squarem_prime(const hfg_config *cfg, const hfg_region_params *p0, const hfg_region_params *p1,
                      const hfg_region_params *p2, double alpha_rate, hfg_region_params *prime) {
    if (!cfg || !p0 || !p1 || !p2 || !prime) return HFG_ERR_INVALID;
    double *s0[MAX_SLOTS], *s1[MAX_SLOTS], *s2[MAX_SLOTS], *sp[MAX_SLOTS];
    for (int reg = 0; reg < cfg->n_regions; reg++) {
        hfg_region_params *q = &prime[reg];
        if (q != &p0[reg]) *q = p0[reg];
        const int n = param_slots(cfg, (hfg_region_params *) &p0[reg], s0);
        param_slots(cfg, (hfg_region_params *) &p1[reg], s1);
        param_slots(cfg, (hfg_region_params *) &p2[reg], s2);
        param_slots(cfg, q, sp);
        for (int i = 0; i < n; i++) {
            const double r = *s1[i] - *s0[i];
            const double v = *s2[i] - *s1[i] - r;
            *sp[i] = *s0[i] - 2 * r * alpha_rate + v * pow(alpha_rate, 2);
        }
        for (int s = 0; s < HFG_NS; s++) {
            if (!gaussian_state(cfg, s)) continue;
            double sum = 0.0;
            for (int c = 0; c < cfg->n_comps[s]; c++) sum += q->weight[s][c];
            if (!(0.0 < sum)) return HFG_ERR_INVALID; /* the reference exits: "Sum of weights is not > 0" */
            const double inv = 1.0 / sum;
            for (int c = 0; c < cfg->n_comps[s]; c++) q->weight[s][c] *= inv;
        }
        for (int a = 0; a < HFG_NS; a++) {
            double row = 0.0;
            for (int b = 0; b < HFG_NS; b++) row += q->trans[a][b];
            for (int b = 0; b < HFG_NS; b++) q->trans[a][b] = q->trans[a][b] / row * (1.0 - HFG_TERM_PROB);
        }
        for (int a = 0; a < HFG_NS; a++) q->trans[a][HFG_NS] = HFG_TERM_PROB;
        q->trans[HFG_NS][HFG_NS] = 0.0;
    }
    return HFG_OK;
}

/* SquareAccelerator_shrinkAlphaAndRecomputeModelPrime (hmm.c:869-883): halve the step towards -1; within `margin` of -1
 * the candidate becomes p0 itself. */
int hfg_squarem_shrink(const hfg_config *cfg, const hfg_region_params *p0, const hfg_region_params *p1,
                       const hfg_region_params *p2, double margin, double *alpha_rate, hfg_region_params *prime) {
    if (!alpha_rate) return HFG_ERR_INVALID;
    *alpha_rate = (*alpha_rate - 1) / 2;
    if (*alpha_rate > (-1 - margin)) {
        *alpha_rate = -1.0;
        memcpy(prime, p0, sizeof(hfg_region_params) * (size_t) cfg->n_regions);
        return HFG_OK;
    }
    return hfg_squarem_prime(cfg, p0, p1, p2, *alpha_rate, prime);
}
