/*
 * hfg_mstep_inl.h -- the M-step of one region, HMM_estimateParameters (hmm.c:120-127; hmm_utils.c:1791-1903,2185-2219),
 * written once and compiled twice: as plain C for the host mirror hfg_mstep (hfg_host_model.c) and as device code for the
 * tail of the E-step kernel (device-resident EM loop, hfg_api.cu).  Both are built without FMA contraction, so they
 * differ only where libm and the CUDA math library round log/exp differently (the golden-section comparisons).
 * The includer defines HFG_HD (function qualifiers).
 */
#ifndef HFG_MSTEP_INL_H
#define HFG_MSTEP_INL_H

#define MIN_COUNT_FOR_UPDATE 10.0  /* hmm_utils.h:11 */
#define TRUNC_POINT_FRACTION 0.25  /* hmm_utils.h:12 */
#define PSEUDO_COUNT 0.001         /* hmm.c:16 */
#define GOLDEN_TOL 1e-6            /* hmm_utils.c:86 */

HFG_HD int hfg_is_gaussian_state(int model_type, int s) {
    return !(model_type == HFG_MODEL_TRUNC_EXP_GAUSSIAN && s == HFG_STATE_ERR);
}

/* mean/var binding coefficient of component c of state s (hmm_utils.c:191-238): Err 0.1, Dup 0.5, Hap 1, Col 2,3,.. */
HFG_HD double hfg_binding(int s, int c) {
    return s == 0 ? 0.1 : (s == 1 ? 0.5 : (s == 2 ? 1.0 : 2.0 + 1.0 * c));
}

/* log-likelihood of the truncated exponential as a function of the rate (hmm_utils.c:949-956) */
HFG_HD double hfg_trunc_exp_objective(double rate, double trunc, double sum_x, double sum_w) {
    return sum_w * log(rate) - sum_w * log(1.0 - exp(-rate * trunc)) - sum_x * rate;
}

/* golden-section maximiser on (0, trunc] (hmm_utils.c:969-1011) */
HFG_HD double hfg_fit_rate(double trunc, double sum_x, double sum_w) {
    const double inv_phi = (sqrt(5.0) - 1.0) / 2.0, inv_phi2 = (3.0 - sqrt(5.0)) / 2.0;
    double lo = 0.0, hi = trunc, span = hi - lo;
    if (span <= GOLDEN_TOL) return (hi + lo) / 2.0;
    const int steps = (int) ceil(log(GOLDEN_TOL / span) / log(inv_phi));
    double x1 = lo + inv_phi2 * span, x2 = lo + inv_phi * span;
    double y1 = hfg_trunc_exp_objective(x1, trunc, sum_x, sum_w), y2 = hfg_trunc_exp_objective(x2, trunc, sum_x, sum_w);
    for (int k = 0; k < steps - 1; k++) {
        span = inv_phi * span;
        if (y1 > y2) {
            hi = x2; x2 = x1; y2 = y1;
            x1 = lo + inv_phi2 * span;
            y1 = hfg_trunc_exp_objective(x1, trunc, sum_x, sum_w);
        } else {
            lo = x1; x1 = x2; y1 = y2;
            x2 = lo + inv_phi * span;
            y2 = hfg_trunc_exp_objective(x2, trunc, sum_x, sum_w);
        }
    }
    return y1 > y2 ? (lo + x2) / 2.0 : (x1 + hi) / 2.0;
}

#ifdef __CUDACC__
/* The same maximiser for a whole warp: every lane calls it with identical arguments and gets the bits the serial routine
 * above would produce.  The serial search is a chain of ~30 dependent objective evaluations (two logs and an exp each:
 * ~50 us for one GPU thread, more than half an E-step).  The POINT evaluated at a step depends only on the outcomes of
 * the comparisons so far, not on the objective values, so the 2 + 4 + 8 + 16 candidate points of the next four steps are
 * evaluated speculatively, one per lane, and the true path is then walked with shuffles: 4 steps per round. */
static __device__ double hfg_fit_rate_warp(double trunc, double sum_x, double sum_w) {
    const unsigned full = 0xffffffffu;
    const int lane = (int) (threadIdx.x & 31);
    const double inv_phi = (sqrt(5.0) - 1.0) / 2.0, inv_phi2 = (3.0 - sqrt(5.0)) / 2.0;
    double lo = 0.0, hi = trunc, span = hi - lo;
    if (span <= GOLDEN_TOL) return (hi + lo) / 2.0;
    const int steps = (int) ceil(log(GOLDEN_TOL / span) / log(inv_phi));
    double x1 = lo + inv_phi2 * span, x2 = lo + inv_phi * span;
    double y1, y2;
    {
        const double y = hfg_trunc_exp_objective(lane == 0 ? x1 : x2, trunc, sum_x, sum_w);
        y1 = __shfl_sync(full, y, 0);
        y2 = __shfl_sync(full, y, 1);
    }
    /* lane -> (step t of the round, outcomes b_0..b_t of its comparisons, b_0 in the top bit): lanes 0-1 step 0,
     * 2-5 step 1, 6-13 step 2, 14-29 step 3 */
    int my_t = 0;
    while ((4 << my_t) - 2 <= lane) my_t++;
    const int my_bits = lane - ((2 << my_t) - 2);
    const int total = steps - 1;
    for (int k = 0; k < total; k += 4) {
        const int d = total - k < 4 ? total - k : 4;
        /* positions along this lane's assumed path (the arithmetic of the serial loop, same order) */
        double slo = lo, sx1 = x1, sx2 = x2, sspan = span, xnew = x1;
        for (int t = 0; t <= my_t && t < 4; t++) {
            const int b = (my_bits >> (my_t - t)) & 1;
            sspan = inv_phi * sspan;
            if (b) {
                sx2 = sx1;
                sx1 = slo + inv_phi2 * sspan;
                xnew = sx1;
            } else {
                slo = sx1;
                sx1 = sx2;
                sx2 = slo + inv_phi * sspan;
                xnew = sx2;
            }
        }
        const double ynew = hfg_trunc_exp_objective(xnew, trunc, sum_x, sum_w);
        /* the true path */
        int path = 0;
        for (int t = 0; t < d; t++) {
            const int b = y1 > y2;
            path = (path << 1) | b;
            const double y = __shfl_sync(full, ynew, ((2 << t) - 2) + path);
            span = inv_phi * span;
            if (b) {
                hi = x2; x2 = x1; y2 = y1;
                x1 = lo + inv_phi2 * span;
                y1 = y;
            } else {
                lo = x1; x1 = x2; y1 = y2;
                x2 = lo + inv_phi * span;
                y2 = y;
            }
        }
    }
    return y1 > y2 ? (lo + x2) / 2.0 : (x1 + hi) / 2.0;
}
#endif

/* relative-change test of Gaussian_updateParameter / TruncExponential_updateParameter (hmm_utils.c:855-858,1051-1053) */
HFG_HD int hfg_settled(double before, double after, double tol, double floor_) {
    const double change = floor_ < before ? fabs(after / before - 1.0) : 0.0;
    return change < tol;
}

#ifndef HFG_FIT_RATE
#define HFG_FIT_RATE hfg_fit_rate
#endif

/* The M-step of one region in three independent pieces (they touch disjoint parameters; the device build runs them
 * on three warps at once).  Each returns 1 when every parameter it updated moved by less than the tolerance. */

/* Gaussian means, variances and mixture weights */
HFG_HD int hfg_mstep_gauss(int model_type, const int32_t *n_comps, hfg_region_params *p, const hfg_region_stats *st,
                           double convergence_tol) {
    int all_settled = 1;
    /* mean, then variance: ONE pooled ("bound") estimate shared by every Gaussian component through its binding
     * coefficient (EmissionDistSeries_getBoundParameterEstimator / _estimateOneParameterType, hmm_utils.c:1791-1858) */
    for (int which = 0; which < 2; which++) {
        const double (*num)[HFG_MAX_COMPS] = which == 0 ? st->mean_num : st->var_num;
        const double (*den)[HFG_MAX_COMPS] = which == 0 ? st->mean_den : st->var_den;
        double (*dst)[HFG_MAX_COMPS] = which == 0 ? p->mean : p->var;
        double pooled_num = 0.0, pooled_den = 0.0;
        for (int s = 0; s < HFG_NS; s++) {
            if (!hfg_is_gaussian_state(model_type, s)) continue;
            for (int c = 0; c < n_comps[s]; c++) {
                pooled_num += num[s][c] / hfg_binding(s, c);
                pooled_den += den[s][c];
            }
        }
        if (!(MIN_COUNT_FOR_UPDATE < pooled_den)) continue; /* hmm_utils.c:1846 */
        const double unit = pooled_num / pooled_den;
        for (int s = 0; s < HFG_NS; s++) {
            if (!hfg_is_gaussian_state(model_type, s)) continue;
            for (int c = 0; c < n_comps[s]; c++) {
                const double v = unit * hfg_binding(s, c);
                all_settled &= hfg_settled(dst[s][c], v, convergence_tol, 1.0e-4);
                dst[s][c] = v;
            }
        }
    }
    /* mixture weights: unbound, each component from its own estimator (binding coefficient 0) */
    for (int s = 0; s < HFG_NS; s++) {
        if (!hfg_is_gaussian_state(model_type, s)) continue;
        for (int c = 0; c < n_comps[s]; c++) {
            const double d = st->weight_den[s][c];
            if (!(MIN_COUNT_FOR_UPDATE < d)) continue;
            const double v = st->weight_num[s][c] / d;
            all_settled &= hfg_settled(p->weight[s][c], v, convergence_tol, 1.0e-4);
            p->weight[s][c] = v;
        }
    }
    return all_settled;
}

/* rate of the truncated exponential: golden-section fit against the truncation point still in force
 * (hmm_utils.c:1872-1882); the truncation point itself follows the new Hap mean afterwards (hfg_mstep_region) */
HFG_HD int hfg_mstep_rate(int model_type, hfg_region_params *p, const hfg_region_stats *st, double convergence_tol) {
    int all_settled = 1;
    if (model_type == HFG_MODEL_TRUNC_EXP_GAUSSIAN && MIN_COUNT_FOR_UPDATE < st->lambda_den) {
        const double v = HFG_FIT_RATE(p->trunc_point, st->lambda_num, st->lambda_den);
        all_settled &= hfg_settled(p->lambda, v, convergence_tol, 1.0e-4);
        p->lambda = v;
    }
    return all_settled;
}

HFG_HD int hfg_mstep_trans(hfg_region_params *p, const hfg_region_stats *st, double convergence_tol) {
    int all_settled = 1;
    /* transition rows (Transition_estimateTransitionMatrix, hmm_utils.c:2185-2219) */
    for (int a = 0; a < HFG_NS; a++) {
        double row = 0.0;
        for (int b = 0; b < HFG_NS; b++) row += st->trans_count[a][b] + PSEUDO_COUNT;
        for (int b = 0; b < HFG_NS; b++) {
            const double v = (st->trans_count[a][b] + PSEUDO_COUNT) / row * (1.0 - HFG_TERM_PROB);
            all_settled &= hfg_settled(p->trans[a][b], v, convergence_tol, 1.0e-6);
            p->trans[a][b] = v;
        }
        p->trans[a][HFG_NS] = HFG_TERM_PROB;
    }
    for (int b = 0; b < HFG_NS; b++) p->trans[HFG_NS][b] = 1.0 / HFG_NS;
    p->trans[HFG_NS][HFG_NS] = 0.0;
    return all_settled;
}

/* one region, in the reference's order; returns 1 when every updated parameter moved by less than the tolerance */
HFG_HD int hfg_mstep_region(int model_type, const int32_t *n_comps, hfg_region_params *p, const hfg_region_stats *st,
                            double convergence_tol) {
    int all_settled = hfg_mstep_gauss(model_type, n_comps, p, st, convergence_tol);
    all_settled &= hfg_mstep_rate(model_type, p, st, convergence_tol);
    if (model_type == HFG_MODEL_TRUNC_EXP_GAUSSIAN) p->trunc_point = p->mean[HFG_STATE_HAP][0] * TRUNC_POINT_FRACTION;
    all_settled &= hfg_mstep_trans(p, st, convergence_tol);
    return all_settled;
}

#endif /* HFG_MSTEP_INL_H */
