/*
 * hfg_nb_mstep_inl.h -- the M-step of the negative-binomial model, one copy for the host (hfg_nb.c) and for the kernel tail of
 * the device-resident loop (hfg_nb_dev.cuh).  The includer defines HFG_HD (function qualifiers).
 */
#define NB_MIN_COUNT 10.0   /* MIN_COUNT_FOR_PARAMETER_UPDATE, hmm_utils.h:11 */
#define NB_PSEUDO 0.001     /* TRANSITION_PSEUDO_COUNT_VALUE, hmm.c:16 */

/* binding coefficients of ParameterBinding_getDefault1DArrayForNegativeBinomial (hmm_utils.c:240-290): theta is tied with
 * coefficient 1 across all states and components, lambda with 0.1 / 0.5 / 1 / 2 + c, weights are free */
HFG_HD double nb_lambda_coef(int s, int c) {
    switch (s) {
        case HFG_STATE_ERR: return 0.1;
        case HFG_STATE_DUP: return 0.5;
        case HFG_STATE_HAP: return 1.0;
        default: return 2.0 + 1.0 * c;
    }
}

HFG_HD int nb_settled(double old_value, double new_value, double tol) { /* hmm_utils.c:565-582 */
    const double diff = 1.0e-4 < old_value ? fabs(new_value / old_value - 1.0) : 0.0;
    return diff < tol;
}

/* EmissionDistSeries_estimateParameters for MODEL_NEGATIVE_BINOMIAL (hmm_utils.c:1791-1858,1884-1900) followed by
 * Transition_estimateTransitionMatrix (:2185-2219).  Returns 1 when every updated value moved by less than tol. */
HFG_HD int hfg_nb_mstep_region_inl(const int32_t *n_comps, hfg_region_params *p, const hfg_region_stats *st, double tol) {
    int settled = 1;
    for (int type = 0; type < 2; type++) { /* theta, then lambda: one pooled ("bound") estimate each */
        double num = 0.0, den = 0.0;
        for (int s = 0; s < HFG_NS; s++)
            for (int c = 0; c < n_comps[s]; c++) {
                const double f = type == 0 ? 1.0 : nb_lambda_coef(s, c);
                num += (type == 0 ? st->mean_num[s][c] : st->var_num[s][c]) / f;
                den += type == 0 ? st->mean_den[s][c] : st->var_den[s][c];
            }
        const double pooled = den == 0 ? 0.0 : num / den; /* ParameterEstimator_getEstimation, :76-92 */
        if (!(NB_MIN_COUNT < den)) continue;
        for (int s = 0; s < HFG_NS; s++)
            for (int c = 0; c < n_comps[s]; c++) {
                double *dst = type == 0 ? &p->mean[s][c] : &p->var[s][c];
                const double v = pooled * (type == 0 ? 1.0 : nb_lambda_coef(s, c));
                settled &= nb_settled(*dst, v, tol);
                *dst = v;
            }
    }
    for (int s = 0; s < HFG_NS; s++)
        for (int c = 0; c < n_comps[s]; c++) {
            const double den = st->weight_den[s][c];
            if (!(NB_MIN_COUNT < den)) continue;
            const double v = st->weight_num[s][c] / den;
            settled &= nb_settled(p->weight[s][c], v, tol);
            p->weight[s][c] = v;
        }
    for (int a = 0; a < HFG_NS; a++) {
        double row = 0.0;
        for (int b = 0; b < HFG_NS; b++) row += st->trans_count[a][b] + NB_PSEUDO;
        for (int b = 0; b < HFG_NS; b++) {
            const double old_value = p->trans[a][b];
            const double v = (st->trans_count[a][b] + NB_PSEUDO) / row * (1.0 - HFG_TERM_PROB);
            p->trans[a][b] = v;
            const double diff = 1.0e-6 < old_value ? fabs(v / old_value - 1.0) : 0.0;
            settled &= diff < tol;
        }
    }
    for (int a = 0; a < HFG_NS; a++) p->trans[a][HFG_NS] = HFG_TERM_PROB;
    for (int b = 0; b < HFG_NS; b++) p->trans[HFG_NS][b] = 1.0 / HFG_NS;
    p->trans[HFG_NS][HFG_NS] = 0.0;
    return settled;
}
