/*
 * hfg_nb.c -- host side of the negative-binomial model (`--modelType negative_binomial`, MODEL_NEGATIVE_BINOMIAL;
 * submodules/hmm_utils/hmm_utils.c:240-290,320-640,1662-1672).
 *
 * In this model the emission of a state depends on the window's coverage value alone (no previous-window dependency, no
 * contig-end factor), and the reference does not update its estimators pair by pair: every chunk accumulates, per (region,
 * state), a histogram of the pair mass over the coverage value and feeds each non-empty bin once through
 * NegativeBinomial_updateEstimator (hmm.c:615-617,643-649).  Both facts make the host the right place for everything that
 * is O(regions x states x 251 x components) and needs libm's lgamma / a long-double digamma:
 *   hfg_nb_emission_table        the pmf of every (region, state, x)            -> uploaded with the parameters
 *   hfg_nb_stats_from_histogram  the estimator sums from the (region, state, x) histogram the device reduces
 *   the M-step, model initialisation and feasibility test (reached through hfg_mstep / hfg_model_init /
 *   hfg_params_feasible, hfg_host_model.c)
 * The device side of a blocking call is a table look-up in the key-matrix phase, per-tile pair masses out of the statistics
 * phase and their fold into the histogram by the whole grid (hfg_estep_v3_kernel<THREADS, true>, hfg_nb_dev.cuh,
 * hfg_api.cu::run_blocking_nb; tests/test_gpu_nb.py); the device-resident loop keeps all of it on the device.  The functions here are checked
 * against the oracle on the CPU (tests/test_host_nb.py).
 *
 * Parameters: hfg_region_params.mean[s][c] = theta, .var[s][c] = lambda, .weight[s][c] = mixture weight; statistics:
 * hfg_region_stats.mean_* = theta estimator, .var_* = lambda estimator, .weight_* = weight estimator.
 */
#include <math.h>
#include <string.h>

#include "hfg_internal.h"
#include "hfg_digamma_coef.h" /* generated: tools/gen_digamma_coef.py */

#define HFG_HD static inline
#include "hfg_nb_mstep_inl.h" /* NB_MIN_COUNT, NB_PSEUDO, hfg_nb_mstep_region_inl: shared with the kernel tail */
#undef HFG_HD

/* digamma in long double: the routine the reference vendors (submodules/digamma/digamma.c:36-116, R. J. Mathar 2005, after
 * J. Wimp 1961) restated -- reflection below 0, psi(x) = psi(x + 1) - 1/x below 1, duplication above 3, closed forms at 1,
 * 2, 3, Chebyshev series in T_n(x - 2) in between. */
long double hfg_digammal(long double x) {
    const long double pi = 3.14159265358979323846264338327950288L;
    const long double euler = 0.577215664901532860606512090082402431L;
    const long double ln2 = 0.693147180559945309417232121458176568L;
    if (x < 0.0L) return hfg_digammal(1.0L - x) + pi / tanl(pi * (1.0L - x));
    if (x < 1.0L) return hfg_digammal(1.0L + x) - 1.0L / x;
    if (x == 1.0L) return -euler;
    if (x == 2.0L) return 1.0L - euler;
    if (x == 3.0L) return 1.5L - euler;
    if (x > 3.0L) return 0.5L * (hfg_digammal(x / 2.0L) + hfg_digammal((x + 1.0L) / 2.0L)) + ln2;
    const long double t = x - 2.0L;
    long double t_prev = 1.0L, t_cur = t;
    long double sum = hfg_digamma_coef[0] + hfg_digamma_coef[1] * t_cur;
    for (int n = 2; n < HFG_DIGAMMA_NCOEF; n++) {
        const long double t_next = 2.0L * t * t_cur - t_prev;
        sum += hfg_digamma_coef[n] * t_next;
        t_prev = t_cur;
        t_cur = t_next;
    }
    return sum;
}

/* r of the (theta, lambda) parametrisation, NegativeBinomial_getR (hmm_utils.c:455-458) */
static double nb_r(double theta, double lambda) { return -1 * lambda / log(theta); }

/* what the pmf of one component needs besides x, evaluated once per component instead of once per (component, x): the
 * same libm calls on the same arguments, so the values -- and the sum below, taken in the reference's order -- do not change */
typedef struct NbComp {
    double theta, r, w, lgamma_r, r_log_theta, log_1m_theta;
    double bt; /* -theta / (1 - theta) - 1 / log(theta), the factor of the theta estimator (hmm_utils.c:545) */
} NbComp;

static void nb_comp_setup(const hfg_region_params *p, int s, int n_comps, NbComp *c) {
    for (int k = 0; k < n_comps; k++) {
        c[k].theta = p->mean[s][k];
        c[k].r = nb_r(c[k].theta, p->var[s][k]);
        c[k].w = p->weight[s][k];
        c[k].lgamma_r = lgamma(c[k].r);
        c[k].r_log_theta = c[k].r * log(c[k].theta);
        c[k].log_1m_theta = log(1 - c[k].theta);
        c[k].bt = -1 * c[k].theta / (1 - c[k].theta) - 1 / log(c[k].theta);
    }
}

/* lgamma(x + 1) for the 251 coverage values */
static const double *nb_lgamma_x1(void) {
    static double table[HFG_NB_TABLE_X];
    static int ready = 0;
    if (!ready) { /* benign race: every thread writes the same values */
        for (int x = 0; x < HFG_NB_TABLE_X; x++) table[x] = lgamma(x + 1);
        __atomic_store_n(&ready, 1, __ATOMIC_RELEASE);
    }
    return table;
}

/* weighted pmf of every component at x, floored at 1e-40 (NegativeBinomial_getComponentProbs, hmm_utils.c:494-515).
 * Returns 1 if one of them is NaN (the reference exits there). */
static int nb_component_probs(const NbComp *c, int n_comps, int x, const double *lg_x1, double *probs) {
    int nan = 0;
    for (int k = 0; k < n_comps; k++) {
        double v = c[k].w * exp(lgamma(c[k].r + x) - c[k].lgamma_r - lg_x1[x] + c[k].r_log_theta + (double) x * c[k].log_1m_theta);
        if (v != v) nan = 1;
        if (v < 1e-40) v = 1e-40;
        probs[k] = v;
    }
    return nan;
}

int hfg_nb_emission_table(const hfg_config *cfg, const hfg_region_params *params, double *table) {
    if (!cfg || !params || !table || cfg->model_type != HFG_MODEL_NEGATIVE_BINOMIAL) return HFG_ERR_INVALID;
    int nan = 0;
    double probs[HFG_MAX_COMPS];
    NbComp comp[HFG_MAX_COMPS];
    const double *lg_x1 = nb_lgamma_x1();
    for (int r = 0; r < cfg->n_regions; r++)
        for (int s = 0; s < HFG_NS; s++) {
            nb_comp_setup(&params[r], s, cfg->n_comps[s], comp);
            for (int x = 0; x < HFG_NB_TABLE_X; x++) {
                nan |= nb_component_probs(comp, cfg->n_comps[s], x, lg_x1, probs);
                double total = 0.0;
                for (int c = 0; c < cfg->n_comps[s]; c++) total += probs[c]; /* NegativeBinomial_getProb, :479-484 */
                table[((size_t) r * HFG_NS + s) * HFG_NB_TABLE_X + x] = total;
            }
        }
    return nan ? HFG_ERR_NAN : HFG_OK;
}

int hfg_nb_stats_from_histogram(const hfg_config *cfg, const hfg_region_params *params, const double *histogram,
                                hfg_region_stats *stats) {
    if (!cfg || !params || !histogram || !stats || cfg->model_type != HFG_MODEL_NEGATIVE_BINOMIAL) return HFG_ERR_INVALID;
    int nan = 0;
    double probs[HFG_MAX_COMPS], psi[HFG_MAX_COMPS][HFG_NB_TABLE_X];
    NbComp comp[HFG_MAX_COMPS];
    const double *lg_x1 = nb_lgamma_x1();
    for (int reg = 0; reg < cfg->n_regions; reg++) {
        const hfg_region_params *p = &params[reg];
        hfg_region_stats *st = &stats[reg];
        for (int s = 0; s < HFG_NS; s++) {
            const int nc = cfg->n_comps[s];
            nb_comp_setup(p, s, nc, comp);
            for (int c = 0; c < nc; c++) {
                st->mean_num[s][c] = st->mean_den[s][c] = 0.0;
                st->var_num[s][c] = st->var_den[s][c] = 0.0;
                st->weight_num[s][c] = st->weight_den[s][c] = 0.0;
                /* psi(r + x) by the recurrence from one digamma(r) (NegativeBinomial_fillDigammaTable, :392-406) */
                const double r = comp[c].r;
                psi[c][0] = (double) hfg_digammal(r);
                for (int x = 1; x < HFG_NB_TABLE_X; x++) psi[c][x] = psi[c][x - 1] + 1.0 / (r + x - 1);
            }
            /* bins in ascending order, empty ones skipped (EmissionDistSeries_updateAllEstimatorsUsingCountData, :1662-1672);
             * one bin = NegativeBinomial_updateEstimator(x, mass) (:536-563) */
            for (int x = 0; x < HFG_NB_BINS; x++) {
                const double mass = histogram[((size_t) reg * HFG_NS + s) * HFG_NB_BINS + x];
                if (!(0 < mass)) continue;
                nan |= nb_component_probs(comp, nc, x, lg_x1, probs);
                double total = 0.0;
                for (int c = 0; c < nc; c++) total += probs[c];
                for (int c = 0; c < nc; c++) {
                    const double r = comp[c].r, bt = comp[c].bt;
                    const double w = mass * probs[c] / total;
                    const double delta = r * (psi[c][x] - psi[c][0]);
                    st->var_num[s][c] += w * delta;
                    st->var_den[s][c] += w;
                    st->mean_num[s][c] += w * delta * bt;
                    st->mean_den[s][c] += w * delta * bt + w * (x - delta);
                    st->weight_num[s][c] += w;
                    for (int k = 0; k < nc; k++) st->weight_den[s][k] += w; /* hmm_utils.c:66-74 */
                }
            }
        }
    }
    return nan ? HFG_ERR_NAN : HFG_OK;
}

/* theta / lambda of a component initialised from its mean with variance 1.5 x mean
 * (NegativeBinomial_constructByMean(.., 1.5, ..), hmm_utils.c:335-342,433-453,1635-1639) */
void hfg_nb_init_component(double mean, double *theta, double *lambda) {
    const double var = mean * 1.5;
    *theta = mean / var;
    const double r = pow(mean, 2) / (var - mean);
    *lambda = -1 * r * log(*theta);
}

/* EmissionDistSeries_estimateParameters for MODEL_NEGATIVE_BINOMIAL + Transition_estimateTransitionMatrix: hfg_nb_mstep_inl.h */
int hfg_nb_mstep_region(const int32_t *n_comps, hfg_region_params *p, const hfg_region_stats *st, double tol) {
    return hfg_nb_mstep_region_inl(n_comps, p, st, tol);
}
