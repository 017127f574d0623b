/*
 * hfg_summary.c -- the prediction summary tables of `hmm_flagger` (prediction_summary_<suffix>.tsv) on FLAT label arrays.
 *
 * Replaces, for the metric types overlap_based and base_level and the single default size bin (ALL_SIZES):
 *   writeBenchmarkingStats                         programs/src/hmm_flagger.c:134-161
 *   SummaryTableList_createAndWriteAllTables       programs/submodules/summary_table/summary_table.c:1663-1747
 *   SummaryTableList_updateByUpdaterArgs           summary_table.c:930-1224   (the block scan)
 *   convertBaseLevelToOverlapBased                 summary_table.c:825-841
 *   SummaryTableListFullCatalog_write              summary_table.c:1385-1588  (row order and formats)
 * The reference walks 750k heap-allocated CoverageInfo/Inference objects once per (category, metric, comparison) on a
 * thread pool; here the labels are the flat int8 array the E-step returns and the window coordinates come from the chunk
 * descriptors.  Not written: the truth_based_auN metric and the *.benchmarking*.tsv files (they exist only when the
 * input carries truth labels), --binArrayFile.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/hfg_io.h"

enum { CMP_TRUTH_VS_PREDICTION = 0, CMP_PREDICTION_VS_TRUTH = 1, CMP_TRUTH = 2, CMP_PREDICTION = 3 }; /* summary_table.h:33-38 */
enum { METRIC_OVERLAP = 0, METRIC_BASE = 1 };
enum { CAT_REGION = 0, CAT_ANNOTATION = 1 };
static const char *METRIC_NAME[2] = {"overlap_based", "base_level"};
static const char *CATEGORY_NAME[2] = {"region", "annotation"};
static const char *COMPARISON_NAME[4] = {"TRUTH_VS_PREDICTION", "PREDICTION_VS_TRUTH", "TRUTH", "PREDICTION"};

/* does window g belong to category `index`?  (CoverageInfo_overlapRegionIndex / _overlapAnnotationIndex,
 * submodules/ptBlock/ptBlock.c:245-255: annotation 0 is "no annotation bit set") */
static int in_category(const hfg_cov_data *d, int64_t g, int cat_type, int index) {
    const uint64_t flag = d->annotation_flag[g];
    if (cat_type == CAT_REGION) return index == (int) (flag >> 58);
    const uint64_t bits = flag & ~0xFC00000000000000ULL;
    if (bits == 0 && index == 0) return 1;
    return index > 0 && ((1ULL << (index - 1)) & flag) != 0;
}

/* one confusion table [n][n] (+ row totals) for one category index, filled by the block scan */
static void scan_category(const hfg_cov_data *d, const int8_t *ref, const int8_t *query, int n, int cat_type, int index,
                          int metric, double overlap_threshold, double *table) {
    double *row = calloc((size_t) n, sizeof(double));
    int pre_ref = -1, ref_start = -1, pre_end = -1;
    int have_prev = 0, prev_in = 0;
    const char *pre_ctg = NULL;
    for (int c = 0; c < d->n_chunks; c++) {
        const hfg_chunk_desc *ch = &d->chunks[c];
        const char *ctg = d->contig_names[c];
        for (int i = 0; i < ch->n_windows; i++) {
            const int64_t g = ch->offset + i;
            const int start = ch->s + i * ch->window_len;
            int end = ch->s + (i + 1) * ch->window_len - 1;
            if (end > ch->e) end = ch->e;
            int r = ref[g], q = query[g];
            if (r == -1) r = n - 1; /* the last row / column is "Unk" */
            if (q == -1) q = n - 1;
            const int ctg_changed = have_prev && strcmp(pre_ctg, ctg) != 0;
            const int ref_changed = r != pre_ref;
            const int cur_in = in_category(d, g, cat_type, index);
            const int continued = cur_in && prev_in, started = cur_in && !prev_in, ended = !cur_in && prev_in;
            /* a block of one reference label inside the category has ended: add it to the table */
            if (pre_ref != -1 && ((continued && ref_changed) || (prev_in && ctg_changed) || ended)) {
                const int len = pre_end - ref_start + 1;
                if (metric == METRIC_OVERLAP) {
                    int hit = 0;
                    for (int k = 0; k < n; k++) {
                        const double ratio = row[k] / len;
                        if (overlap_threshold < ratio) hit = 1;
                        row[k] = overlap_threshold < ratio ? 1 : 0;
                    }
                    if (!hit) row[n - 1] = 1;
                }
                for (int k = 0; k < n; k++) table[(size_t) pre_ref * n + k] += row[k];
            }
            if ((!cur_in && ctg_changed) || ended) {
                ref_start = -1;
                memset(row, 0, sizeof(double) * (size_t) n);
            }
            if ((continued && ref_changed) || (cur_in && ctg_changed) || started) {
                ref_start = start;
                memset(row, 0, sizeof(double) * (size_t) n);
            }
            if (cur_in) row[q] += end - start + 1;
            have_prev = 1;
            prev_in = cur_in;
            pre_ref = r;
            pre_ctg = ctg;
            pre_end = end;
        }
    }
    if (have_prev && prev_in && pre_ref != -1) {
        const int len = pre_end - ref_start + 1;
        if (metric == METRIC_OVERLAP) {
            int hit = 0;
            for (int k = 0; k < n; k++) {
                const double ratio = row[k] / len;
                if (overlap_threshold < ratio) hit = 1;
                row[k] = overlap_threshold < ratio ? 1 : 0;
            }
            if (!hit) row[n - 1] = 1;
        }
        for (int k = 0; k < n; k++) table[(size_t) pre_ref * n + k] += row[k];
    }
    free(row);
}

static void write_values(FILE *f, const double *v, int n) {
    for (int k = 0; k < n; k++) fprintf(f, "%s%.2f", k ? "\t" : "", v[k]);
}

int hfg_write_summary_tsv(const char *path, const hfg_cov_data *d, const int8_t *prediction, const int8_t *truth,
                          const char *const *label_names, int n_labels, double overlap_ratio_threshold, char *err,
                          size_t errlen) {
    if (!path || !d || n_labels < 1 || (!prediction && !truth)) {
        snprintf(err, errlen, "hfg_write_summary_tsv: bad argument");
        return HFG_ERR_INVALID;
    }
    FILE *f = fopen(path, "w");
    if (!f) {
        snprintf(err, errlen, "Error: %s cannot be opened.", path);
        return HFG_ERR_INVALID;
    }
    const int n = n_labels + 1; /* + "Unk" */
    fprintf(f, "#Statistic\tMetric_Type\tEntry_Type\tCategory_Type\tCategory_Name\tSize_Bin_Name\tRef_Label");
    for (int k = 0; k < n; k++) {
        if (label_names) fprintf(f, "\t%s", label_names[k]);
        else if (k == n - 1) fprintf(f, "\tlabel_unk");
        else fprintf(f, "\tlabel_%d", k);
    }
    fprintf(f, "\n");
    double *table = malloc(sizeof(double) * (size_t) n * n), *vals = malloc(sizeof(double) * (size_t) n);
    for (int cat_type = 0; cat_type < 2; cat_type++) {
        const int n_cat = cat_type == CAT_REGION ? d->n_regions : d->n_annotations;
        for (int metric = 0; metric < 2; metric++) {
            for (int cmp = 0; cmp < 4; cmp++) {
                const int need_truth = cmp != CMP_PREDICTION, need_pred = cmp != CMP_TRUTH;
                if ((need_truth && !truth) || (need_pred && !prediction)) continue;
                const int8_t *ref = (cmp == CMP_TRUTH_VS_PREDICTION || cmp == CMP_TRUTH) ? truth : prediction;
                const int8_t *query = (cmp == CMP_TRUTH_VS_PREDICTION || cmp == CMP_PREDICTION) ? prediction : truth;
                const int single_row = cmp == CMP_TRUTH || cmp == CMP_PREDICTION;
                /* the reference writes all counts of a (category type, metric, comparison) first, then all percentages:
                 * keep the tables of every category index */
                double *all = calloc((size_t) n_cat * n * n, sizeof(double));
                for (int ci = 0; ci < n_cat; ci++) {
                    memset(table, 0, sizeof(double) * (size_t) n * n);
                    scan_category(d, ref, query, n, cat_type, ci, metric, overlap_ratio_threshold, table);
                    memcpy(all + (size_t) ci * n * n, table, sizeof(double) * (size_t) n * n);
                }
                for (int pct = 0; pct < 2; pct++) {
                    for (int ci = 0; ci < n_cat; ci++) {
                        const double *t = all + (size_t) ci * n * n;
                        char cname[64];
                        const char *cat_name = d->annotation_names && cat_type == CAT_ANNOTATION ? d->annotation_names[ci] : NULL;
                        if (!cat_name) {
                            snprintf(cname, sizeof(cname), "region_%d", ci);
                            cat_name = cname;
                        }
                        double total = 0.0;
                        for (int k = 0; k < n * n; k++) total += t[k];
                        if (single_row) {
                            /* total per reference label (SummaryTableList_writeTotalPerRow[Percentage]IntoFile) */
                            for (int r = 0; r < n; r++) {
                                double s = 0.0;
                                for (int k = 0; k < n; k++) s += t[(size_t) r * n + k];
                                vals[r] = pct ? (0 < total ? s / total * 100.0 : 0.0) : s;
                            }
                            fprintf(f, "%s\t%s\t%s\t%s\t%s\tALL_SIZES\tALL_LABELS\t", COMPARISON_NAME[cmp], METRIC_NAME[metric],
                                    pct ? "percentage" : "count", CATEGORY_NAME[cat_type], cat_name);
                            write_values(f, vals, n);
                            fprintf(f, "\n");
                        } else {
                            for (int r = 0; r < n; r++) {
                                double s = 0.0;
                                for (int k = 0; k < n; k++) s += t[(size_t) r * n + k];
                                for (int k = 0; k < n; k++)
                                    vals[k] = pct ? (0 < s ? t[(size_t) r * n + k] / s * 100.0 : 0.0) : t[(size_t) r * n + k];
                                fprintf(f, "%s\t%s\t%s\t%s\t%s\tALL_SIZES\t", COMPARISON_NAME[cmp], METRIC_NAME[metric],
                                        pct ? "percentage" : "count", CATEGORY_NAME[cat_type], cat_name);
                                if (label_names) fprintf(f, "%s\t", label_names[r]);
                                else fprintf(f, "%d\t", r);
                                write_values(f, vals, n);
                                fprintf(f, "\n");
                            }
                        }
                    }
                }
                free(all);
            }
        }
    }
    free(table);
    free(vals);
    fclose(f);
    return HFG_OK;
}
