/*
 * hfg_summary.c -- the prediction summary tables of `hmm_flagger` (prediction_summary_<suffix>.tsv) on FLAT label arrays.
 *
 * Replaces:
 *   writeBenchmarkingStats                         programs/src/hmm_flagger.c:134-161
 *   SummaryTableList_createAndWriteAllTables       programs/submodules/summary_table/summary_table.c:1663-1747
 *   SummaryTableList_updateByUpdaterArgs           summary_table.c:930-1224   (the block scan)
 *   convertBaseLevelToOverlapBased                 summary_table.c:825-841
 *   SummaryTableListFullCatalog_write              summary_table.c:1385-1588  (row order and formats)
 *   SummaryTableList_writeFinalStatisticsIntoFile  summary_table.c:461-742    (<prefix>.benchmarking.tsv)
 *   SummaryTableList_writeFinalAunStatisticsIntoFile summary_table.c:744-813  (<prefix>.benchmarking.auN_ratio.tsv)
 * The reference walks 750k heap-allocated CoverageInfo/Inference objects once per (category, metric, comparison) on a
 * thread pool; here the labels are the flat int8 array the E-step returns and the window coordinates come from the chunk
 * descriptors.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/hfg_io.h"

enum { CMP_TRUTH_VS_PREDICTION = 0, CMP_PREDICTION_VS_TRUTH = 1, CMP_TRUTH = 2, CMP_PREDICTION = 3 }; /* summary_table.h:33-38 */
enum { METRIC_OVERLAP = 0, METRIC_BASE = 1, METRIC_AUN = 2 }; /* summary_table.h:22-26 */
enum { CAT_REGION = 0, CAT_ANNOTATION = 1 };
static const char *METRIC_NAME[3] = {"overlap_based", "base_level", "truth_based_auN"};
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

/* One table = n x n counts, then the n row totals and the grand total, accumulated increment by increment in the
 * reference's order (SummaryTable_increment, summary_table.c:69-89): the auN entries are fractions, so the percentages
 * printed from the totals depend on the order of the additions. */
#define TBL_STRIDE(n) ((size_t) (n) * (n) + (n) + 1)
#define TBL_ROWTOT(t, n, r) ((t)[(size_t) (n) * (n) + (r)])
#define TBL_TOTAL(t, n) ((t)[(size_t) (n) * (n) + (n)])

/* size bins of the reference-label blocks (IntBinArray, submodules/common/common.c:670-750): a block of length len is
 * counted in every bin with start <= len < end */
typedef struct SizeBins {
    int n;
    int *start, *end;
    char **name;
} SizeBins;

static void bins_free(SizeBins *b) {
    for (int i = 0; i < b->n; i++) free(b->name[i]);
    free(b->start);
    free(b->end);
    free(b->name);
}

/* NULL path: the single bin [0, 1e9) "ALL_SIZES" (summary_table.c:1672-1678); else a tab-delimited file
 * "start<TAB>end<TAB>name", '#' lines skipped, numbers read with atof (IntBinArray_constructFromFile) */
static int bins_load(const char *path, SizeBins *b) {
    memset(b, 0, sizeof(*b));
    if (!path) {
        b->n = 1;
        b->start = malloc(sizeof(int));
        b->end = malloc(sizeof(int));
        b->name = malloc(sizeof(char *));
        b->start[0] = 0;
        b->end[0] = (int) 1e9;
        b->name[0] = strdup("ALL_SIZES");
        return 1;
    }
    FILE *f = fopen(path, "r");
    if (!f) return 0;
    char line[4096];
    int cap = 0;
    while (fgets(line, sizeof(line), f)) {
        size_t len = strlen(line);
        if (len && line[len - 1] == '\n') line[--len] = '\0';
        if (line[0] == '#' || len == 0) continue;
        char *t1 = strchr(line, '\t');
        if (!t1) continue;
        char *t2 = strchr(t1 + 1, '\t');
        if (!t2) continue;
        if (b->n == cap) {
            cap = cap ? 2 * cap : 8;
            b->start = realloc(b->start, sizeof(int) * (size_t) cap);
            b->end = realloc(b->end, sizeof(int) * (size_t) cap);
            b->name = realloc(b->name, sizeof(char *) * (size_t) cap);
        }
        *t1 = *t2 = '\0';
        char *t3 = strchr(t2 + 1, '\t');
        if (t3) *t3 = '\0';
        b->start[b->n] = (int) atof(line);
        b->end[b->n] = (int) atof(t1 + 1);
        b->name[b->n] = strdup(t2 + 1);
        b->n++;
    }
    fclose(f);
    return b->n > 0;
}

/* growable list of block lengths (the auN metric keeps the query-label blocks of the current reference block) */
typedef struct IntList {
    int *v;
    int n, cap;
} IntList;

static void intlist_push(IntList *l, int x) {
    if (l->n == l->cap) {
        l->cap = l->cap ? 2 * l->cap : 16;
        l->v = realloc(l->v, sizeof(int) * (size_t) l->cap);
    }
    l->v[l->n++] = x;
}

/* adds one finished reference-label block to the tables of the size bins it falls in (summary_table.c:1040-1102 and
 * :1150-1214).  tables / aux: [n_bins] consecutive tables of this category index. */
static void flush_block(double *tables, double *row, int n, int metric, double overlap_threshold, int pre_ref, int len,
                        IntList *qlen, int pre_query, int pre_end, int query_start, const double *aux, const SizeBins *bins) {
    if (metric == METRIC_OVERLAP) {
        int hit = 0;
        for (int k = 0; k < n; k++) {
            const double ratio = row[k] / len;
            if (overlap_threshold < ratio) hit = 1;
            row[k] = overlap_threshold < ratio ? 1 : 0;
        }
        if (!hit) row[n - 1] = 1;
    }
    if (metric == METRIC_AUN) {
        /* the last query block of the reference block, then the sum of squared query-block lengths per query label */
        if (pre_query != -1) intlist_push(&qlen[pre_query], pre_end - query_start + 1);
        for (int k = 0; k < n; k++)
            for (int b = 0; b < qlen[k].n; b++) row[k] += (double) qlen[k].v[b] * qlen[k].v[b];
    }
    for (int bi = 0; bi < bins->n; bi++) {
        if (!(bins->start[bi] <= len && len < bins->end[bi])) continue;
        double *table = tables + (size_t) bi * TBL_STRIDE(n);
        /* auN: over the total length of this reference label in the category and bin (base_level truth-vs-truth table) */
        const double denom = metric == METRIC_AUN ? (aux + (size_t) bi * TBL_STRIDE(n))[(size_t) pre_ref * n + pre_ref] : 1.0;
        for (int k = 0; k < n; k++) {
            const double v = row[k] / denom;
            table[(size_t) pre_ref * n + k] += v;
            TBL_ROWTOT(table, n, pre_ref) += v;
            TBL_TOTAL(table, n) += v;
        }
    }
}

/* the confusion tables [n_bins][n][n] of one category index, filled by the block scan (SummaryTableList_updateByUpdaterArgs,
 * summary_table.c:930-1224).  aux: the base_level truth-vs-truth table of the same category index (auN only). */
static void scan_category(const hfg_cov_data *d, const int8_t *ref, const int8_t *query, int n, int cat_type, int index,
                          int metric, double overlap_threshold, const double *aux, const SizeBins *bins, double *table) {
    double *row = calloc((size_t) n, sizeof(double));
    IntList *qlen = calloc((size_t) n, sizeof(IntList));
    int pre_ref = -1, pre_query = -1, ref_start = -1, query_start = -1, pre_end = -1;
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
            const int ref_changed = r != pre_ref, query_changed = q != pre_query;
            const int cur_in = in_category(d, g, cat_type, index);
            const int continued = cur_in && prev_in, started = cur_in && !prev_in, ended = !cur_in && prev_in;
            /* a block of one reference label inside the category has ended: add it to the table */
            if (pre_ref != -1 && ((continued && ref_changed) || (prev_in && ctg_changed) || ended))
                flush_block(table, row, n, metric, overlap_threshold, pre_ref, pre_end - ref_start + 1, qlen, pre_query, pre_end,
                            query_start, aux, bins);
            /* the query label changed inside a reference block */
            if (cur_in && metric == METRIC_AUN && pre_query != -1 && query_changed && (continued && !ref_changed) && !ctg_changed)
                intlist_push(&qlen[pre_query], pre_end - query_start + 1);
            if ((!cur_in && ctg_changed) || ended) {
                ref_start = -1;
                query_start = -1;
                memset(row, 0, sizeof(double) * (size_t) n);
            }
            if ((continued && ref_changed) || (cur_in && ctg_changed) || started) {
                ref_start = start;
                memset(row, 0, sizeof(double) * (size_t) n);
                for (int k = 0; k < n; k++) qlen[k].n = 0;
            }
            if ((continued && ref_changed) || (continued && query_changed) || (cur_in && ctg_changed) || started) query_start = start;
            if (cur_in && metric != METRIC_AUN) row[q] += end - start + 1;
            have_prev = 1;
            prev_in = cur_in;
            pre_ref = r;
            pre_query = q;
            pre_ctg = ctg;
            pre_end = end;
        }
    }
    if (have_prev && prev_in && pre_ref != -1)
        flush_block(table, row, n, metric, overlap_threshold, pre_ref, pre_end - ref_start + 1, qlen, pre_query, pre_end, query_start,
                    aux, bins);
    for (int k = 0; k < n; k++) free(qlen[k].v);
    free(qlen);
    free(row);
}

static void write_values(FILE *f, const double *v, int n) {
    for (int k = 0; k < n; k++) fprintf(f, "%s%.2f", k ? "\t" : "", v[k]);
}

static const char *row_name(const char *const *label_names, int r, char *buf, size_t buflen) {
    if (label_names) return label_names[r];
    snprintf(buf, buflen, "%d", r);
    return buf;
}

static void pct_or_na(char *out, size_t len, int defined, double value) {
    if (defined) snprintf(out, len, "%.2f", value);
    else snprintf(out, len, "NA");
}

/* <prefix>.benchmarking.tsv: precision / recall / F1 per label and their averages
 * (SummaryTableList_writeFinalStatisticsIntoFile, summary_table.c:461-742).  recall: tables with the truth as reference,
 * precision: tables with the prediction as reference; [n_cat][n][n] each. */
static void write_final_statistics(FILE *f, const double *recall, const double *precision, int n_cat, int n, const char *metric,
                                   const char *category, const char *const *cat_names, const char *const *label_names,
                                   const SizeBins *bins) {
    const int n_labels = n - 1, HAP = 2;
    for (int cb = 0; cb < n_cat * bins->n; cb++) {
        const int ci = cb / bins->n;
        const char *bname = bins->name[cb % bins->n];
        const double *rt = recall + (size_t) cb * TBL_STRIDE(n), *pt = precision + (size_t) cb * TBL_STRIDE(n);
        double tot_tp_r = 0, tot_tp_p = 0, tot_r = 0, tot_p = 0, sum_r = 0, sum_p = 0, sum_r_nh = 0, sum_p_nh = 0;
        double rec_r = 0, rec_p = 0, rec_r_nh = 0, rec_p_nh = 0;
        int nz_r = 0, nz_p = 0, nz_r_nh = 0, nz_p_nh = 0;
        char cbuf[64], nbuf[32];
        const char *cname = cat_names ? cat_names[ci] : (snprintf(cbuf, sizeof(cbuf), "region_%d", ci), cbuf);
        for (int r = 0; r < n_labels; r++) {
            const double all_r = TBL_ROWTOT(rt, n, r), all_p = TBL_ROWTOT(pt, n, r);
            const double tp_r = rt[(size_t) r * n + r], tp_p = pt[(size_t) r * n + r];
            tot_tp_r += tp_r;
            tot_tp_p += tp_p;
            const double fn = all_r - tp_r, fp = all_p - tp_p;
            tot_r += tp_r + fn;
            tot_p += tp_p + fp;
            const double rp = tp_r / (tp_r + fn + 1.0e-9) * 100.0, pp = tp_p / (tp_p + fp + 1.0e-9) * 100.0;
            const int def_r = 1e-9 < (tp_r + fn), def_p = 1e-9 < (tp_p + fp);
            nz_r += def_r;
            nz_p += def_p;
            if (r != HAP) {
                nz_r_nh += def_r;
                nz_p_nh += def_p;
            }
            sum_r += rp;
            sum_p += pp;
            if (r != HAP) {
                sum_r_nh += rp;
                sum_p_nh += pp;
            }
            if (def_r) {
                rec_r += 0.0 < rp ? 1.0 / rp : 1.0e9;
                if (r != HAP) rec_r_nh += 0.0 < rp ? 1.0 / rp : 1.0e9;
            }
            if (def_p) {
                rec_p += 0.0 < pp ? 1.0 / pp : 1.0e9;
                if (r != HAP) rec_p_nh += 0.0 < pp ? 1.0 / pp : 1.0e9;
            }
            const double f1 = 2 * pp * rp / (pp + rp + 1.0e-9);
            char rs[20], ps[20], fs[20];
            pct_or_na(rs, sizeof(rs), def_r, rp);
            pct_or_na(ps, sizeof(ps), def_p, pp);
            pct_or_na(fs, sizeof(fs), def_r && def_p, f1);
            fprintf(f, "%s\t%s\t%s\t%s\t%s\t%.2f\t%.2f\t%.2f\t%.2f\t%.2f\t%.2f\t%s\t%s\t%s\tNA\tNA\n", metric, category, cname, bname,
                    row_name(label_names, r, nbuf, sizeof(nbuf)), tp_p, tp_r, fp, fn, tp_p + fp, tp_r + fn, ps, rs, fs);
        }
        const double mac_r = 0 < nz_r ? sum_r / nz_r : 0.0, mac_p = 0 < nz_p ? sum_p / nz_p : 0.0;
        const double mac_r_nh = 0 < nz_r_nh ? sum_r_nh / nz_r_nh : 0.0, mac_p_nh = 0 < nz_p_nh ? sum_p_nh / nz_p_nh : 0.0;
        const double har_r = 0 < nz_r ? (double) nz_r / rec_r : 0.0, har_p = 0 < nz_p ? (double) nz_p / rec_p : 0.0;
        const double har_r_nh = 0 < nz_r_nh ? (double) nz_r_nh / rec_r_nh : 0.0, har_p_nh = 0 < nz_p_nh ? (double) nz_p_nh / rec_p_nh : 0.0;
        struct { const char *name; double p, r; int dp, dr; } avg[4] = {
            {"MACRO_AVERAGE", mac_p, mac_r, 0 < nz_p, 0 < nz_r},
            {"MACRO_AVERAGE_NO_HAP", mac_p_nh, mac_r_nh, 0 < nz_p_nh, 0 < nz_r_nh},
            {"HARMONIC_MEAN", har_p, har_r, 0 < nz_p, 0 < nz_r},
            {"HARMONIC_MEAN_NO_HAP", har_p_nh, har_r_nh, 0 < nz_p_nh, 0 < nz_r_nh}};
        for (int k = 0; k < 4; k++) {
            char ps[20], rs[20], fs[20];
            pct_or_na(ps, sizeof(ps), avg[k].dp, avg[k].p);
            pct_or_na(rs, sizeof(rs), avg[k].dr, avg[k].r);
            pct_or_na(fs, sizeof(fs), avg[k].dp && avg[k].dr, 2 * avg[k].r * avg[k].p / (avg[k].r + avg[k].p + 1.0e-9));
            fprintf(f, "%s\t%s\t%s\t%s\t%s\tNA\tNA\tNA\tNA\tNA\tNA\t%s\t%s\t%s\tNA\tNA\n", metric, category, cname, bname, avg[k].name,
                    ps, rs, fs);
        }
        fprintf(f, "%s\t%s\t%s\t%s\tACCURACY\t%.2f\t%.2f\tNA\tNA\t%.2f\t%.2f\tNA\tNA\tNA\t%.2f\t%.2f\n", metric, category, cname, bname,
                tot_tp_p, tot_tp_r, tot_p, tot_r, tot_tp_p / (tot_p + 1.0e-9) * 100.0, tot_tp_r / (tot_r + 1e-9) * 100.0);
    }
}

/* <prefix>.benchmarking.auN_ratio.tsv (SummaryTableList_writeFinalAunStatisticsIntoFile, summary_table.c:744-813) */
static void write_aun_statistics(FILE *f, const double *num, const double *den, int n_cat, int n, const char *category,
                                 const char *const *cat_names, const char *const *label_names, const SizeBins *bins) {
    for (int cb = 0; cb < n_cat * bins->n; cb++) {
        const int ci = cb / bins->n;
        const char *bname = bins->name[cb % bins->n];
        const double *nt = num + (size_t) cb * TBL_STRIDE(n), *dt = den + (size_t) cb * TBL_STRIDE(n);
        char cbuf[64], nbuf[32], s1[20], s2[20];
        const char *cname = cat_names ? cat_names[ci] : (snprintf(cbuf, sizeof(cbuf), "region_%d", ci), cbuf);
        double sum = 0.0, rec = 0.0;
        int nz = 0;
        for (int r = 0; r < n - 1; r++) {
            const double de = dt[(size_t) r * n + r], nu = nt[(size_t) r * n + r], aun = nu / (de + 1e-9);
            nz += 0 < de ? 1 : 0;
            sum += aun;
            if (0 < de) rec += 0.0 < aun ? 1.0 / aun : 1.0e9;
            fprintf(f, "%s\t%s\t%s\t%s\t%.2f\n", category, cname, bname, row_name(label_names, r, nbuf, sizeof(nbuf)), aun);
        }
        pct_or_na(s1, sizeof(s1), 0 < nz, 0 < nz ? sum / nz : 0.0);
        pct_or_na(s2, sizeof(s2), 0 < nz, 0 < nz ? (double) nz / rec : 0.0);
        fprintf(f, "%s\t%s\t%s\tAVERAGE\t%s\n", category, cname, bname, s1);
        fprintf(f, "%s\t%s\t%s\tHARMONIC_MEAN\t%s\n", category, cname, bname, s2);
    }
}

int hfg_write_summary_tsv(const char *path, const hfg_cov_data *d, const int8_t *prediction, const int8_t *truth,
                          const char *const *label_names, int n_labels, double overlap_ratio_threshold,
                          const char *bin_array_file, char *err, size_t errlen) {
    if (!path || !d || n_labels < 1 || (!prediction && !truth)) {
        snprintf(err, errlen, "hfg_write_summary_tsv: bad argument");
        return HFG_ERR_INVALID;
    }
    SizeBins bins;
    if (!bins_load(bin_array_file, &bins)) {
        snprintf(err, errlen, "Error: Unable to read size bins from %s", bin_array_file);
        bins_free(&bins);
        return HFG_ERR_INVALID;
    }
    FILE *f = fopen(path, "w");
    if (!f) {
        snprintf(err, errlen, "Error: %s cannot be opened.", path);
        bins_free(&bins);
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
    /* precision / recall files exist only when both kinds of labels do (summary_table.c:1424-1450) */
    FILE *f_stats = NULL, *f_aun = NULL;
    if (truth && prediction) {
        char p2[4200];
        const size_t plen = strlen(path) >= 4 ? strlen(path) - 4 : strlen(path); /* ".tsv" */
        snprintf(p2, sizeof(p2), "%.*s.benchmarking.tsv", (int) plen, path);
        f_stats = fopen(p2, "w");
        snprintf(p2, sizeof(p2), "%.*s.benchmarking.auN_ratio.tsv", (int) plen, path);
        f_aun = fopen(p2, "w");
        if (!f_stats || !f_aun) {
            snprintf(err, errlen, "Error: %s cannot be opened.", p2);
            fclose(f);
            if (f_stats) fclose(f_stats);
            if (f_aun) fclose(f_aun);
            bins_free(&bins);
            return HFG_ERR_INVALID;
        }
        fprintf(f_stats, "#Metric_Type\tCategory_Type\tCategory_Name\tSize_Bin_Name\tLabel\tTP_Prediction_Ref\tTP_Truth_Ref\tFP\tFN\t"
                         "Total_Prediction_Ref\tTotal_Truth_Ref\tPrecision\tRecall\tF1-Score\tAccuracy_Prediction_Ref\tAccuracy_Truth_Ref\n");
        fprintf(f_aun, "#Category_Type\tCategory_Name\tSize_Bin_Name\tLabel\tauN_Ratio\n");
    }
    double *vals = malloc(sizeof(double) * (size_t) n);
    for (int cat_type = 0; cat_type < 2; cat_type++) {
        const int n_cat = cat_type == CAT_REGION ? d->n_regions : d->n_annotations;
        const char *const *cat_names = cat_type == CAT_ANNOTATION ? (const char *const *) d->annotation_names : NULL;
        double *tab[3][4];
        memset(tab, 0, sizeof(tab));
        /* all tables of this category type first (the auN tables need the base_level truth table) */
        for (int metric = 0; metric < 3; metric++) {
            for (int cmp = 0; cmp < 4; cmp++) {
                const int need_truth = cmp != CMP_PREDICTION, need_pred = cmp != CMP_TRUTH;
                if ((need_truth && !truth) || (need_pred && !prediction)) continue;
                if (metric == METRIC_AUN && (cmp == CMP_PREDICTION || cmp == CMP_PREDICTION_VS_TRUTH)) continue;
                const int8_t *ref = (cmp == CMP_TRUTH_VS_PREDICTION || cmp == CMP_TRUTH) ? truth : prediction;
                const int8_t *query = (cmp == CMP_TRUTH_VS_PREDICTION || cmp == CMP_PREDICTION) ? prediction : truth;
                tab[metric][cmp] = calloc((size_t) n_cat * bins.n * TBL_STRIDE(n), sizeof(double));
                for (int ci = 0; ci < n_cat; ci++)
                    scan_category(d, ref, query, n, cat_type, ci, metric, overlap_ratio_threshold,
                                  metric == METRIC_AUN ? tab[METRIC_BASE][CMP_TRUTH] + (size_t) ci * bins.n * TBL_STRIDE(n) : NULL, &bins,
                                  tab[metric][cmp] + (size_t) ci * bins.n * TBL_STRIDE(n));
            }
        }
        for (int metric = 0; metric < 3; metric++) {
            for (int cmp = 0; cmp < 4; cmp++) {
                const double *all = tab[metric][cmp];
                if (!all) continue;
                const int single_row = cmp == CMP_TRUTH || cmp == CMP_PREDICTION;
                /* the reference writes all counts of a (category type, metric, comparison) first, then all percentages */
                for (int pct = 0; pct < 2; pct++) {
                    for (int cb = 0; cb < n_cat * bins.n; cb++) {
                        const int ci = cb / bins.n;
                        const char *bname = bins.name[cb % bins.n];
                        const double *t = all + (size_t) cb * TBL_STRIDE(n);
                        char cname[64], nbuf[32];
                        const char *cat_name = cat_names ? cat_names[ci] : (snprintf(cname, sizeof(cname), "region_%d", ci), cname);
                        const double total = TBL_TOTAL(t, n);
                        if (single_row) {
                            /* total per reference label (SummaryTableList_writeTotalPerRow[Percentage]IntoFile) */
                            for (int r = 0; r < n; r++) {
                                const double s = TBL_ROWTOT(t, n, r);
                                vals[r] = pct ? (0 < total ? s / total * 100.0 : 0.0) : s;
                            }
                            fprintf(f, "%s\t%s\t%s\t%s\t%s\t%s\tALL_LABELS\t", COMPARISON_NAME[cmp], METRIC_NAME[metric],
                                    pct ? "percentage" : "count", CATEGORY_NAME[cat_type], cat_name, bname);
                            write_values(f, vals, n);
                            fprintf(f, "\n");
                        } else {
                            for (int r = 0; r < n; r++) {
                                const double s = TBL_ROWTOT(t, n, r);
                                for (int k = 0; k < n; k++)
                                    vals[k] = pct ? (0 < s ? t[(size_t) r * n + k] / s * 100.0 : 0.0) : t[(size_t) r * n + k];
                                fprintf(f, "%s\t%s\t%s\t%s\t%s\t%s\t%s\t", COMPARISON_NAME[cmp], METRIC_NAME[metric],
                                        pct ? "percentage" : "count", CATEGORY_NAME[cat_type], cat_name, bname,
                                        row_name(label_names, r, nbuf, sizeof(nbuf)));
                                write_values(f, vals, n);
                                fprintf(f, "\n");
                            }
                        }
                    }
                }
            }
            if (f_stats && metric != METRIC_AUN)
                write_final_statistics(f_stats, tab[metric][CMP_TRUTH_VS_PREDICTION], tab[metric][CMP_PREDICTION_VS_TRUTH], n_cat, n,
                                       METRIC_NAME[metric], CATEGORY_NAME[cat_type], cat_names, label_names, &bins);
        }
        if (f_aun)
            write_aun_statistics(f_aun, tab[METRIC_AUN][CMP_TRUTH_VS_PREDICTION], tab[METRIC_AUN][CMP_TRUTH], n_cat, n,
                                 CATEGORY_NAME[cat_type], cat_names, label_names, &bins);
        for (int metric = 0; metric < 3; metric++)
            for (int cmp = 0; cmp < 4; cmp++) free(tab[metric][cmp]);
    }
    free(vals);
    bins_free(&bins);
    fclose(f);
    if (f_stats) fclose(f_stats);
    if (f_aun) fclose(f_aun);
    return HFG_OK;
}

/* ---- the scores of the alpha-tuning driver ---------------------------------------------------------------------------- */

/* F1-Score of the HARMONIC_MEAN_NO_HAP row (write_final_statistics above), rounded as printed; NaN for "NA" */
static double harmonic_f1_no_hap(const double *rt, const double *pt, int n) {
    const int HAP = 2;
    double rec_r = 0, rec_p = 0;
    int nz_r = 0, nz_p = 0;
    for (int r = 0; r < n - 1; r++) {
        if (r == HAP) continue;
        const double all_r = TBL_ROWTOT(rt, n, r), all_p = TBL_ROWTOT(pt, n, r);
        const double tp_r = rt[(size_t) r * n + r], tp_p = pt[(size_t) r * n + r];
        const double fn = all_r - tp_r, fp = all_p - tp_p;
        const double rp = tp_r / (tp_r + fn + 1.0e-9) * 100.0, pp = tp_p / (tp_p + fp + 1.0e-9) * 100.0;
        if (1e-9 < (tp_r + fn)) {
            nz_r++;
            rec_r += 0.0 < rp ? 1.0 / rp : 1.0e9;
        }
        if (1e-9 < (tp_p + fp)) {
            nz_p++;
            rec_p += 0.0 < pp ? 1.0 / pp : 1.0e9;
        }
    }
    if (!(0 < nz_r && 0 < nz_p)) return 0.0 / 0.0;
    const double har_r = (double) nz_r / rec_r, har_p = (double) nz_p / rec_p;
    char buf[32];
    snprintf(buf, sizeof(buf), "%.2f", 2 * har_r * har_p / (har_r + har_p + 1.0e-9));
    return atof(buf);
}

int hfg_benchmark_scores(const hfg_cov_data *d, const int8_t *prediction, const int8_t *truth, int n_labels,
                         double overlap_ratio_threshold, const char *bin_array_file, const char *annotation_label,
                         const char *size_label, double scores[3], char *err, size_t errlen) {
    if (!d || !prediction || !truth || !annotation_label || !size_label || !scores || n_labels < 3) {
        snprintf(err, errlen, "hfg_benchmark_scores: bad argument");
        return HFG_ERR_INVALID;
    }
    SizeBins bins;
    if (!bins_load(bin_array_file, &bins)) {
        snprintf(err, errlen, "Error: Unable to read size bins from %s", bin_array_file);
        bins_free(&bins);
        return HFG_ERR_INVALID;
    }
    int ci = -1, bi = -1;
    for (int k = 0; k < d->n_annotations; k++)
        if (strcmp(d->annotation_names[k], annotation_label) == 0) ci = k;
    for (int k = 0; k < bins.n; k++)
        if (strcmp(bins.name[k], size_label) == 0) bi = k;
    if (ci < 0 || bi < 0) {
        snprintf(err, errlen, "hfg_benchmark_scores: annotation '%s' or size bin '%s' not found", annotation_label, size_label);
        bins_free(&bins);
        return HFG_ERR_INVALID;
    }
    const int n = n_labels + 1;
    const size_t stride = (size_t) bins.n * TBL_STRIDE(n);
    /* tables of this annotation only: [metric 0..1][T-vs-P, P-vs-T], base_level truth-vs-truth, auN T-vs-P and truth */
    double *tp[2], *pt_[2], *tt = calloc(stride, sizeof(double)), *aun_tp = calloc(stride, sizeof(double)),
                            *aun_tt = calloc(stride, sizeof(double));
    for (int metric = 0; metric < 2; metric++) {
        tp[metric] = calloc(stride, sizeof(double));
        pt_[metric] = calloc(stride, sizeof(double));
        scan_category(d, truth, prediction, n, CAT_ANNOTATION, ci, metric, overlap_ratio_threshold, NULL, &bins, tp[metric]);
        scan_category(d, prediction, truth, n, CAT_ANNOTATION, ci, metric, overlap_ratio_threshold, NULL, &bins, pt_[metric]);
    }
    scan_category(d, truth, truth, n, CAT_ANNOTATION, ci, METRIC_BASE, overlap_ratio_threshold, NULL, &bins, tt);
    scan_category(d, truth, prediction, n, CAT_ANNOTATION, ci, METRIC_AUN, overlap_ratio_threshold, tt, &bins, aun_tp);
    scan_category(d, truth, truth, n, CAT_ANNOTATION, ci, METRIC_AUN, overlap_ratio_threshold, tt, &bins, aun_tt);
    const size_t off = (size_t) bi * TBL_STRIDE(n);
    scores[0] = harmonic_f1_no_hap(tp[METRIC_OVERLAP] + off, pt_[METRIC_OVERLAP] + off, n);
    scores[1] = harmonic_f1_no_hap(tp[METRIC_BASE] + off, pt_[METRIC_BASE] + off, n);
    {
        /* 100 x the HARMONIC_MEAN row of the auN ratio file (write_aun_statistics), rounded as printed */
        const double *nt = aun_tp + off, *dt = aun_tt + off;
        double rec = 0.0;
        int nz = 0;
        for (int r = 0; r < n - 1; r++) {
            const double de = dt[(size_t) r * n + r], aun = nt[(size_t) r * n + r] / (de + 1e-9);
            if (0 < de) {
                nz++;
                rec += 0.0 < aun ? 1.0 / aun : 1.0e9;
            }
        }
        if (0 < nz) {
            char buf[32];
            snprintf(buf, sizeof(buf), "%.2f", (double) nz / rec);
            scores[2] = 100 * atof(buf);
        } else {
            scores[2] = 0.0 / 0.0;
        }
    }
    for (int metric = 0; metric < 2; metric++) {
        free(tp[metric]);
        free(pt_[metric]);
    }
    free(tt);
    free(aun_tp);
    free(aun_tt);
    bins_free(&bins);
    return HFG_OK;
}
