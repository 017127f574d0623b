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
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

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

/* Window coordinates and contig runs, built once per call; the scans below then need no chunk look-up, no strcmp. */
typedef struct WinIndex {
    int64_t n;
    int32_t *start, *end, *ctg; /* ctg: run id of the contig name, +1 whenever a chunk's name differs from the previous one's */
    int64_t *cum;               /* cum[g] = bases in the windows before g (n + 1 entries) */
    int32_t *run_end;           /* last window of the stretch around g in which neither label array nor the contig changes */
} WinIndex;

static void winindex_free(WinIndex *w) {
    free(w->start);
    free(w->end);
    free(w->ctg);
    free(w->cum);
    free(w->run_end);
    memset(w, 0, sizeof(*w));
}

/* run_end[g] for the label arrays a, b (either may be NULL) */
static void winindex_set_runs(const WinIndex *w, const int8_t *a, const int8_t *b, int32_t *run_end) {
    for (int64_t g = w->n - 1; g >= 0; g--) {
        const int same = g + 1 < w->n && w->ctg[g] == w->ctg[g + 1] && (!a || a[g] == a[g + 1]) && (!b || b[g] == b[g + 1]);
        run_end[g] = same ? run_end[g + 1] : (int32_t) g;
    }
}

/* a, b: the label arrays the scans will compare (either may be NULL) */
static int winindex_build(const hfg_cov_data *d, const int8_t *a, const int8_t *b, WinIndex *w) {
    memset(w, 0, sizeof(*w));
    int64_t total = 0;
    for (int c = 0; c < d->n_chunks; c++) total += d->chunks[c].n_windows;
    if (total >= INT32_MAX) return 0;
    w->n = total;
    const size_t cap = (size_t) (total > 0 ? total : 1);
    w->start = malloc(sizeof(int32_t) * cap);
    w->end = malloc(sizeof(int32_t) * cap);
    w->ctg = malloc(sizeof(int32_t) * cap);
    w->cum = malloc(sizeof(int64_t) * (cap + 1));
    w->run_end = malloc(sizeof(int32_t) * cap);
    if (!w->start || !w->end || !w->ctg || !w->cum || !w->run_end) {
        winindex_free(w);
        return 0;
    }
    int64_t g = 0;
    int32_t run = 0;
    for (int c = 0; c < d->n_chunks; c++) {
        const hfg_chunk_desc *ch = &d->chunks[c];
        if (c > 0 && strcmp(d->contig_names[c - 1], d->contig_names[c]) != 0) run++;
        for (int i = 0; i < ch->n_windows; i++, g++) {
            const int end = ch->s + (i + 1) * ch->window_len - 1;
            w->start[g] = ch->s + i * ch->window_len;
            w->end[g] = end > ch->e ? ch->e : end;
            w->ctg[g] = run;
        }
    }
    w->cum[0] = 0;
    for (g = 0; g < total; g++) w->cum[g + 1] = w->cum[g] + (w->end[g] - w->start[g] + 1);
    winindex_set_runs(w, a, b, w->run_end);
    return 1;
}

/* The windows the scan of one category has to look at: those inside the category and, to close its last block, the window
 * right after each run of them.  Everything else only hands state from window to window that nothing reads (a block can
 * neither grow nor end outside the category), so a sparse category -- a bias annotation, a region -- costs its own size,
 * not the genome's.  g: window numbers in scan order; in: 1 inside the category. */
typedef struct Visit {
    int64_t n;
    int32_t *g;
    uint8_t *in;
    int32_t *stretch_last; /* for an entry inside the category: the last window of its run of consecutive such entries */
} Visit;

static void visit_free(Visit *v) {
    free(v->g);
    free(v->in);
    free(v->stretch_last);
    memset(v, 0, sizeof(*v));
}

static int visit_build(const hfg_cov_data *d, const WinIndex *w, int cat_type, int index, Visit *v) {
    memset(v, 0, sizeof(*v));
    int64_t count = 0;
    int prev_in = 0;
    for (int64_t g = 0; g < w->n; g++) {
        const int cur = in_category(d, g, cat_type, index);
        count += cur || prev_in;
        prev_in = cur;
    }
    v->g = malloc(sizeof(int32_t) * (size_t) (count > 0 ? count : 1));
    v->in = malloc((size_t) (count > 0 ? count : 1));
    v->stretch_last = malloc(sizeof(int32_t) * (size_t) (count > 0 ? count : 1));
    if (!v->g || !v->in || !v->stretch_last) {
        visit_free(v);
        return 0;
    }
    prev_in = 0;
    for (int64_t g = 0; g < w->n; g++) {
        const int cur = in_category(d, g, cat_type, index);
        if (cur || prev_in) {
            v->g[v->n] = (int32_t) g;
            v->in[v->n] = (uint8_t) cur;
            v->n++;
        }
        prev_in = cur;
    }
    for (int64_t k = v->n - 1; k >= 0; k--)
        v->stretch_last[k] = (v->in[k] && k + 1 < v->n && v->in[k + 1] && v->g[k + 1] == v->g[k] + 1) ? v->stretch_last[k + 1] : v->g[k];
    return 1;
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
 * summary_table.c:930-1224).  aux: the base_level truth-vs-truth table of the same category index (auN only).  The scan is
 * the reference's window-by-window state machine, run over the windows of `visit` only: a visited window whose predecessor
 * was skipped starts a block exactly as the first window of the genome does.  Returns 0 when out of memory. */
static int scan_category(const WinIndex *w, const Visit *visit, const int8_t *ref, const int8_t *query, int n, int metric,
                         double overlap_threshold, const double *aux, const SizeBins *bins, double *table) {
    double *row = calloc((size_t) n, sizeof(double));
    IntList *qlen = calloc((size_t) n, sizeof(IntList));
    if (!row || !qlen) {
        free(row);
        free(qlen);
        return 0;
    }
    int pre_ref = -1, pre_query = -1, ref_start = -1, query_start = -1, pre_end = -1;
    int have_prev = 0, prev_in = 0;
    int32_t pre_ctg = -1, pre_g = -2;
    for (int64_t k = 0; k < visit->n; k++) {
        const int32_t g = visit->g[k];
        const int cur_in = visit->in[k];
        if (g != pre_g + 1) { /* the windows in between were outside the category */
            have_prev = 0;
            prev_in = 0;
        }
        const int start = w->start[g], end = w->end[g];
        const int32_t ctg = w->ctg[g];
        int r = ref[g], q = query[g];
        if (r == -1) r = n - 1; /* the last row / column is "Unk" */
        if (q == -1) q = n - 1;
        const int ctg_changed = have_prev && pre_ctg != ctg;
        const int ref_changed = r != pre_ref, query_changed = q != pre_query;
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
            for (int j = 0; j < n; j++) qlen[j].n = 0;
        }
        if ((continued && ref_changed) || (continued && query_changed) || (cur_in && ctg_changed) || started) query_start = start;
        if (cur_in && metric != METRIC_AUN) row[q] += end - start + 1;
        have_prev = 1;
        prev_in = cur_in;
        pre_ref = r;
        pre_query = q;
        pre_ctg = ctg;
        pre_end = end;
        pre_g = g;
        /* the windows that follow with nothing changed -- next in scan order, inside the category, same contig, same pair of
         * labels -- take none of the branches above: they only lengthen the current block.  Their end is known from the
         * two precomputed run tables, their bases from the prefix sums (integers: one addition equals the many) */
        if (cur_in) {
            int32_t last = w->run_end[g];
            if (visit->stretch_last[k] < last) last = visit->stretch_last[k];
            if (last > g) {
                if (metric != METRIC_AUN) row[q] += (double) (w->cum[last + 1] - w->cum[g + 1]);
                k += last - g;
                pre_g = last;
                pre_end = w->end[last];
            }
        }
    }
    if (have_prev && prev_in && pre_ref != -1)
        flush_block(table, row, n, metric, overlap_threshold, pre_ref, pre_end - ref_start + 1, qlen, pre_query, pre_end, query_start,
                    aux, bins);
    for (int j = 0; j < n; j++) free(qlen[j].v);
    free(qlen);
    free(row);
    return 1;
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

/* the scans of one category type, one category per job */
typedef struct CatJobs {
    const hfg_cov_data *d;
    const WinIndex *wi;
    const int8_t *prediction, *truth;
    const SizeBins *bins;
    int n, n_cat[2]; /* categories per type: regions, annotations */
    double overlap_threshold;
    double *tab[2][3][4]; /* [category type][metric][comparison] -> [n_cat][n_bins] tables, NULL where not applicable */
    int next, oom;
    pthread_mutex_t mu;
} CatJobs;

static void *cat_worker(void *arg) {
    CatJobs *j = arg;
    for (;;) {
        const int n_jobs = j->n_cat[0] + j->n_cat[1];
        pthread_mutex_lock(&j->mu);
        const int job = j->oom ? n_jobs : j->next++;
        pthread_mutex_unlock(&j->mu);
        if (job >= n_jobs) return NULL;
        /* annotations first: "whole_genome" covers every window and is the longest job */
        const int cat_type = job < j->n_cat[CAT_ANNOTATION] ? CAT_ANNOTATION : CAT_REGION;
        const int ci = cat_type == CAT_ANNOTATION ? job : job - j->n_cat[CAT_ANNOTATION];
        double *(*tab)[4] = j->tab[cat_type];
        Visit visit;
        int ok = visit_build(j->d, j->wi, cat_type, ci, &visit);
        const size_t off = (size_t) ci * j->bins->n * TBL_STRIDE(j->n);
        for (int metric = 0; metric < 3 && ok; metric++) /* base_level before truth_based_auN, which divides by it */
            for (int cmp = 0; cmp < 4 && ok; cmp++) {
                if (!tab[metric][cmp]) continue;
                const int8_t *ref = (cmp == CMP_TRUTH_VS_PREDICTION || cmp == CMP_TRUTH) ? j->truth : j->prediction;
                const int8_t *query = (cmp == CMP_TRUTH_VS_PREDICTION || cmp == CMP_PREDICTION) ? j->prediction : j->truth;
                ok = scan_category(j->wi, &visit, ref, query, j->n, metric, j->overlap_threshold,
                                   metric == METRIC_AUN ? tab[METRIC_BASE][CMP_TRUTH] + off : NULL, j->bins, tab[metric][cmp] + off);
            }
        visit_free(&visit);
        if (!ok) {
            pthread_mutex_lock(&j->mu);
            j->oom = 1;
            pthread_mutex_unlock(&j->mu);
        }
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
    WinIndex wi;
    memset(&wi, 0, sizeof(wi));
    int oom = !vals || !winindex_build(d, prediction, truth, &wi);
    /* all tables first (the auN tables need the base_level truth table): one job per category, each building the list of
     * windows to visit once for its ten scans, on a few threads (the reference uses its thread pool here) */
    CatJobs jobs = {d, &wi, prediction, truth, &bins, n, {d->n_regions, d->n_annotations}, overlap_ratio_threshold, {{{NULL}}}, 0, 0,
                    PTHREAD_MUTEX_INITIALIZER};
    for (int cat_type = 0; cat_type < 2; cat_type++)
        for (int metric = 0; metric < 3; metric++)
            for (int cmp = 0; cmp < 4; cmp++) {
                const int need_truth = cmp != CMP_PREDICTION, need_pred = cmp != CMP_TRUTH;
                if ((need_truth && !truth) || (need_pred && !prediction)) continue;
                if (metric == METRIC_AUN && (cmp == CMP_PREDICTION || cmp == CMP_PREDICTION_VS_TRUTH)) continue;
                jobs.tab[cat_type][metric][cmp] = calloc((size_t) jobs.n_cat[cat_type] * bins.n * TBL_STRIDE(n), sizeof(double));
                if (!jobs.tab[cat_type][metric][cmp]) oom = 1;
            }
    if (!oom) {
        int n_threads = (int) sysconf(_SC_NPROCESSORS_ONLN);
        if (n_threads > 8) n_threads = 8;
        if (n_threads > jobs.n_cat[0] + jobs.n_cat[1]) n_threads = jobs.n_cat[0] + jobs.n_cat[1];
        pthread_t th[8];
        int started = 0;
        for (int t = 1; t < n_threads; t++)
            if (pthread_create(&th[started], NULL, cat_worker, &jobs) == 0) started++;
        cat_worker(&jobs); /* this thread works too, and alone if no thread could be started */
        for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
        oom = jobs.oom;
    }
    for (int cat_type = 0; cat_type < 2 && !oom; cat_type++) {
        const int n_cat = jobs.n_cat[cat_type];
        const char *const *cat_names = cat_type == CAT_ANNOTATION ? (const char *const *) d->annotation_names : NULL;
        double *(*tab)[4] = jobs.tab[cat_type];
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
    }
    for (int cat_type = 0; cat_type < 2; cat_type++)
        for (int metric = 0; metric < 3; metric++)
            for (int cmp = 0; cmp < 4; cmp++) free(jobs.tab[cat_type][metric][cmp]);
    free(vals);
    winindex_free(&wi);
    bins_free(&bins);
    fclose(f);
    if (f_stats) fclose(f_stats);
    if (f_aun) fclose(f_aun);
    if (oom) {
        snprintf(err, errlen, "hfg_write_summary_tsv: out of memory (or more than 2^31 windows)");
        return HFG_ERR_NOMEM;
    }
    return HFG_OK;
}

/* what hfg_benchmark_scores keeps with a data object between calls */
typedef struct ScoreCache {
    int ci;      /* annotation index the visit list was built for */
    WinIndex wi; /* its run_end belongs to no prediction (all-NULL labels) and is not used */
    Visit visit;
} ScoreCache;

static pthread_mutex_t g_score_cache_mu = PTHREAD_MUTEX_INITIALIZER;

void hfg_score_cache_free(void *cache) {
    ScoreCache *sc = cache;
    if (!sc) return;
    winindex_free(&sc->wi);
    visit_free(&sc->visit);
    free(sc);
}

/* a handful of independent scans, one thread each (the calling thread takes the first) */
typedef struct ScanJob {
    const WinIndex *w;
    const Visit *visit;
    const int8_t *ref, *query;
    int n, metric;
    double overlap_threshold;
    const double *aux;
    const SizeBins *bins;
    double *table;
    int ok;
} ScanJob;

static void *scan_job_run(void *arg) {
    ScanJob *j = arg;
    j->ok = scan_category(j->w, j->visit, j->ref, j->query, j->n, j->metric, j->overlap_threshold, j->aux, j->bins, j->table);
    return NULL;
}

#define MAX_SCAN_JOBS 8
static int run_scans(ScanJob *jobs, int count) {
    pthread_t th[MAX_SCAN_JOBS];
    int threaded[MAX_SCAN_JOBS] = {0};
    if (count > MAX_SCAN_JOBS) count = MAX_SCAN_JOBS; /* (callers pass 5 and 2) */
    for (int i = 1; i < count; i++) threaded[i] = pthread_create(&th[i], NULL, scan_job_run, &jobs[i]) == 0;
    for (int i = 0; i < count; i++)
        if (!threaded[i]) scan_job_run(&jobs[i]); /* job 0, and any job whose thread could not be started */
    int ok = 1;
    for (int i = 0; i < count; i++) {
        if (threaded[i]) pthread_join(th[i], NULL);
        ok &= jobs[i].ok;
    }
    return ok;
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
    double *tp[2] = {NULL, NULL}, *pt_[2] = {NULL, NULL}, *tt = calloc(stride, sizeof(double)),
           *aun_tp = calloc(stride, sizeof(double)), *aun_tt = calloc(stride, sizeof(double));
    /* the window coordinates and the windows of the annotation do not change between candidates of an alpha-tuning run:
     * kept with the data object (d->score_cache); only the run table depends on the prediction */
    WinIndex wi;
    Visit visit;
    memset(&wi, 0, sizeof(wi));
    memset(&visit, 0, sizeof(visit));
    int ok = tt && aun_tp && aun_tt;
    int32_t *run_end = NULL;
    if (ok) {
        pthread_mutex_lock(&g_score_cache_mu);
        ScoreCache *sc = ((hfg_cov_data *) d)->score_cache;
        if (!sc || sc->ci != ci || sc->wi.n != d->n_windows) {
            hfg_score_cache_free(sc);
            ((hfg_cov_data *) d)->score_cache = sc = calloc(1, sizeof(ScoreCache));
            if (sc) {
                sc->ci = ci;
                if (!winindex_build(d, NULL, NULL, &sc->wi) || !visit_build(d, &sc->wi, CAT_ANNOTATION, ci, &sc->visit)) {
                    hfg_score_cache_free(sc);
                    ((hfg_cov_data *) d)->score_cache = sc = NULL;
                }
            }
        }
        if (sc) {
            wi = sc->wi; /* shared, read-only; the run table below is this call's own */
            visit = sc->visit;
        }
        pthread_mutex_unlock(&g_score_cache_mu);
        run_end = sc ? malloc(sizeof(int32_t) * (size_t) (wi.n > 0 ? wi.n : 1)) : NULL;
        ok = sc && run_end;
        if (ok) {
            winindex_set_runs(&wi, prediction, truth, run_end);
            wi.run_end = run_end;
        }
    }
    for (int metric = 0; metric < 2; metric++) {
        tp[metric] = calloc(stride, sizeof(double));
        pt_[metric] = calloc(stride, sizeof(double));
        ok = ok && tp[metric] && pt_[metric];
    }
    if (ok) {
        /* five independent scans side by side, then the two auN scans, which divide by the base_level truth table */
        ScanJob jobs[7] = {
            {&wi, &visit, truth, prediction, n, METRIC_OVERLAP, overlap_ratio_threshold, NULL, &bins, tp[METRIC_OVERLAP], 0},
            {&wi, &visit, prediction, truth, n, METRIC_OVERLAP, overlap_ratio_threshold, NULL, &bins, pt_[METRIC_OVERLAP], 0},
            {&wi, &visit, truth, prediction, n, METRIC_BASE, overlap_ratio_threshold, NULL, &bins, tp[METRIC_BASE], 0},
            {&wi, &visit, prediction, truth, n, METRIC_BASE, overlap_ratio_threshold, NULL, &bins, pt_[METRIC_BASE], 0},
            {&wi, &visit, truth, truth, n, METRIC_BASE, overlap_ratio_threshold, NULL, &bins, tt, 0},
            {&wi, &visit, truth, prediction, n, METRIC_AUN, overlap_ratio_threshold, tt, &bins, aun_tp, 0},
            {&wi, &visit, truth, truth, n, METRIC_AUN, overlap_ratio_threshold, tt, &bins, aun_tt, 0}};
        ok = run_scans(jobs, 5) && run_scans(jobs + 5, 2);
    }
    free(run_end);
    if (!ok) {
        for (int metric = 0; metric < 2; metric++) {
            free(tp[metric]);
            free(pt_[metric]);
        }
        free(tt);
        free(aun_tp);
        free(aun_tt);
        bins_free(&bins);
        snprintf(err, errlen, "hfg_benchmark_scores: out of memory (or more than 2^31 windows)");
        return HFG_ERR_NOMEM;
    }
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
