/*
 * hmm_flagger_b200.c -- a stand-alone `hmm_flagger` built only from this repository: the reference's command line
 * (programs/src/hmm_flagger.c:578-608, defaults :613-650, presets :21-58,518-575), its `.cov/.cov.gz/.bin` inputs
 * (include/hfg_io.h) and its per-run outputs, with the EM loop of runHMMFlagger (:285-488) driven through libhfg
 * (include/hfg.h).  Host code is plain C, like the reference.
 *
 * Outputs written into --outputDir (formats as the reference):
 *   loglikelihood.tsv, transition_{initial,iteration_k,final}.tsv, emission_{...}.tsv, final_flagger_prediction.bed,
 *   posterior_prediction_final.bed (-P), chunks.c_<C>.w_<W>.bin (-B), prediction_summary_{initial,iteration_k,final}.tsv
 *   and, for inputs with truth labels, their .benchmarking.tsv / .benchmarking.auN_ratio.tsv companions (all on the flat
 *   label array, hfg_write_summary_tsv; -k for every iteration).
 * --accelerate (SQUAREM) runs through hfg_squarem_iteration; --modelType negative_binomial through the same device-resident
 * loop as the other models (hfg_nb_dev.cuh; its blocking E-steps keep the host's libm table and estimator update).  Not supported: --initialRandomDev other than 0 (refused with a message).
 */
#include <getopt.h>
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/resource.h>
#include <sys/stat.h>
#include <sys/time.h>
#include <time.h>

#include "../../include/hfg.h"
#include "../../include/hfg_io.h"

static const char *STATE_NAMES[4] = {"Err", "Dup", "Hap", "Col"};
/* chunk.c:10-21 */
static const char *LABEL_NAMES[6] = {"Err", "Dup", "Hap", "Col", "Unk", "Msj"};
static const char *LABEL_COLORS[6] = {"162,0,37", "250,104,0", "0,138,0", "170,0,255", "99, 99, 96", "250,200,0"};

static const char *stamp(void) {
    static char buf[64];
    time_t t = time(NULL);
    strftime(buf, sizeof(buf), "%Y-%m-%d %H:%M:%S", localtime(&t));
    return buf;
}

/* the CUDA context (about a second on a B200 box) is created on a thread of its own while the input is read */
static void *warmup_thread(void *arg) {
    hfg_device_warmup(*(int *) arg);
    return NULL;
}

static double now_s(void) {
    struct timeval tv;
    gettimeofday(&tv, NULL);
    return tv.tv_sec + 1e-6 * tv.tv_usec;
}

static void die(const char *msg) {
    fprintf(stderr, "[%s] Error: %s\n", stamp(), msg);
    exit(EXIT_FAILURE);
}

static void *xmalloc(size_t bytes) {
    void *p = malloc(bytes ? bytes : 1);
    if (!p) die("out of host memory");
    return p;
}

static int is_gauss(const hfg_config *cfg, int s) {
    return !(cfg->model_type == HFG_MODEL_TRUNC_EXP_GAUSSIAN && s == HFG_STATE_ERR);
}

/* HMM_printTransitionMatrixInTsvFormat, hmm.c:137-181 */
static void write_transition_tsv(const char *dir, const char *suffix, const hfg_config *cfg, const hfg_region_params *p) {
    char path[4096];
    snprintf(path, sizeof(path), "%s/transition_%s.tsv", dir, suffix);
    FILE *f = fopen(path, "w+");
    if (!f) die("cannot write the transition tsv");
    fprintf(f, "#Region\tState\tErr\tDup\tHap\tCol\tEnd\n");
    for (int r = 0; r < cfg->n_regions; r++)
        for (int pre = 0; pre <= HFG_NUM_STATES; pre++) {
            fprintf(f, "%d\t%s", r, pre < HFG_NUM_STATES ? STATE_NAMES[pre] : "Start");
            for (int s = 0; s <= HFG_NUM_STATES; s++) fprintf(f, "\t%.5e", p[r].trans[pre][s]);
            fprintf(f, "\n");
        }
    fclose(f);
}

/* HMM_printEmissionParametersInTsvFormat, hmm.c:183-239; parameter lists hmm_utils.c:1539-1576 */
static void write_emission_tsv(const char *dir, const char *suffix, const hfg_config *cfg, const hfg_region_params *p) {
    char path[4096];
    snprintf(path, sizeof(path), "%s/emission_%s.tsv", dir, suffix);
    FILE *f = fopen(path, "w+");
    if (!f) die("cannot write the emission tsv");
    fprintf(f, "#State\tDistribution\tComponents\tParameter");
    for (int r = 0; r < cfg->n_regions; r++) fprintf(f, "\tValues_Region_%d", r);
    fprintf(f, "\n");
    for (int s = 0; s < HFG_NUM_STATES; s++) {
        if (!is_gauss(cfg, s)) {
            static const char *names[2] = {"Mean", "Trunc_Point"};
            for (int k = 0; k < 2; k++) {
                fprintf(f, "%s\tTruncated Exponential\t1\t%s", STATE_NAMES[s], names[k]);
                for (int r = 0; r < cfg->n_regions; r++) fprintf(f, "\t%.5e", k == 0 ? 1.0 / p[r].lambda : p[r].trunc_point);
                fprintf(f, "\n");
            }
        } else {
            static const char *names[3] = {"Mean", "Var", "Weight"};
            const int nb = cfg->model_type == HFG_MODEL_NEGATIVE_BINOMIAL;
            for (int k = 0; k < 3; k++) {
                fprintf(f, "%s\t%s\t%d\t%s", STATE_NAMES[s], nb ? "Negative Binomial" : "Gaussian", cfg->n_comps[s], names[k]);
                for (int r = 0; r < cfg->n_regions; r++) {
                    const double *v = k == 0 ? p[r].mean[s] : (k == 1 ? p[r].var[s] : p[r].weight[s]);
                    fprintf(f, "\t");
                    for (int c = 0; c < cfg->n_comps[s]; c++) {
                        double value = v[c];
                        if (nb && k < 2) {
                            /* the table shows mean and variance, the model keeps (theta, lambda) in those slots
                             * (NegativeBinomial_getMean / _getVar, hmm_utils.c:460-470) */
                            const double theta = p[r].mean[s][c], rr = -1 * p[r].var[s][c] / log(theta);
                            value = k == 0 ? rr * (1 - theta) / theta : rr * (1 - theta) / pow(theta, 2);
                        }
                        fprintf(f, c ? ",%.5e" : "%.5e", value);
                    }
                }
                fprintf(f, "\n");
            }
        }
    }
    fclose(f);
}

typedef struct Block {
    int s, e, label;
} Block;

/* flush the blocks of one contig: merge neighbours with equal labels (mergeBlocksWithSameLabels, chunk.c:949-983 --
 * including its quirk that the first merged block starts at 0) and print them (chunk.c:1053-1069) */
static void flush_contig(FILE *f, const char *ctg, const Block *b, int n) {
    if (n == 0) return;
    int pre_label = -1, pre_start = 0, pre_end = 0;
    for (int i = 0; i < n; i++) {
        if (pre_label != -1 && b[i].label != pre_label) {
            fprintf(f, "%s\t%d\t%d\t%s\t0\t.\t%d\t%d\t%s\n", ctg, pre_start, pre_end + 1, LABEL_NAMES[pre_label], pre_start,
                    pre_end + 1, LABEL_COLORS[pre_label]);
            pre_start = b[i].s;
        }
        pre_end = b[i].e;
        pre_label = b[i].label;
    }
    fprintf(f, "%s\t%d\t%d\t%s\t0\t.\t%d\t%d\t%s\n", ctg, pre_start, pre_end + 1, LABEL_NAMES[pre_label], pre_start,
            pre_end + 1, LABEL_COLORS[pre_label]);
}

/* ChunksCreator_writePredictionIntoFinalBED, chunk.c:985-1124 */
static void write_final_bed(const char *path, const char *track, const hfg_cov_data *d, const int8_t *labels,
                            const int *min_len) {
    FILE *f = fopen(path, "w");
    if (!f) die("cannot write the final BED");
    fprintf(f, "track name=%s visibility=1 itemRgb=\"On\"\n", track);
    Block *blocks = xmalloc(sizeof(Block) * (size_t) (d->n_windows + 1));
    int nb = 0, bed_start = 0, pre_end = 0, pre_label = -1;
    const char *pre_ctg = NULL;
    for (int c = 0; c < d->n_chunks; c++) {
        const hfg_chunk_desc *ch = &d->chunks[c];
        const char *ctg = d->contig_names[c];
        for (int i = 0; i < ch->n_windows; i++) {
            const int start = ch->s + i * ch->window_len;
            int end = ch->s + (i + 1) * ch->window_len - 1;
            if (end > ch->e) end = ch->e;
            int label = labels[ch->offset + i];
            if (label < 0) label = 4; /* "Unk" */
            if (pre_label == -1 || pre_ctg == NULL) bed_start = start;
            const int label_changed = pre_label != -1 && label != pre_label;
            const int ctg_changed = pre_ctg != NULL && strcmp(pre_ctg, ctg) != 0;
            if (label_changed || ctg_changed) {
                const int len = pre_end + 1 - bed_start;
                blocks[nb].s = bed_start;
                blocks[nb].e = pre_end;
                blocks[nb].label = (pre_label < 4 && len < min_len[pre_label]) ? 2 : pre_label; /* short -> Hap */
                nb++;
                bed_start = start;
            }
            if (ctg_changed) {
                flush_contig(f, pre_ctg, blocks, nb);
                nb = 0;
            }
            pre_end = end;
            pre_label = label;
            pre_ctg = ctg;
        }
    }
    if (pre_label != -1) {
        const int len = pre_end + 1 - bed_start;
        blocks[nb].s = bed_start;
        blocks[nb].e = pre_end;
        blocks[nb].label = (pre_label < 4 && len < min_len[pre_label]) ? 2 : pre_label;
        nb++;
        flush_contig(f, pre_ctg, blocks, nb);
    }
    free(blocks);
    fclose(f);
}

/* writePosteriorIntoBED, src/hmm_flagger.c:240-282 */
static void write_posterior_bed(const char *dir, const hfg_cov_data *d, const double *post, const int8_t *labels) {
    char path[4096];
    snprintf(path, sizeof(path), "%s/posterior_prediction_final.bed", dir);
    FILE *f = fopen(path, "w+");
    if (!f) die("cannot write the posterior BED");
    fprintf(f, "#ctg\tstart\tend\t");
    for (int s = 0; s < HFG_NUM_STATES; s++) fprintf(f, "posterior_%s_%d\t", STATE_NAMES[s], s);
    fprintf(f, "prediction\n");
    for (int c = 0; c < d->n_chunks; c++) {
        const hfg_chunk_desc *ch = &d->chunks[c];
        for (int i = 0; i < ch->n_windows; i++) {
            const int start = ch->s + i * ch->window_len;
            int end = ch->s + (i + 1) * ch->window_len - 1;
            if (end > ch->e) end = ch->e;
            fprintf(f, "%s\t%d\t%d\t", d->contig_names[c], start, end + 1);
            for (int s = 0; s < HFG_NUM_STATES; s++) fprintf(f, "%.2f\t", post[(ch->offset + i) * 4 + s]);
            fprintf(f, "%s\n", STATE_NAMES[labels[ch->offset + i]]);
        }
    }
    fclose(f);
}

/* ChunksCreator_writeChunksIntoBinaryFile, chunk.c:596-709 */
static void write_bin(const char *path, const hfg_cov_data *d) {
    FILE *f = fopen(path, "wb+");
    if (!f) die("cannot write the bin file");
    fwrite(&d->n_annotations, 4, 1, f);
    for (int i = 0; i < d->n_annotations; i++) {
        int32_t n = (int32_t) strlen(d->annotation_names[i]) + 1;
        fwrite(&n, 4, 1, f);
        fwrite(d->annotation_names[i], 1, (size_t) n, f);
    }
    fwrite(&d->n_regions, 4, 1, f);
    fwrite(d->region_coverages, 4, (size_t) d->n_regions, f);
    fwrite(&d->n_labels, 4, 1, f);
    uint8_t b[3] = {(uint8_t) d->truth_available, (uint8_t) d->prediction_available, (uint8_t) d->start_only};
    fwrite(b, 1, 3, f);
    fwrite(&d->avg_alignment_len, 4, 1, f);
    fwrite(&d->chunk_len, 4, 1, f);
    fwrite(&d->window_len, 4, 1, f);
    for (int c = 0; c < d->n_chunks; c++) {
        const hfg_chunk_desc *ch = &d->chunks[c];
        int32_t n = (int32_t) strlen(d->contig_names[c]) + 1;
        fwrite(&n, 4, 1, f);
        fwrite(d->contig_names[c], 1, (size_t) n, f);
        fwrite(&ch->ctg_len, 4, 1, f);
        fwrite(&ch->s, 4, 1, f);
        fwrite(&ch->e, 4, 1, f);
        fwrite(&ch->n_windows, 4, 1, f);
        const int64_t o = ch->offset;
        const size_t L = (size_t) ch->n_windows;
        fwrite(d->cov + o, 2, L, f);
        fwrite(d->cov_high_mapq + o, 2, L, f);
        fwrite(d->cov_high_clip + o, 2, L, f);
        fwrite(d->annotation_flag + o, 8, L, f);
        fwrite(d->truth + o, 1, L, f);
        fwrite(d->prediction + o, 1, L, f);
    }
    fclose(f);
}

/* ChunksCreator_subsetChunksToContigs: keep only the chunks whose contig is listed (one name per line) */
static void subset_contigs(hfg_cov_data *d, const char *list_path) {
    FILE *f = fopen(list_path, "r");
    if (!f) die("cannot open the contigs list");
    char **names = NULL;
    int n = 0;
    char line[1024];
    while (fgets(line, sizeof(line), f)) {
        line[strcspn(line, "\r\n")] = '\0';
        if (!line[0]) continue;
        names = realloc(names, sizeof(char *) * (size_t) (n + 1));
        if (!names) die("out of host memory");
        names[n++] = strdup(line);
    }
    fclose(f);
    int out = 0;
    int64_t w = 0;
    for (int c = 0; c < d->n_chunks; c++) {
        int keep = 0;
        for (int i = 0; i < n && !keep; i++) keep = strcmp(names[i], d->contig_names[c]) == 0;
        if (!keep) continue;
        const int64_t o = d->chunks[c].offset;
        const size_t L = (size_t) d->chunks[c].n_windows;
        memmove(d->cov + w, d->cov + o, 2 * L);
        memmove(d->cov_high_mapq + w, d->cov_high_mapq + o, 2 * L);
        memmove(d->cov_high_clip + w, d->cov_high_clip + o, 2 * L);
        memmove(d->annotation_flag + w, d->annotation_flag + o, 8 * L);
        memmove(d->region + w, d->region + o, L);
        memmove(d->truth + w, d->truth + o, L);
        memmove(d->prediction + w, d->prediction + o, L);
        d->chunks[out] = d->chunks[c];
        d->chunks[out].offset = w;
        memmove(d->contig_names[out], d->contig_names[c], HFG_CONTIG_NAME_MAX);
        w += (int64_t) L;
        out++;
    }
    d->n_chunks = out;
    d->n_windows = w;
    for (int i = 0; i < n; i++) free(names[i]);
    free(names);
}

static int ends_with(const char *s, const char *suffix) {
    const size_t a = strlen(s), b = strlen(suffix);
    return a >= b && strcmp(s + a - b, suffix) == 0;
}

/* writeBenchmarkingStats (src/hmm_flagger.c:134-161) on the flat label array */
static void write_summary(const char *dir, const char *suffix, const hfg_cov_data *d, const int8_t *labels,
                          char *const *label_names, double overlap_thr, const char *bin_array_file) {
    char path[4096], err[512];
    snprintf(path, sizeof(path), "%s/prediction_summary_%s.tsv", dir, suffix);
    if (hfg_write_summary_tsv(path, d, labels, d->truth_available ? d->truth : NULL, (const char *const *) label_names,
                              HFG_NUM_STATES, overlap_thr, bin_array_file, err, sizeof(err)) != HFG_OK)
        die(err);
}

static struct option long_options[] = {{"input", required_argument, NULL, 'i'},
                                       {"preset", required_argument, NULL, 'x'},
                                       {"iterations", required_argument, NULL, 'n'},
                                       {"convergenceTol", required_argument, NULL, 't'},
                                       {"disableAdjustContigEnds", no_argument, NULL, 'e'},
                                       {"minReadFractionAtEnds", required_argument, NULL, 'f'},
                                       {"modelType", required_argument, NULL, 'm'},
                                       {"maxHighMapqRatio", required_argument, NULL, 'q'},
                                       {"minHighMapqRatio", required_argument, NULL, 'Q'},
                                       {"chunkLen", required_argument, NULL, 'C'},
                                       {"windowLen", required_argument, NULL, 'W'},
                                       {"contigsList", required_argument, NULL, 'c'},
                                       {"threads", required_argument, NULL, '@'},
                                       {"collapsedComps", required_argument, NULL, 'p'},
                                       {"alphaTsv", required_argument, NULL, 'A'},
                                       {"binArrayFile", required_argument, NULL, 'a'},
                                       {"writeParameterStatsPerIteration", no_argument, NULL, 'w'},
                                       {"writeBenchmarkingStatsPerIteration", no_argument, NULL, 'k'},
                                       {"writePosteriorProbs", no_argument, NULL, 'P'},
                                       {"outputDir", required_argument, NULL, 'o'},
                                       {"overlapRatioThreshold", required_argument, NULL, 'v'},
                                       {"labelNames", required_argument, NULL, 'l'},
                                       {"initialRandomDev", required_argument, NULL, 'D'},
                                       {"trackName", required_argument, NULL, 'N'},
                                       {"dumpBin", no_argument, NULL, 'B'},
                                       {"accelerate", no_argument, NULL, 's'},
                                       {"minimumLengths", required_argument, NULL, 'M'},
                                       {"device", required_argument, NULL, 'g'},
                                       {NULL, 0, NULL, 0}};

int main(int argc, char *argv[]) {
    const char *track = "final_hmm_flagger", *preset = "hifi", *input = NULL, *alpha_tsv = NULL, *contigs = NULL, *out_dir = NULL;
    int iterations = 100, adjust_ends = 1, collapsed = -1, write_params = 0, write_post = 0, chunk_len = 20000000;
    int window_len = -1, dump_bin = 0, device = 0, model_type = -1, accelerate = 0, write_bench = 0;
    double overlap_thr = 0.4; /* --overlapRatioThreshold (src/hmm_flagger.c:632) */
    char *label_names[64];
    int n_label_names = 0;
    const char *bin_array_file = NULL;
    double tol = 0.001, max_mapq = 0.25, min_mapq = 0.75, min_frac = -1.0;
    int min_len[4] = {0, 0, 0, 0};
    int c;
    while (~(c = getopt_long(argc, argv, "i:x:f:en:t:m:q:Q:C:W:c:@:p:A:a:wkPo:v:l:D:BN:M:sg:", long_options, NULL))) {
        switch (c) {
            case 'i': input = optarg; break;
            case 'x': preset = optarg; break;
            case 'n': iterations = atoi(optarg); break;
            case 'B': dump_bin = 1; break;
            case 'N': track = optarg; break;
            case 't': tol = atof(optarg); break;
            case 'e': adjust_ends = 0; break;
            case 'f': min_frac = atof(optarg); break;
            case 'm':
                if (strcmp(optarg, "trunc_exp_gaussian") == 0) model_type = HFG_MODEL_TRUNC_EXP_GAUSSIAN;
                else if (strcmp(optarg, "gaussian") == 0) model_type = HFG_MODEL_GAUSSIAN;
                else if (strcmp(optarg, "negative_binomial") == 0) model_type = HFG_MODEL_NEGATIVE_BINOMIAL;
                else die("--modelType should be trunc_exp_gaussian, gaussian or negative_binomial");
                break;
            case 'c': contigs = optarg; break;
            case 'p': collapsed = atoi(optarg); break;
            case 'A': alpha_tsv = optarg; break;
            case 'C': chunk_len = atoi(optarg); break;
            case 'W': window_len = atoi(optarg); break;
            case 'w': write_params = 1; break;
            case 'P': write_post = 1; break;
            case 'o': out_dir = optarg; break;
            case 'q': max_mapq = atof(optarg); break;
            case 'Q': min_mapq = atof(optarg); break;
            case 'g': device = atoi(optarg); break;
            case 'M': {
                int a, b, d3;
                if (sscanf(optarg, "%d,%d,%d", &a, &b, &d3) != 3) die("--minimumLengths should contain 3 comma-delimited integers");
                min_len[0] = a; min_len[1] = b; min_len[3] = d3; /* Err, Dup, Col (src/hmm_flagger.c:744-746) */
                break;
            }
            case 's': accelerate = 1; break;
            case 'k': write_bench = 1; break;
            case 'v': overlap_thr = atof(optarg); break;
            case 'a': bin_array_file = optarg; break;
            case 'l': {
                /* --labelNames: comma-separated, "Unk" appended (src/hmm_flagger.c:721-724) */
                char *copy = strdup(optarg);
                for (char *tok = strtok(copy, ","); tok && n_label_names < 62; tok = strtok(NULL, ",")) label_names[n_label_names++] = tok;
                label_names[n_label_names++] = "Unk";
                break;
            }
            case '@':
                break; /* accepted for command-line compatibility: there is no thread pool here */
            case 'D':
                /* the reference perturbs the initial means with rand() (src/hmm_flagger.c:207-219): not reproduced here */
                if (atof(optarg) != 0.0) die("--initialRandomDev other than 0 is not supported by this binary");
                break;
            default:
                fprintf(stderr, "Usage: %s -i <INPUT.cov|.cov.gz|.bin> -o <OUTPUT_DIR> [options of hmm_flagger v1.2.0] [--device N]\n", argv[0]);
                return 1;
        }
    }
    const double t_start = now_s();
    if (!input) die("Input path cannot be NULL.");
    /* one visible device: the driver then initialises that GPU only (an 8-GPU box otherwise pays for all eight) */
    static int warm_device = 0;
    if (!getenv("CUDA_VISIBLE_DEVICES")) {
        char dev[16];
        snprintf(dev, sizeof(dev), "%d", device);
        setenv("CUDA_VISIBLE_DEVICES", dev, 1);
        device = 0;
    }
    warm_device = device;
    pthread_t warm_tid;
    const int warm_started = pthread_create(&warm_tid, NULL, warmup_thread, &warm_device) == 0;
    if (n_label_names && n_label_names - 1 != HFG_NUM_STATES)
        die("Number of label names does not match the number of labels (4: Err,Dup,Hap,Col)."); /* summary_table.c:1682-1689 */
    if (tol <= 0.0 || tol > 1.0) die("convergence tol should be between 0 and 1.");
    struct stat st;
    if (!out_dir) die("--outputDir, -o should be specified.");
    if (stat(out_dir, &st) != 0 || !S_ISDIR(st.st_mode)) die("Output directory does not exist!");

    /* presets (src/hmm_flagger.c:21-58): window size, minReadFraction and model type.  The preset alpha arrays are
     * declared `int[4][4]` in the reference, so a preset WITHOUT --alphaTsv yields an all-zero alpha -- mirrored. */
    int p_window;
    double p_frac;
    if (strcmp(preset, "hifi") == 0) { p_window = 16000; p_frac = 0.95; }
    else if (strcmp(preset, "ont-r9") == 0) { p_window = 16000; p_frac = 1.0; }
    else if (strcmp(preset, "ont-r10") == 0) { p_window = 8000; p_frac = 0.8; }
    else die("preset can be one of hifi, ont-r9, ont-r10.");
    double alpha[16];
    memset(alpha, 0, sizeof(alpha));
    if (alpha_tsv) { /* MatrixDouble_parseFromFile (data_types.c:490-518) + range check (src/hmm_flagger.c:503-512) */
        FILE *f = fopen(alpha_tsv, "r");
        if (!f) die("cannot open the alpha tsv");
        for (int i = 0; i < 16; i++)
            if (fscanf(f, "%lf", &alpha[i]) != 1 || alpha[i] < 0.0 || alpha[i] > 1.0) die("alpha tsv: 4x4 values between 0 and 1 expected");
        fclose(f);
    }
    if (min_frac < 0.0 && adjust_ends) min_frac = p_frac;
    if (window_len < 0) window_len = p_window;
    if (model_type < 0) model_type = HFG_MODEL_TRUNC_EXP_GAUSSIAN;
    if (adjust_ends && (min_frac > 1.0 || min_frac < 0.0)) die("--minReadFractionAtEnds, -f should be between 0 and 1.");
    if (window_len <= 0) die("windowLen cannot be <= 0.");

    /* 1. chunks */
    fprintf(stderr, "[%s] Parsing/Creating coverage chunks. \n", stamp());
    /* where the wall time goes (printed behind the reference's "Real time" line): read, GPU set-up, EM, summary tables, BED */
    double ph_read = 0, ph_setup = 0, ph_summary = 0, ph_bed = 0, ph_t = now_s();
    hfg_cov_data *d = NULL;
    char err[512] = "";
    int rc;
    if (ends_with(input, ".bin")) rc = hfg_read_bin(input, &d, err, sizeof(err));
    else if (ends_with(input, ".cov") || ends_with(input, ".cov.gz")) rc = hfg_read_cov(input, chunk_len, window_len, &d, err, sizeof(err));
    else die("input file should either cov/cov.gz or a binary file made with --dumpBin.");
    if (rc != HFG_OK) die(err);
    if (contigs) subset_contigs(d, contigs);
    if (d->n_chunks == 0) die("no chunks to process");
    if (dump_bin) {
        char path[4096];
        snprintf(path, sizeof(path), "%s/chunks.c_%d.w_%d.bin", out_dir, d->chunk_len, d->window_len);
        write_bin(path, d);
    }
    fprintf(stderr, "[%s] %d chunks are parsed (%lld windows). \n", stamp(), d->n_chunks, (long long) d->n_windows);
    ph_read = now_s() - ph_t;

    /* 2. number of collapsed components (src/hmm_flagger.c:1003-1023) */
    if (collapsed == -1) {
        int max_cov = 0;
        for (int64_t i = 0; i < d->n_windows; i++) max_cov = d->cov[i] > max_cov ? d->cov[i] : max_cov;
        collapsed = hfg_best_num_collapsed_comps(max_cov, d->region_coverages, d->n_regions);
    }

    /* 3. model */
    hfg_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.model_type = model_type;
    cfg.n_regions = d->n_regions;
    cfg.n_comps[0] = cfg.n_comps[1] = cfg.n_comps[2] = 1;
    cfg.n_comps[3] = collapsed;
    cfg.adjust_contig_ends = adjust_ends;
    cfg.mean_read_length = d->avg_alignment_len;
    cfg.min_read_fraction_at_ends = adjust_ends ? min_frac : 0.0;
    cfg.max_high_mapq_ratio = max_mapq;
    cfg.min_high_mapq_ratio = min_mapq;
    cfg.min_highly_clipped_ratio = 1.0; /* src/hmm_flagger.c:222 */
    cfg.device = device;
    hfg_region_params *params = xmalloc(sizeof(hfg_region_params) * (size_t) cfg.n_regions);
    hfg_region_stats *stats = xmalloc(sizeof(hfg_region_stats) * (size_t) cfg.n_regions);
    if (hfg_model_init(&cfg, d->region_coverages, d->window_len, d->start_only, params) != HFG_OK) die("invalid model configuration");

    /* 4. EM (runHMMFlagger, src/hmm_flagger.c:285-488) */
    fprintf(stderr, "[%s] Running EM for estimating parameters. \n", stamp());
    char path[4096], suffix[64];
    snprintf(path, sizeof(path), "%s/loglikelihood.tsv", out_dir);
    FILE *ll_file = fopen(path, "w+");
    if (!ll_file) die("cannot write loglikelihood.tsv");
    fprintf(ll_file, "#Iteration\tEffective_Iteration\tLoglikelihood\n");
    write_transition_tsv(out_dir, "initial", &cfg, params);
    write_emission_tsv(out_dir, "initial", &cfg, params);
    /* the GPU context: no CPU fallback -- without a usable device the run stops here */
    hfg_ctx *ctx = NULL;
    ph_t = now_s();
    if (warm_started) pthread_join(warm_tid, NULL);
    if (hfg_create(&ctx, &cfg) != HFG_OK) die(hfg_last_error(NULL));
    if (hfg_set_chunks(ctx, d->n_chunks, d->chunks, d->cov, d->cov_high_mapq, d->cov_high_clip, d->region) != HFG_OK)
        die(hfg_last_error(ctx));
    ph_setup = now_s() - ph_t;
    const double ph_em0 = now_s();
    int8_t *labels = xmalloc((size_t) d->n_windows);
    int iter = 1, converged = 0, final_done = 0;
    double loglik = 0.0;
    hfg_region_params *params_before = xmalloc(sizeof(hfg_region_params) * (size_t) cfg.n_regions);
    hfg_region_stats *stats_scratch = xmalloc(sizeof(hfg_region_stats) * (size_t) cfg.n_regions);
    /* Iterations whose results are needed on the host between E-steps run through the blocking call + host M-step: the first
     * one always (its labels are prediction_summary_initial.tsv, src/hmm_flagger.c:361-379), all of them with
     * --accelerate, -w or -k.  Otherwise the REST of the loop -- every further E-step, the M-steps, the convergence test and
     * the final inference -- is queued on the device at once (hfg_em_*: parameters stay in HBM, the M-step runs in the tail
     * of the E-step kernel) and the host comes back for the log-likelihoods, the parameters and the labels. */
    const int per_iteration_outputs = accelerate || write_params || write_bench;
    while (iter <= iterations && !converged) {
        if (!per_iteration_outputs && iter > 1 && iterations - iter + 2 <= 4096) {
            const int remaining = iterations - iter + 1;
            double *lls = xmalloc(sizeof(double) * ((size_t) remaining + 1));
            int n_esteps = 0, rc_d = hfg_em_begin(ctx, alpha, params, tol, remaining + 1);
            for (int it = 0; it < remaining && rc_d == HFG_OK; it++) rc_d = hfg_em_enqueue(ctx, 0);
            if (rc_d == HFG_OK) rc_d = hfg_em_enqueue(ctx, 1);
            if (rc_d == HFG_OK) rc_d = hfg_em_finish(ctx, params, lls, &n_esteps, &converged, labels);
            if (rc_d != HFG_OK) {
                fprintf(stderr, "%s\n", hfg_last_error(ctx));
                exit(EXIT_FAILURE);
            }
            for (int k = 0; k < n_esteps; k++) fprintf(ll_file, "%d\t%d\t%.4f\n", iter - 1 + k, iter - 1 + k, lls[k]);
            iter += n_esteps - 1; /* n_esteps - 1 EM iterations, then the final inference */
            free(lls);
            final_done = 1;
            break;
        }
        /* --accelerate: E(p0), M, E(p1), M, SQUAREM candidate p' chosen with forward-only passes, E(p') (:382-416) */
        double rate = 0.0;
        const int want_labels = iter == 1 || write_bench;
        if (accelerate && want_labels) memcpy(params_before, params, sizeof(hfg_region_params) * (size_t) cfg.n_regions);
        const int rc_e = accelerate ? hfg_squarem_iteration(ctx, alpha, params, stats, tol, &loglik, &rate)
                                    : hfg_em_iteration(ctx, alpha, params, stats, &loglik, want_labels ? labels : NULL);
        if (rc_e != HFG_OK) {
            fprintf(stderr, "%s\n", hfg_last_error(ctx));
            exit(EXIT_FAILURE);
        }
        if (accelerate) fprintf(stderr, "[%s] Computed alpha rate for accelerating EM = %.4f\n", stamp(), rate);
        fprintf(ll_file, "%d\t%d\t%.4f\n", iter - 1, accelerate ? 3 * (iter - 1) : iter - 1, loglik);
        if (want_labels) {
            /* the summary holds the labels of the E-step with the iteration's STARTING parameters, written before any
             * acceleration (src/hmm_flagger.c:344-379): with --accelerate that pass ran inside hfg_squarem_iteration
             * without keeping its labels, so it is repeated here (first iteration, or every one with -k) */
            if (accelerate) {
                double ll0;
                if (hfg_em_iteration(ctx, alpha, params_before, stats_scratch, &ll0, labels) != HFG_OK) die(hfg_last_error(ctx));
            }
            if (iter == 1) snprintf(suffix, sizeof(suffix), "initial");
            else snprintf(suffix, sizeof(suffix), accelerate ? "iteration_accelerated_%d" : "iteration_%d", iter - 1);
            ph_t = now_s();
            write_summary(out_dir, suffix, d, labels, n_label_names ? label_names : NULL, overlap_thr, bin_array_file);
            ph_summary += now_s() - ph_t;
        }
        hfg_mstep(&cfg, params, stats, tol, &converged);
        if (write_params) {
            snprintf(suffix, sizeof(suffix), accelerate ? "iteration_accelerated_%d" : "iteration_%d", iter);
            write_transition_tsv(out_dir, suffix, &cfg, params);
            write_emission_tsv(out_dir, suffix, &cfg, params);
        }
        iter++;
    }
    if (converged) fprintf(stderr, "[%s] Parameters converged after %d iterations (tol=%.2e)\n", stamp(), iter - 1, tol);
    else fprintf(stderr, "[%s] Parameter estimation stopped (not yet converged based on the given tolerance) after %d iterations (tol=%.2e)\n", stamp(), iter - 1, tol);
    /* final inference with the final parameters (:464) */
    if (!final_done) {
        if (hfg_em_iteration(ctx, alpha, params, stats, &loglik, labels) != HFG_OK) {
            fprintf(stderr, "%s\n", hfg_last_error(ctx));
            exit(EXIT_FAILURE);
        }
        fprintf(ll_file, "%d\t%d\t%.4f\n", iter - 1, accelerate ? 3 * (iter - 1) : iter - 1, loglik);
    }
    const double ph_em = now_s() - ph_em0 - ph_summary;
    ph_t = now_s();
    write_summary(out_dir, "final", d, labels, n_label_names ? label_names : NULL, overlap_thr, bin_array_file);
    ph_summary += now_s() - ph_t;
    fclose(ll_file);
    write_transition_tsv(out_dir, "final", &cfg, params);
    write_emission_tsv(out_dir, "final", &cfg, params);
    if (write_post) {
        double *post = xmalloc(sizeof(double) * 4 * (size_t) d->n_windows);
        if (hfg_get_posteriors(ctx, post) != HFG_OK) die(hfg_last_error(ctx));
        write_posterior_bed(out_dir, d, post, labels);
        free(post);
    }

    /* 5. final BED */
    fprintf(stderr, "[%s] Writing final BED file. \n", stamp());
    snprintf(path, sizeof(path), "%s/final_flagger_prediction.bed", out_dir);
    ph_t = now_s();
    write_final_bed(path, track, d, labels, min_len);
    ph_bed = now_s() - ph_t;

    hfg_destroy(ctx);
    hfg_cov_free(d);
    free(params);
    free(params_before);
    free(stats);
    free(stats_scratch);
    free(labels);
    fprintf(stderr, "[%s] Done! \n", stamp());
    struct rusage ru;
    getrusage(RUSAGE_SELF, &ru);
    const double real = now_s() - t_start;
    const double cpu = ru.ru_utime.tv_sec + ru.ru_stime.tv_sec + 1e-6 * (ru.ru_utime.tv_usec + ru.ru_stime.tv_usec);
    fprintf(stderr, "Real time:  %.3f sec; CPU: %.3f sec; Peak RSS: %.3f GB; CPU usage: %.1f%%\n", real, cpu,
            ru.ru_maxrss / 1024.0 / 1024.0, 100.0 * cpu / (real > 0 ? real : 1));
    fprintf(stderr, "Phases: read %.4f s; GPU set-up %.4f s; EM %.4f s; summary tables %.4f s; BED %.4f s\n", ph_read, ph_setup, ph_em,
            ph_summary, ph_bed);
    return 0;
}
