/*
 * hfg_io.h -- the data formats on either side of the hot path: readers for the inputs
 * `hmm_flagger` accepts, and the summary-table writer on flat labels.  Plain C, no GPU involved.  The arrays a reader returns are exactly what hfg_set_chunks
 * (include/hfg.h) takes.
 *
 * Replaces, with the same window semantics:
 *   ChunksCreator_constructFromCov + ChunksCreator_parseChunks      programs/submodules/chunk/chunk.c:141-236,486-547
 *   TrackReader_readNextTrackCov, CoverageHeader_construct          programs/submodules/track_reader/track_reader.c:48-96,751-818
 *   Chunk_addTrack / Chunk_addWindow                                programs/submodules/chunk/chunk.c:393-481
 *   ChunksCreator_parseChunksFromBinaryFile                         programs/submodules/chunk/chunk.c:713-828
 */
#ifndef HFG_IO_H
#define HFG_IO_H

#include "hfg.h"

#ifdef __cplusplus
extern "C" {
#endif

#define HFG_CONTIG_NAME_MAX 200 /* Chunk.ctg, chunk.h:16 */

/* Everything a coverage input holds once cut into chunks of windows (CoverageHeader, track_reader.h:42-54, + the
 * ChunksCreator's chunk list, chunk.h:36-50), as flat arrays. */
typedef struct hfg_cov_data {
    int32_t n_annotations;
    char **annotation_names;
    int32_t n_regions;
    int32_t *region_coverages;
    int32_t n_labels, truth_available, prediction_available, start_only, avg_alignment_len;
    int32_t chunk_len, window_len;
    int32_t n_chunks;
    hfg_chunk_desc *chunks;                         /* [n_chunks], offsets into the window arrays */
    char (*contig_names)[HFG_CONTIG_NAME_MAX];      /* [n_chunks] */
    int64_t n_windows;
    uint16_t *cov, *cov_high_mapq, *cov_high_clip;  /* [n_windows] CoverageInfo.coverage* (ptBlock.h:79-92) */
    uint64_t *annotation_flag;                      /* [n_windows] incl. the region index in bits 58..63 */
    uint8_t *region;                                /* [n_windows] CoverageInfo_getRegionIndex */
    int8_t *truth, *prediction;                     /* [n_windows] Inference labels, -1 = none (ptBlock.h:50-55) */
    void *score_cache; /* owned by the library: what hfg_benchmark_scores keeps between calls (window coordinates, the
                          windows of the scored annotation); freed by hfg_cov_free.  Objects must come from the readers. */
} hfg_cov_data;

/* `.cov` or `.cov.gz` (run-length blocks tiling every contig) -> chunks of `chunk_len` bases cut into windows of
 * `window_len` bases.  One sequential pass, O(blocks + windows).  Returns hfg_status; on failure *out is NULL and err
 * holds a message. */
int hfg_read_cov(const char *path, int32_t chunk_len, int32_t window_len, hfg_cov_data **out, char *err, size_t errlen);

/* `.bin` chunk dump written by `hmm_flagger --dumpBin` (chunk.c:596-709); chunk and window lengths come from the file. */
int hfg_read_bin(const char *path, hfg_cov_data **out, char *err, size_t errlen);

void hfg_cov_free(hfg_cov_data *data);

/* ---- writers ------------------------------------------------------------------------------------------------------ */

/* prediction_summary_<suffix>.tsv (writeBenchmarkingStats, programs/src/hmm_flagger.c:134-161;
 * SummaryTableList_createAndWriteAllTables, programs/submodules/summary_table/summary_table.c:1663-1747) from FLAT label
 * arrays: per region and per annotation, the number of label blocks (overlap_based) and of bases (base_level) for every
 * label -- and, with truth labels, the confusion tables and the truth_based_auN metric -- counts then percentages, in the
 * reference's row order and formats.  prediction / truth: one int8 per window (-1 = none, counted as "Unk"); either may
 * be NULL (the comparisons that need it are left out).  When both are given, <path minus .tsv>.benchmarking.tsv
 * (precision / recall / F1 / accuracy) and .benchmarking.auN_ratio.tsv are written next to it, as the reference does.
 * label_names: n_labels + 1 names, the last one for "Unk" (NULL: label_0 ... label_unk).  bin_array_file: the size bins
 * of --binArrayFile ("start<TAB>end<TAB>name" lines; a block is counted in every bin with start <= length < end), or
 * NULL for the single bin ALL_SIZES. */
int hfg_write_summary_tsv(const char *path, const hfg_cov_data *data, const int8_t *prediction, const int8_t *truth,
                          const char *const *label_names, int n_labels, double overlap_ratio_threshold,
                          const char *bin_array_file, char *err, size_t errlen);

/* The three scores the reference's alpha-tuning driver reads back from the *.benchmarking*.tsv files of a run
 * (programs/src/tune_alpha_hmm_flagger.py:82-111): the F1-Score of the HARMONIC_MEAN_NO_HAP row of the overlap_based and
 * of the base_level table, and 100 x the HARMONIC_MEAN auN ratio, for one annotation and one size bin (default names
 * "whole_genome" / "ALL_SIZES"), each rounded to two decimals as the files print them (NaN where they print NA).
 * Computed from the flat label arrays without writing any file: what a tuning loop needs per candidate alpha matrix.
 * The window coordinates and the annotation's windows are kept with `data` after the first call (data->score_cache), so
 * later calls cost only the label-run scans (~6 ms at 750k windows); calls on one data object must not run concurrently. */
int hfg_benchmark_scores(const hfg_cov_data *data, const int8_t *prediction, const int8_t *truth, int n_labels,
                         double overlap_ratio_threshold, const char *bin_array_file, const char *annotation_label,
                         const char *size_label, double scores[3], char *err, size_t errlen);

/* Test hook: gunzip `path` with the reader's own gzip decoder (csrc/hfg_inflate.c) into out[cap], in pieces of piece_bytes
 * (0 = the reader's 4 MB); *len = bytes written.  0 on success, 1 with a message otherwise (not gzip, corrupt, CRC). */
int hfg_debug_gunzip(const char *path, uint8_t *out, size_t cap, size_t *len, size_t piece_bytes, char *err, size_t errlen);

#ifdef __cplusplus
}
#endif
#endif
