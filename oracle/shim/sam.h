/*
 * oracle/shim/sam.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Declaration-only stand-in for htslib 1.13 (reference Dockerfile:27-36), which is
 * absent from this image.  The hmm_flagger call graph never reads a BAM; these
 * declarations exist so that the reference's ptBlock/ptAlignment/cigar_it
 * translation units COMPILE and LINK unmodified.  Every function aborts if reached.
 * Struct layouts follow the public htslib API only as far as the reference
 * dereferences them.
 */
#ifndef HFG_ORACLE_SAM_STUB_H
#define HFG_ORACLE_SAM_STUB_H
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

typedef int64_t hts_pos_t;

typedef struct bam1_core_t {
    hts_pos_t pos;
    int32_t tid;
    uint16_t bin;
    uint8_t qual;
    uint8_t l_extranul;
    uint16_t flag;
    uint16_t l_qname;
    uint32_t n_cigar;
    int32_t l_qseq;
    int32_t mtid;
    hts_pos_t mpos;
    hts_pos_t isize;
} bam1_core_t;

typedef struct bam1_t {
    bam1_core_t core;
    uint64_t id;
    uint8_t *data;
    int l_data;
    uint32_t m_data;
    uint32_t mempolicy;
} bam1_t;

typedef struct sam_hdr_t {
    int32_t n_targets, ignore_sam_err;
    size_t l_text;
    uint32_t *target_len;
    const int8_t *cigar_tab;
    char **target_name;
    char *text;
    void *sdict;
    void *hrecs;
    uint32_t ref_count;
} sam_hdr_t;
typedef sam_hdr_t bam_hdr_t;

typedef struct htsFile { int dummy; } htsFile;
typedef htsFile samFile;
typedef struct hts_idx_t { int dummy; } hts_idx_t;
typedef struct hts_itr_t { int dummy; } hts_itr_t;

#define BAM_CMATCH 0
#define BAM_CINS 1
#define BAM_CDEL 2
#define BAM_CREF_SKIP 3
#define BAM_CSOFT_CLIP 4
#define BAM_CHARD_CLIP 5
#define BAM_CPAD 6
#define BAM_CEQUAL 7
#define BAM_CDIFF 8
#define BAM_CBACK 9
#define BAM_CIGAR_SHIFT 4
#define BAM_CIGAR_MASK 0xf
#define bam_cigar_op(c) ((c) & BAM_CIGAR_MASK)
#define bam_cigar_oplen(c) ((c) >> BAM_CIGAR_SHIFT)

#define BAM_FPAIRED 1
#define BAM_FPROPER_PAIR 2
#define BAM_FUNMAP 4
#define BAM_FMUNMAP 8
#define BAM_FREVERSE 16
#define BAM_FMREVERSE 32
#define BAM_FREAD1 64
#define BAM_FREAD2 128
#define BAM_FSECONDARY 256
#define BAM_FQCFAIL 512
#define BAM_FDUP 1024
#define BAM_FSUPPLEMENTARY 2048

#define bam_is_rev(b) (((b)->core.flag & BAM_FREVERSE) != 0)
#define bam_get_qname(b) ((char *) (b)->data)
#define bam_get_cigar(b) ((uint32_t *) ((b)->data + (b)->core.l_qname))
#define bam_get_seq(b) ((b)->data + ((b)->core.n_cigar << 2) + (b)->core.l_qname)
#define bam_get_qual(b) ((b)->data + ((b)->core.n_cigar << 2) + (b)->core.l_qname + (((b)->core.l_qseq + 1) >> 1))
#define bam_seqi(s, i) ((s)[(i) >> 1] >> ((~(i) & 1) << 2) & 0xf)

#define HFG_HTS_STUB(name) \
    do { fprintf(stderr, "[htslib stub] %s reached: BAM input is outside the oracle's scope\n", name); abort(); } while (0)

static inline samFile *sam_open(const char *fn, const char *mode) { HFG_HTS_STUB("sam_open"); return NULL; }
static inline int sam_close(samFile *fp) { HFG_HTS_STUB("sam_close"); return 0; }
static inline sam_hdr_t *sam_hdr_read(samFile *fp) { HFG_HTS_STUB("sam_hdr_read"); return NULL; }
static inline void sam_hdr_destroy(sam_hdr_t *h) { HFG_HTS_STUB("sam_hdr_destroy"); }
static inline int sam_hdr_name2tid(sam_hdr_t *h, const char *ref) { HFG_HTS_STUB("sam_hdr_name2tid"); return -1; }
static inline const char *sam_hdr_tid2name(const sam_hdr_t *h, int tid) { HFG_HTS_STUB("sam_hdr_tid2name"); return NULL; }
static inline int sam_read1(samFile *fp, sam_hdr_t *h, bam1_t *b) { HFG_HTS_STUB("sam_read1"); return -1; }
static inline hts_idx_t *sam_index_load(samFile *fp, const char *fn) { HFG_HTS_STUB("sam_index_load"); return NULL; }
static inline hts_itr_t *sam_itr_queryi(const hts_idx_t *idx, int tid, hts_pos_t beg, hts_pos_t end) { HFG_HTS_STUB("sam_itr_queryi"); return NULL; }
static inline hts_itr_t *sam_itr_querys(const hts_idx_t *idx, sam_hdr_t *hdr, const char *region) { HFG_HTS_STUB("sam_itr_querys"); return NULL; }
static inline int sam_itr_next(samFile *fp, hts_itr_t *itr, bam1_t *b) { HFG_HTS_STUB("sam_itr_next"); return -1; }
static inline void hts_idx_destroy(hts_idx_t *idx) { HFG_HTS_STUB("hts_idx_destroy"); }
static inline void hts_itr_destroy(hts_itr_t *itr) { HFG_HTS_STUB("hts_itr_destroy"); }
static inline bam1_t *bam_init1(void) { HFG_HTS_STUB("bam_init1"); return NULL; }
static inline void bam_destroy1(bam1_t *b) { HFG_HTS_STUB("bam_destroy1"); }
static inline bam1_t *bam_copy1(bam1_t *d, const bam1_t *s) { HFG_HTS_STUB("bam_copy1"); return NULL; }
static inline uint8_t *bam_aux_get(const bam1_t *b, const char tag[2]) { HFG_HTS_STUB("bam_aux_get"); return NULL; }
static inline int64_t bam_aux2i(const uint8_t *s) { HFG_HTS_STUB("bam_aux2i"); return 0; }
static inline char *bam_aux2Z(const uint8_t *s) { HFG_HTS_STUB("bam_aux2Z"); return NULL; }
static inline hts_pos_t bam_endpos(const bam1_t *b) { HFG_HTS_STUB("bam_endpos"); return 0; }

#endif
