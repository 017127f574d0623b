/*
 * hfg_layout_inl.h -- the per-window quantities of the data layout, written once and compiled twice: plain C for the host
 * builder / checker (hfg_layout.c) and device code for the on-device layout build (hfg_layout_dev.cuh).  Integer
 * truncations and IEEE double arithmetic only, so both give the same bits.  The includer defines HFG_LHD.
 */
#ifndef HFG_LAYOUT_INL_H
#define HFG_LAYOUT_INL_H

/* submodules/common/common.c:142-148: min/max are int functions; double arguments are truncated at the call */
HFG_LHD int hfg_imin(int a, int b) { return a < b ? a : b; }
HFG_LHD int hfg_imax(int a, int b) { return a < b ? b : a; }

/* EM_computeAdjustmentBeta (hmm.c:301-316) for window i of a chunk */
HFG_LHD double hfg_beta_of(int adjust_contig_ends, double frac, int Lr, int ctg_len, int s, int e, int window_len, int i) {
    if (!adjust_contig_ends) return 1.0;
    const int mid = hfg_imin((int) (s + (double) window_len * (i + 0.5)), (int) ((s + (double) window_len * i + e) / 2));
    const int lo = hfg_imax(mid - Lr + 1, (int) (-(1 - frac) * Lr));
    const int hi = hfg_imin(mid, (int) (ctg_len - frac * Lr));
    const double b = (double) (hi - lo) / Lr;
    return b <= 0.25 ? 0.25 : b;
}

/* validity of the Dup / Col / END columns for a window (hmm_utils.c:2229-2264): bit0 Dup invalid, bit1 Col invalid,
 * bit2 END column valid */
HFG_LHD uint32_t hfg_validity_mask(double max_high_mapq_ratio, double min_high_mapq_ratio, double min_highly_clipped_ratio,
                                   uint16_t cov, uint16_t mapq, uint16_t clip) {
    const double rm = (double) mapq / (0.1 + cov);
    const double rc = (double) clip / (0.1 + cov);
    uint32_t m = 0;
    if (rm > max_high_mapq_ratio) m |= 1u;            /* Dup invalid */
    if (rm < min_high_mapq_ratio) m |= 2u;            /* Col invalid */
    if (!(rc < min_highly_clipped_ratio)) m |= 4u;    /* END column valid */
    return m;
}

/* the packed observation word of window w (0-based) of a chunk of L windows; prev_cov / prev_region are those of window
 * w-1 (ignored for w == 0) */
HFG_LHD uint32_t hfg_pack_word(uint32_t mask, uint16_t cov, uint16_t prev_cov, uint8_t region, uint8_t prev_region, int w,
                               int L, int is_edge) {
    uint32_t word = HFG_OBS_VALID;
    word |= (uint32_t) (uint8_t) cov;
    if (w > 0) word |= (uint32_t) (uint8_t) prev_cov << 8;
    word |= (uint32_t) region << 16;
    word |= mask << 22;
    if (w > 0 && region != prev_region) word |= HFG_OBS_REGION_CHANGE;
    if (w == 0) word |= HFG_OBS_CHUNK_START;
    if (w == 1) word |= HFG_OBS_SECOND;
    if (w == L - 1) word |= HFG_OBS_CHUNK_END;
    if (is_edge) word |= HFG_OBS_EDGE;
    return word;
}

#endif /* HFG_LAYOUT_INL_H */
