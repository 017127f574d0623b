/*
 * hfg_layout_dev.cuh -- the observation keys, their window lists and tiles, built ON THE DEVICE.
 *
 * The host builder (hfg_layout.c: per-slice hash tables on pthreads, merge, scatter) costs ~4 ms for a 3 Gbp assembly --
 * as much as thirty E-steps.  Here the host only cuts the genome into segments (a byte scan over the region indices); the
 * raw window arrays go to the device as they are and a handful of kernels + CUB radix sorts / scans do the rest:
 *   pack     one thread per segment: packed observation word, beta bits of contig-end windows, slot of every window
 *   sort     stable LSD radix sort of the windows by (word, beta bits): equal keys become runs, windows ascending in a run
 *   keys     run heads -> provisional key ids, counts, first window; final numbering = radix sort of the keys by
 *            (region, count descending, first window) -- the same data-determined order as the host builder
 *   lists    exclusive scans give every key its stretch of klist and of the tile table; one pass scatters the key words
 *            (segment-transposed) and the window lists, one writes the tiles
 * The result is bit-identical to the host builder's (tests/test_gpu_parity.py::test_device_layout_equals_host_layout,
 * hfg_debug_layout_compare).  Integer work; temporary arrays live in the arena region that later holds the key table.
 */
#pragma once

#include <cub/cub.cuh>
#include <cuda_runtime.h>
#include <stdint.h>

#include "hfg_internal.h"

#define HFG_LHD static __device__ __forceinline__
#include "hfg_layout_inl.h"
#undef HFG_LHD

namespace hfgl {

struct PackArgs {
    /* raw windows (device copies of the arrays handed to hfg_set_chunks) */
    const uint16_t *cov, *mapq, *clip;
    const uint8_t *region;
    /* chunks */
    const hfg_chunk_desc *chunks;
    const int32_t *edge_head, *edge_tail;
    /* segments */
    const int32_t *seg_start, *seg_len, *seg_chunk;
    int32_t n_seg, capacity;
    /* configuration */
    int32_t adjust_contig_ends, mean_read_length;
    double min_read_fraction_at_ends, max_high_mapq_ratio, min_high_mapq_ratio, min_highly_clipped_ratio;
    /* out, indexed by global window */
    uint32_t *word;      /* full packed word */
    uint64_t *beta_bits; /* bits of beta for contig-end windows, 0 elsewhere */
    uint32_t *slot;      /* k * capacity + j: where the window's key word goes in wkeyT */
    uint32_t *idx;       /* identity, the payload of the first sort */
};

__global__ void pack_kernel(const PackArgs a) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= a.n_seg) return;
    const int c = a.seg_chunk[j];
    const hfg_chunk_desc ch = a.chunks[c];
    const int L = ch.n_windows, first = (int) (a.seg_start[j] - ch.offset), len = a.seg_len[j];
    const int head = a.edge_head[c], tail = a.edge_tail[c];
    for (int k = 0; k < len; k++) {
        const int w = first + k;
        const int64_t g = ch.offset + w;
        const int is_edge = w < head || w >= L - tail;
        const uint16_t cv = a.cov[g];
        const uint32_t mask = hfg_validity_mask(a.max_high_mapq_ratio, a.min_high_mapq_ratio, a.min_highly_clipped_ratio, cv,
                                                a.mapq[g], a.clip[g]);
        const uint32_t word = hfg_pack_word(mask, cv, w > 0 ? a.cov[g - 1] : (uint16_t) 0, a.region[g],
                                            w > 0 ? a.region[g - 1] : (uint8_t) 0, w, L, is_edge);
        uint64_t bb = 0;
        if (is_edge) {
            const double b = hfg_beta_of(a.adjust_contig_ends, a.min_read_fraction_at_ends, a.mean_read_length, ch.ctg_len, ch.s,
                                         ch.e, ch.window_len, w);
            bb = (uint64_t) __double_as_longlong(b);
        }
        a.word[g] = word;
        a.beta_bits[g] = bb;
        a.slot[g] = (uint32_t) k * (uint32_t) a.capacity + (uint32_t) j;
        a.idx[g] = (uint32_t) g;
    }
}

/* keys of the second (stable) sort: the packed word without the chunk-end bit, in the order left by the first sort */
__global__ void gather_words_kernel(const uint32_t *word, const uint32_t *idx1, uint32_t *kw, int64_t W) {
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < W) kw[i] = word[idx1[i]] & ~HFG_OBS_CHUNK_END;
}

/* 1 where a new key starts in the sorted order */
__global__ void head_flags_kernel(const uint32_t *kw_sorted, const uint32_t *idx2, const uint64_t *beta_bits, int32_t *head,
                                  int64_t W) {
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= W) return;
    int h = 1;
    if (i > 0) h = kw_sorted[i] != kw_sorted[i - 1] || beta_bits[idx2[i]] != beta_bits[idx2[i - 1]];
    head[i] = h;
}

/* per provisional key (run of the sorted order): start of the run, word, beta bits, first (smallest) window */
__global__ void run_heads_kernel(const int32_t *head, const int32_t *head_scan, const uint32_t *kw_sorted, const uint32_t *idx2,
                                 const uint64_t *beta_bits, int32_t *run_start, uint32_t *kword, uint64_t *kbb, int64_t W) {
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= W || !head[i]) return;
    const int32_t t = head_scan[i] - 1; /* inclusive scan of the flags */
    run_start[t] = (int32_t) i;
    kword[t] = kw_sorted[i];
    kbb[t] = beta_bits[idx2[i]];
}

__device__ __forceinline__ bool has_stats(uint32_t word) { return !(word & (HFG_OBS_CHUNK_START | HFG_OBS_SECOND)); }

/* sort key of the final numbering: region | 2^28-1-count | first window (hfg_layout.c: key_order_sort) */
__global__ void order_keys_kernel(const int32_t *run_start, const uint32_t *kword, const uint32_t *idx2, uint64_t *skey,
                                  int32_t *ident, int32_t P, int64_t W) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= P) return;
    const int64_t end = t + 1 < P ? run_start[t + 1] : W;
    const int32_t count = (int32_t) (end - run_start[t]);
    const uint32_t first = idx2[run_start[t]]; /* the stable sorts keep the windows of a run ascending */
    skey[t] = ((uint64_t) HFG_OBS_REGION(kword[t]) << 56) | ((uint64_t) (HFG_KEY_MAX - count) << 28) | (uint64_t) first;
    ident[t] = t;
}

/* final key p = rank of provisional key order[p]: tables of the key, list length, tile counts for every candidate tile
 * length (summed over the keys with integer atomics: exact) */
__global__ void key_tables_kernel(const int32_t *order, const int32_t *run_start, const uint32_t *kword, const uint64_t *kbb,
                                  int32_t *new_id, uint32_t *kdesc, double *kbeta, int32_t *cnt_stats, double beta0,
                                  unsigned long long *tiles_for_len, int32_t P, int64_t W) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    __shared__ unsigned long long s_tiles[3 * HFG_TILE + 1];
    for (int i = threadIdx.x; i <= 3 * HFG_TILE; i += blockDim.x) s_tiles[i] = 0ull;
    __syncthreads();
    int32_t my_cs = 0;
    if (p < P) {
        const int t = order[p];
        new_id[t] = p;
        const uint32_t word = kword[t];
        kdesc[p] = word;
        double b = beta0;
        if (word & HFG_OBS_EDGE) b = __longlong_as_double((long long) kbb[t]);
        kbeta[3 * (size_t) p] = b;
        kbeta[3 * (size_t) p + 1] = (word & HFG_OBS_EDGE) ? beta0 / b : 1.0;
        kbeta[3 * (size_t) p + 2] = (word & HFG_OBS_EDGE) ? sqrt(beta0 / b) : 1.0;
        const int64_t end = t + 1 < P ? run_start[t + 1] : W;
        const int32_t count = (int32_t) (end - run_start[t]);
        const int32_t cs = has_stats(word) ? count : 0;
        cnt_stats[p] = cs;
        my_cs = cs;
    }
    /* tiles this block's keys need for every candidate tile length: warp sums, one shared atomic per warp and length */
    for (int tl = HFG_TILE; tl <= 4 * HFG_TILE; tl++) {
        const unsigned v = __reduce_add_sync(0xffffffffu, (unsigned) ((my_cs + tl - 1) / tl));
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&s_tiles[tl - HFG_TILE], (unsigned long long) v);
    }
    __syncthreads();
    for (int i = threadIdx.x; i <= 3 * HFG_TILE; i += blockDim.x)
        if (s_tiles[i]) atomicAdd(&tiles_for_len[i], s_tiles[i]);
}

__global__ void tile_counts_kernel(const int32_t *cnt_stats, int32_t *ntile, int32_t tile_len, int32_t P) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < P) ntile[p] = (cnt_stats[p] + tile_len - 1) / tile_len;
}

/* tiles of every key, and the first tile of every region that has keys (the rest is filled by region_fill_kernel) */
__global__ void tiles_kernel(const int32_t *kbegin, const int32_t *cnt_stats, const int32_t *tbase, const uint32_t *kdesc,
                             int32_t *tile_key, int32_t *tile_begin, int32_t *tile_cnt, int32_t *region_tile_begin,
                             int32_t tile_len, int32_t P) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const int r = (int) HFG_OBS_REGION(kdesc[p]);
    if (p == 0 || (int) HFG_OBS_REGION(kdesc[p - 1]) != r) region_tile_begin[r] = tbase[p];
    const int32_t n = cnt_stats[p], b0 = kbegin[p];
    int32_t t = tbase[p];
    for (int32_t b = 0; b < n; b += tile_len, t++) {
        tile_key[t] = p;
        tile_begin[t] = b0 + b;
        tile_cnt[t] = n - b < tile_len ? n - b : tile_len;
    }
}

/* regions without keys start where the next region with keys starts (one thread: 65 entries) */
__global__ void region_fill_kernel(int32_t *region_tile_begin, int32_t n_tiles) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        region_tile_begin[HFG_MAX_REGIONS] = n_tiles;
        for (int r = HFG_MAX_REGIONS - 1; r >= 0; r--)
            if (region_tile_begin[r] < 0) region_tile_begin[r] = region_tile_begin[r + 1];
    }
}

/* key words (segment-transposed) and the window lists of the keys */
__global__ void scatter_kernel(const int32_t *head_scan, const int32_t *run_start, const int32_t *new_id, const uint32_t *idx2,
                               const uint32_t *word, const uint32_t *slot, const int32_t *kbegin, uint32_t *wkeyT,
                               int32_t *klist, int64_t W) {
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= W) return;
    const int32_t t = head_scan[i] - 1, p = new_id[t];
    const uint32_t g = idx2[i], wd = word[g];
    uint32_t kw = (uint32_t) p;
    if (wd & HFG_OBS_CHUNK_START) kw |= HFG_KEY_CHUNK_START;
    if (wd & HFG_OBS_CHUNK_END) kw |= HFG_KEY_CHUNK_END;
    wkeyT[slot[g]] = kw;
    if (has_stats(wd)) klist[kbegin[p] + (int32_t) (i - run_start[t])] = (int32_t) g;
}

static inline size_t align256(size_t x) { return (x + 255) & ~(size_t) 255; }

/* temporary device memory the build needs for W windows (carved by layout_build_device from one block) */
static size_t temp_bytes(int64_t W, size_t *cub_bytes_out) {
    size_t cub_bytes = 0, b = 0;
    cub::DeviceRadixSort::SortPairs(NULL, b, (const uint64_t *) NULL, (uint64_t *) NULL, (const uint32_t *) NULL, (uint32_t *) NULL,
                                    (int) W);
    cub_bytes = b;
    cub::DeviceRadixSort::SortPairs(NULL, b, (const uint32_t *) NULL, (uint32_t *) NULL, (const uint32_t *) NULL, (uint32_t *) NULL,
                                    (int) W);
    if (b > cub_bytes) cub_bytes = b;
    cub::DeviceRadixSort::SortPairs(NULL, b, (const uint64_t *) NULL, (uint64_t *) NULL, (const int32_t *) NULL, (int32_t *) NULL,
                                    (int) W);
    if (b > cub_bytes) cub_bytes = b;
    cub::DeviceScan::InclusiveSum(NULL, b, (const int32_t *) NULL, (int32_t *) NULL, (int) W);
    if (b > cub_bytes) cub_bytes = b;
    cub::DeviceScan::ExclusiveSum(NULL, b, (const int32_t *) NULL, (int32_t *) NULL, (int) W);
    if (b > cub_bytes) cub_bytes = b;
    cub_bytes = align256(cub_bytes + 256);
    if (cub_bytes_out) *cub_bytes_out = cub_bytes;
    const size_t w = (size_t) W;
    /* word, slot, idx, idx1, kw, kw_sorted, idx2, head, head_scan (4 B each); beta_bits, bb_sorted (8 B); per key (<= W):
     * run_start, kword, ident, order, new_id, cnt_stats, kbegin, ntile, tbase (4 B), kbb, skey, skey_sorted (8 B) */
    return 9 * align256(4 * w) + 2 * align256(8 * w) + 9 * align256(4 * w + 4) + 3 * align256(8 * w) + cub_bytes + 4096;
}

struct DeviceLayoutOut {
    uint32_t *wkeyT;  /* [smax * capacity], zero-filled by the caller */
    int32_t *klist;   /* [<= W] */
    int32_t *tile_key, *tile_begin, *tile_cnt; /* [<= W / HFG_TILE + W + 1] */
    int32_t *region_tile_begin;                /* [HFG_MAX_REGIONS + 1] */
    /* sized by the number of keys, known after phase 1 */
    uint32_t *kdesc;  /* [n_keys] */
    double *kbeta;    /* [n_keys][3] */
};

/* the temporary arrays of one build, carved from the caller's block */
struct DeviceLayoutTemp {
    uint32_t *word, *slot, *idx, *idx1, *kw, *kw_sorted, *idx2, *kword;
    int32_t *head, *head_scan, *run_start, *ident, *order, *new_id, *cnt_stats, *kbegin, *ntile, *tbase;
    uint64_t *beta_bits, *bb_sorted, *kbb, *skey, *skey_sorted;
    char *cub_temp;
    size_t cub_bytes;
    unsigned long long *tiles_for_len;
    int64_t W;
    int32_t P, capacity;
};

/* Phase 1 on `stream`: packed words, the two sorts, run heads.  temp: temp_bytes(W) of device memory.  Synchronises the
 * stream once and returns the number of distinct keys in t->P. */
static cudaError_t layout_build_device_keys(const PackArgs &pa_in, int64_t W, void *temp, cudaStream_t stream,
                                            DeviceLayoutTemp *t) {
    size_t cub_bytes = 0;
    temp_bytes(W, &cub_bytes);
    char *base = (char *) temp;
    size_t off = 0;
    const size_t w = (size_t) W;
#define TAKE(type, name, bytes) t->name = (type *) (base + off); off += align256(bytes)
    TAKE(uint32_t, word, 4 * w);
    TAKE(uint32_t, slot, 4 * w);
    TAKE(uint32_t, idx, 4 * w);
    TAKE(uint32_t, idx1, 4 * w);
    TAKE(uint32_t, kw, 4 * w);
    TAKE(uint32_t, kw_sorted, 4 * w);
    TAKE(uint32_t, idx2, 4 * w);
    TAKE(int32_t, head, 4 * w);
    TAKE(int32_t, head_scan, 4 * w);
    TAKE(uint64_t, beta_bits, 8 * w);
    TAKE(uint64_t, bb_sorted, 8 * w);
    TAKE(int32_t, run_start, 4 * w + 4);
    TAKE(uint32_t, kword, 4 * w + 4);
    TAKE(int32_t, ident, 4 * w + 4);
    TAKE(int32_t, order, 4 * w + 4);
    TAKE(int32_t, new_id, 4 * w + 4);
    TAKE(int32_t, cnt_stats, 4 * w + 4);
    TAKE(int32_t, kbegin, 4 * w + 4);
    TAKE(int32_t, ntile, 4 * w + 4);
    TAKE(int32_t, tbase, 4 * w + 4);
    TAKE(uint64_t, kbb, 8 * w);
    TAKE(uint64_t, skey, 8 * w);
    TAKE(uint64_t, skey_sorted, 8 * w);
    TAKE(char, cub_temp, cub_bytes);
    TAKE(unsigned long long, tiles_for_len, 8 * (3 * HFG_TILE + 1));
#undef TAKE
    t->cub_bytes = cub_bytes;
    t->W = W;
    t->capacity = pa_in.capacity;
    cudaError_t e;
    const int TB = 256;
    const unsigned gw = (unsigned) ((W + TB - 1) / TB);

    PackArgs pa = pa_in;
    pa.word = t->word;
    pa.beta_bits = t->beta_bits;
    pa.slot = t->slot;
    pa.idx = t->idx;
    pack_kernel<<<(pa.n_seg + TB - 1) / TB, TB, 0, stream>>>(pa);
    /* stable LSD sort by (word, beta bits): beta bits first, then the word */
    size_t cb = cub_bytes;
    if ((e = cub::DeviceRadixSort::SortPairs(t->cub_temp, cb, (const uint64_t *) t->beta_bits, t->bb_sorted,
                                             (const uint32_t *) t->idx, t->idx1, (int) W, 0, 64, stream)) != cudaSuccess)
        return e;
    gather_words_kernel<<<gw, TB, 0, stream>>>(t->word, t->idx1, t->kw, W);
    cb = cub_bytes;
    if ((e = cub::DeviceRadixSort::SortPairs(t->cub_temp, cb, (const uint32_t *) t->kw, t->kw_sorted, (const uint32_t *) t->idx1,
                                             t->idx2, (int) W, 0, 32, stream)) != cudaSuccess)
        return e;
    head_flags_kernel<<<gw, TB, 0, stream>>>(t->kw_sorted, t->idx2, t->beta_bits, t->head, W);
    cb = cub_bytes;
    if ((e = cub::DeviceScan::InclusiveSum(t->cub_temp, cb, (const int32_t *) t->head, t->head_scan, (int) W, stream)) != cudaSuccess)
        return e;
    run_heads_kernel<<<gw, TB, 0, stream>>>(t->head, t->head_scan, t->kw_sorted, t->idx2, t->beta_bits, t->run_start, t->kword,
                                            t->kbb, W);
    int32_t P = 0;
    if ((e = cudaMemcpyAsync(&P, t->head_scan + (W - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return e;
    t->P = P;
    return cudaGetLastError();
}

/* Phase 2: final numbering, key tables, lists, tiles, key words.  Synchronises the stream once (tile length). */
static cudaError_t layout_build_device_lists(const DeviceLayoutTemp *t, double beta0, const DeviceLayoutOut &out,
                                             cudaStream_t stream, int64_t *n_list, int32_t *n_tiles, int32_t *tile_len_out) {
    cudaError_t e;
    const int TB = 256;
    const int64_t W = t->W;
    const int32_t P = t->P;
    const unsigned gw = (unsigned) ((W + TB - 1) / TB), gp = (unsigned) ((P + TB - 1) / TB);
    const size_t cub_bytes = t->cub_bytes;
    size_t cb;
    order_keys_kernel<<<gp, TB, 0, stream>>>(t->run_start, t->kword, t->idx2, t->skey, t->ident, P, W);
    cb = cub_bytes;
    if ((e = cub::DeviceRadixSort::SortPairs(t->cub_temp, cb, (const uint64_t *) t->skey, t->skey_sorted, (const int32_t *) t->ident,
                                             t->order, P, 0, 62, stream)) != cudaSuccess)
        return e;
    if ((e = cudaMemsetAsync(t->tiles_for_len, 0, 8 * (3 * HFG_TILE + 1), stream)) != cudaSuccess) return e;
    key_tables_kernel<<<gp, TB, 0, stream>>>(t->order, t->run_start, t->kword, t->kbb, t->new_id, out.kdesc, out.kbeta,
                                             t->cnt_stats, beta0, t->tiles_for_len, P, W);
    cb = cub_bytes;
    if ((e = cub::DeviceScan::ExclusiveSum(t->cub_temp, cb, (const int32_t *) t->cnt_stats, t->kbegin, P, stream)) != cudaSuccess)
        return e;
    unsigned long long h_tiles[3 * HFG_TILE + 1];
    int32_t h_last[2] = {0, 0};
    if ((e = cudaMemcpyAsync(h_tiles, t->tiles_for_len, sizeof(h_tiles), cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return e;
    if ((e = cudaMemcpyAsync(&h_last[0], t->kbegin + (P - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, stream)) != cudaSuccess)
        return e;
    if ((e = cudaMemcpyAsync(&h_last[1], t->cnt_stats + (P - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, stream)) != cudaSuccess)
        return e;
    if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return e;
    /* tile length: HFG_TILE, or the shortest one that leaves every thread of the grid at most one tile (hfg_layout.c) */
    int tile_len = HFG_TILE;
    const long long capacity = t->capacity / (hfg_layout_tile_div > 0 ? hfg_layout_tile_div : 1); /* statistics workers */
    while (!((long long) h_tiles[tile_len - HFG_TILE] <= capacity - capacity / 16 || tile_len >= 4 * HFG_TILE)) tile_len++;
    const int32_t NT = (int32_t) h_tiles[tile_len - HFG_TILE];

    tile_counts_kernel<<<gp, TB, 0, stream>>>(t->cnt_stats, t->ntile, tile_len, P);
    cb = cub_bytes;
    if ((e = cub::DeviceScan::ExclusiveSum(t->cub_temp, cb, (const int32_t *) t->ntile, t->tbase, P, stream)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(out.region_tile_begin, 0xff, sizeof(int32_t) * (HFG_MAX_REGIONS + 1), stream)) != cudaSuccess) return e;
    tiles_kernel<<<gp, TB, 0, stream>>>(t->kbegin, t->cnt_stats, t->tbase, out.kdesc, out.tile_key, out.tile_begin, out.tile_cnt,
                                        out.region_tile_begin, tile_len, P);
    region_fill_kernel<<<1, 32, 0, stream>>>(out.region_tile_begin, NT);
    scatter_kernel<<<gw, TB, 0, stream>>>(t->head_scan, t->run_start, t->new_id, t->idx2, t->word, t->slot, t->kbegin, out.wkeyT,
                                          out.klist, W);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    *n_list = (int64_t) h_last[0] + h_last[1];
    *n_tiles = NT;
    *tile_len_out = tile_len;
    return cudaSuccess;
}

}  // namespace hfgl
