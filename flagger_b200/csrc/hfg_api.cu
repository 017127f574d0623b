/*
 * hfg_api.cu -- the C-ABI of libhfg (include/hfg.h): context, device memory, launches.
 *
 * The compute entry points stand where the reference has EM_runOneIterationForList / EM_runForwardForList
 * (submodules/hmm/hmm.c:739-816).  There is NO CPU fallback: without a usable CUDA device every entry point
 * returns HFG_ERR_CUDA with a message.
 */
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <pthread.h>
#include <string.h>
#include <time.h>

#include "hfg_estep.cuh"
#include "hfg_estep_quad.cuh"
#include "hfg_estep_v3.cuh"
#include "hfg_internal.h"
#include "hfg_layout_dev.cuh"

#define STAGE_SLOTS 4
#define HFG_EM_LOGLIK_SLOTS 4096 /* E-steps a device-resident EM loop can record (carved from the arena) */
#define STATS_DOUBLES ((int) (sizeof(hfg_region_stats) / sizeof(double)))

struct hfg_ctx {
    hfg_config cfg;
    hfg_layout lay;
    int have_chunks;
    int device, num_sms, max_blocks, grid;
    int threads;          /* CTA size: 1024 (quad kernel); first-generation kernel: 512, or 256 when the model needs the shared memory */
    int segs_per_cta;     /* segments a CTA owns: threads / 4 (quad kernel: four lanes per segment) or threads */
    int quad;             /* kernel generation: 3 = hfg_estep_v3_kernel (default), 2 = hfg_estep_quad_kernel (HFG_KERNEL=quad),
                             0 = hfg_estep_kernel (HFG_KERNEL=v1) */
    size_t smem_optin;    /* shared memory a CTA may ask for on this device */
    int32_t n_hot, lab_bytes;
    const void *kernel;   /* the matching kernel instantiation */
    size_t smem_bytes;
    cudaStream_t stream;
    cudaEvent_t ev0, ev1; /* around the E-step kernel */
    cudaEvent_t ev2, ev3; /* around the whole device side of the last blocking call */
    int ev_valid, span_valid;
    int dbg;              /* HFG_DBG at hfg_create: instrumentation switches of the kernel */
    int timing;           /* hfg_debug_set_timing: the blocking calls record ev0..ev3 (and give up their fast path) */
    unsigned long long done_seq; /* blocking fast path: sequence number of the completion word behind h_out */
    /* device */
    uint32_t *d_wkeyT, *d_kdesc, *d_wposT, *d_wkeyH, *d_khot;
    int32_t *d_hot_key, *d_hot_range;
    double *d_scan_stash;
    int32_t *d_seg_start, *d_seg_len, *d_block_reset, *d_err, *d_ticket;
    int32_t *d_klist, *d_tile_key, *d_tile_begin, *d_tile_cnt, *d_region_tile_begin;
    double *d_kbeta, *d_tabM, *d_tabMT, *d_scrFT, *d_scrXB, *d_block_tot, *d_partials, *d_out, *d_seg_loglik, *d_post;
    hfg_region_params *d_params[STAGE_SLOTS];
    int8_t *d_labels;
    long long *d_phase_clock;
    void *d_arena;   /* one device allocation behind all per-run buffers */
    size_t arena_bytes;
    size_t h_out_bytes, h_labels_bytes;
    void *d_keyblock; /* the buffers sized by the number of observation keys (key table, descriptors, betas) */
    size_t keyblock_bytes;
    int32_t capacity_arg; /* the slot count the layout was asked for (before trimming): hfg_debug_layout_compare */
    int8_t *h_labels; /* pinned staging for the label read-back */
    /* pinned host staging */
    hfg_region_params *h_params[STAGE_SLOTS];
    cudaEvent_t stage_ev[STAGE_SLOTS];
    int stage_next;
    double *h_out;
    /* last call (for hfg_get_posteriors) */
    int have_last;
    double last_alpha[16];
    hfg_region_params *last_params;
    /* the blocking entry points replay a captured graph (parameter upload -> kernel -> result read-back): one launch
     * call per E-step instead of six */
    cudaGraphExec_t gexec;
    int graph_disabled;
    EstepArgs graph_args;
    /* multi-GPU peer exchange (hfg_peer_export / hfg_peer_connect) */
    int n_ranks, rank;
    double *d_mailbox;                 /* own mailbox (its own allocation: it is shared through CUDA IPC) */
    double *peer_box[HFG_MAX_PEERS];   /* every rank's mailbox as mapped into this process */
    unsigned long long *d_epoch;
    /* device-resident EM loop (hfg_em_begin / hfg_em_enqueue / hfg_em_finish) */
    hfg_region_params *d_em_params;
    int32_t *d_em_state;
    double *d_em_logliks;
    int em_active, em_max, em_limit, em_enqueued, em_ev_cap;
    double em_alpha[16], em_tol;
    cudaEvent_t *em_ev; /* [2 * em_ev_cap] */
    void *d_flush;
    size_t flush_bytes;
    int64_t launches;
    /* negative-binomial model: emission table from the host, per-tile pair masses back (run_blocking_nb) */
    int nb;
    double *d_nb_table, *h_nb_table;     /* [R][4][HFG_NB_XSTRIDE] device / pinned */
    double *d_nb_tile_col, *h_nb_tile_col; /* [n_tiles][4] */
    double *h_nb_hist;                   /* pinned [R][4][256]: the grid-folded histogram of a blocking call */
    int32_t *h_tile_key;                 /* [n_tiles] host copies for folding the tile masses into the histogram */
    uint32_t *h_kdesc;                   /* [n_keys] */
    int32_t *d_nb_bins;                  /* device-resident loop: [R * 250 + 1] bin offsets, then [n_tiles] the tiles of every bin */
    double *d_nb_lgx1;                   /* [251] lgamma(x + 1) */
    char err[512];
};

static char g_create_err[512] = "";

static int fail(hfg_ctx *ctx, int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(ctx ? ctx->err : g_create_err, 512, fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                                              \
    do {                                                                                                      \
        cudaError_t e_ = (call);                                                                              \
        if (e_ != cudaSuccess)                                                                                \
            return fail(ctx, HFG_ERR_CUDA, "%s failed: %s (no CPU fallback exists)", #call, cudaGetErrorString(e_)); \
    } while (0)

/* worst-case number of component evaluations of an ordinary window: every Gaussian state with four distinct alphas */
static int max_tasks(const hfg_config *cfg) {
    int n = 0;
    for (int s = 0; s < HFG_NS; s++)
        if (!(cfg->model_type == HFG_MODEL_TRUNC_EXP_GAUSSIAN && s == HFG_STATE_ERR)) n += 4 * cfg->n_comps[s];
    return n;
}

static size_t smem_bytes_for(int R, int G, int NT, int threads) {
    size_t doubles = (size_t) R * RT_STRIDE2(G, NT) + 3 * (threads / 32) * 16 + 8 + (size_t) hfg_acc_rows(G) * (threads + 1);
    return doubles * sizeof(double);
}

/* shared memory of the quad kernel (hfg_estep_quad.cuh): region tables, second-level scan buffers, the CTA messages, and
 * the larger of the per-warp statistics rows, the grid totals and -- aliasing everything behind the region tables -- the
 * M-step work area of the device-resident loop */
/* the M-step work area of the kernel tail (hfg_estep_tail): per region of a batch its parameters, statistics and 282 doubles,
 * plus 128 doubles of scratch per warp and one exchange slot per thread for the rate fits, plus 576 doubles (HFG_TAIL_TOP) at the top of the
 * area for the scout warps and the convergence flag */
static size_t mstep_work_bytes(int threads, int regions) {
    return ((size_t) regions * ((sizeof(hfg_region_params) + sizeof(hfg_region_stats)) / sizeof(double) + 282) + (size_t) threads +
            (size_t) (threads / 32) * 128 + 576) * sizeof(double);
}

static size_t smem_bytes_quad(int R, int G, int threads, int smax) {
    const size_t warps = (size_t) threads / 32, nstat = (size_t) hfg_nstat(G);
    const size_t rt = ((size_t) R * QRT_STRIDE(G) + 1) & ~(size_t) 1;
    size_t stat = ((warps * nstat > (size_t) R * nstat ? warps * nstat : (size_t) R * nstat) + 1) & ~(size_t) 1;
    const size_t labels = smax <= HFGQ_LAB_SMAX ? ((size_t) (threads / 4) * smax + 16 + 7) / 8 : 0; /* staged labels (bytes -> doubles) */
    size_t work = 4 * warps * 16 + 8 + (size_t) (threads / 4) * 8 + stat + labels;
    const size_t mstep = mstep_work_bytes(threads, R < 4 ? R : 4) / sizeof(double);
    if (mstep > work) work = mstep;
    return (rt + work) * sizeof(double);
}

/* shared memory of the third-generation kernel before the staged labels and the hot-key table: region tables, second-level
 * scan buffers, CTA messages, per-warp statistics rows / grid totals */
static size_t smem_base_v3(int R, int G, int threads) {
    const size_t warps = (size_t) threads / 32, nstat = (size_t) hfg_nstat(G);
    const size_t rt = ((size_t) R * QRT_STRIDE(G) + 1) & ~(size_t) 1;
    const size_t stat = ((warps * nstat > (size_t) R * nstat ? warps * nstat : (size_t) R * nstat) + 1) & ~(size_t) 1;
    return (rt + 3 * warps * 16 + 8 + stat) * sizeof(double);
}

static int total_gauss_comps(const hfg_config *cfg) {
    int G = 0;
    for (int s = 0; s < HFG_NS; s++)
        if (!(cfg->model_type == HFG_MODEL_TRUNC_EXP_GAUSSIAN && s == HFG_STATE_ERR)) G += cfg->n_comps[s];
    return G;
}

/* quad kernel set-up: slot (k * capacity + j) of every window, then the list position of every listed window at its slot */
__global__ void hfg_wslot_kernel(const int32_t *seg_start, const int32_t *seg_len, int32_t n_seg, int32_t capacity, uint32_t *wslot) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_seg) return;
    const int first = seg_start[j], len = seg_len[j];
    for (int k = 0; k < len; k++) wslot[first + k] = (uint32_t) k * (uint32_t) capacity + (uint32_t) j;
}
/* third-generation kernel set-up.  One CTA: region r gets slots in proportion to its statistics tiles (~ its windows), never
 * more than it has keys, leftovers go to the regions that can still use them; inside a region the keys are numbered hottest
 * first, so its first h_r keys are taken.  hot_key[slot] = key id, khot[key] = slot (preset to 0xffffffff). */
__global__ void hfg_hot_assign_kernel(const uint32_t *kdesc, int32_t n_keys, const int32_t *region_tile_begin, int32_t R,
                                      int32_t n_hot, int32_t *hot_key, uint32_t *khot, int32_t *hot_range) {
    __shared__ int32_t key_begin[HFG_MAX_REGIONS + 1], h[HFG_MAX_REGIONS], base[HFG_MAX_REGIONS + 1];
    const int tid = threadIdx.x;
    if (tid <= R) {
        /* first key whose region is >= tid (the keys are sorted by region) */
        int lo = 0, hi = n_keys;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if ((int) HFG_OBS_REGION(kdesc[mid]) < tid) lo = mid + 1; else hi = mid;
        }
        key_begin[tid] = tid == R ? n_keys : lo;
    }
    __syncthreads();
    if (tid == 0) {
        const long long tiles_all = region_tile_begin[HFG_MAX_REGIONS] > 0 ? region_tile_begin[HFG_MAX_REGIONS] : 1;
        int left = n_hot;
        for (int r = 0; r < R; r++) {
            const int keys = key_begin[r + 1] - key_begin[r];
            const long long tiles = region_tile_begin[r + 1] - region_tile_begin[r];
            int want = (int) ((long long) n_hot * tiles / tiles_all);
            h[r] = want < keys ? want : keys;
            left -= h[r];
        }
        for (int r = 0; r < R && left > 0; r++) {
            const int keys = key_begin[r + 1] - key_begin[r];
            const int add = keys - h[r] < left ? keys - h[r] : left;
            h[r] += add;
            left -= add;
        }
        base[0] = 0;
        for (int r = 0; r < R; r++) base[r + 1] = base[r] + h[r];
        for (int r = 0; r < R; r++) {
            hot_range[3 * r] = key_begin[r];
            hot_range[3 * r + 1] = h[r];
            hot_range[3 * r + 2] = base[r];
        }
    }
    __syncthreads();
    for (int r = 0; r < R; r++)
        for (int i = tid; i < h[r]; i += blockDim.x) {
            hot_key[base[r] + i] = key_begin[r] + i;
            khot[key_begin[r] + i] = (uint32_t) (base[r] + i);
        }
    /* slots no region could use (fewer keys than slots) keep pointing at key 0: the fill copies a valid row */
    for (int i = base[R] + tid; i < n_hot; i += blockDim.x) hot_key[i] = 0;
}
__global__ void hfg_hot_rewrite_kernel(const uint32_t *wkeyT, const uint32_t *khot, size_t slots, uint32_t *wkeyH) {
    const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= slots) return;
    const uint32_t w = wkeyT[i], slot = khot[HFG_KEY_ID(w)];
    wkeyH[i] = slot == 0xffffffffu ? w : ((w & ~0x0fffffffu) | HFG3_HOT | slot);
}
__global__ void hfg_wpos_kernel(const int32_t *klist, int64_t n_list, const uint32_t *wslot, uint32_t *wposT) {
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_list) wposT[wslot[klist[i]]] = (uint32_t) i;
}

extern "C" const char *hfg_last_error(const hfg_ctx *ctx) { return ctx ? ctx->err : g_create_err; }

extern "C" int hfg_create(hfg_ctx **out, const hfg_config *cfg) {
    hfg_ctx *ctx = NULL;
    if (!out || !cfg) return fail(NULL, HFG_ERR_INVALID, "hfg_create: NULL argument");
    *out = NULL;
    const int nb = cfg->model_type == HFG_MODEL_NEGATIVE_BINOMIAL;
    if (!nb && cfg->model_type != HFG_MODEL_TRUNC_EXP_GAUSSIAN && cfg->model_type != HFG_MODEL_GAUSSIAN)
        return fail(NULL, HFG_ERR_INVALID, "hfg_create: model_type %d not supported", cfg->model_type);
    if (cfg->n_regions < 1 || cfg->n_regions > HFG_MAX_REGIONS)
        return fail(NULL, HFG_ERR_INVALID, "hfg_create: n_regions %d outside 1..%d", cfg->n_regions, HFG_MAX_REGIONS);
    for (int s = 0; s < HFG_NS; s++)
        if (cfg->n_comps[s] < 1 || cfg->n_comps[s] > HFG_MAX_COMPS)
            return fail(NULL, HFG_ERR_INVALID, "hfg_create: n_comps[%d]=%d outside 1..%d", s, cfg->n_comps[s], HFG_MAX_COMPS);
    {
        cudaError_t e = cudaSetDevice(cfg->device);
        if (e != cudaSuccess)
            return fail(NULL, HFG_ERR_CUDA, "cudaSetDevice(%d) failed: %s (libhfg has no CPU fallback)", cfg->device,
                        cudaGetErrorString(e));
    }
    ctx = (hfg_ctx *) calloc(1, sizeof(hfg_ctx));
    if (!ctx) return fail(NULL, HFG_ERR_NOMEM, "hfg_create: out of memory");
    ctx->cfg = *cfg;
    ctx->device = cfg->device;
    /* individual attributes: cudaGetDeviceProperties costs about a millisecond */
    struct { int multiProcessorCount; size_t sharedMemPerBlockOptin; } prop = {0, 0};
    int optin = 0;
    cudaError_t e = cudaDeviceGetAttribute(&prop.multiProcessorCount, cudaDevAttrMultiProcessorCount, cfg->device);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, cfg->device);
    if (e != cudaSuccess) {
        fail(NULL, HFG_ERR_CUDA, "cudaDeviceGetAttribute failed: %s", cudaGetErrorString(e));
        free(ctx);
        return HFG_ERR_CUDA;
    }
    prop.sharedMemPerBlockOptin = (size_t) optin;
    ctx->num_sms = prop.multiProcessorCount;
    const int G = total_gauss_comps(cfg);
    ctx->nb = nb;
    /* HFG_KERNEL=v1: A/B switch, the first-generation kernel (one thread per segment, hfg_estep.cuh) */
    ctx->quad = 3;
    if (getenv("HFG_KERNEL") && strcmp(getenv("HFG_KERNEL"), "v1") == 0) ctx->quad = 0;
    if (getenv("HFG_KERNEL") && strcmp(getenv("HFG_KERNEL"), "quad") == 0) ctx->quad = 2;
    ctx->smem_optin = (size_t) optin;
    if (ctx->quad == 3) {
        ctx->threads = HFG3_THREADS;
        ctx->kernel = nb ? (const void *) hfg_estep_v3_kernel<HFG3_THREADS, true> : (const void *) hfg_estep_v3_kernel<HFG3_THREADS>;
        /* HFG_THREADS: A/B switch of the CTA size (fewer, longer segments: cheaper scans, fewer warps to hide latency) */
        const int want = getenv("HFG_THREADS") ? atoi(getenv("HFG_THREADS")) : 0;
        if (!nb && (want == 256 || want == 320 || want == 384 || want == 512 || want == 640)) {
            ctx->threads = want;
            ctx->kernel = want == 256 ? (const void *) hfg_estep_v3_kernel<256> : want == 320 ? (const void *) hfg_estep_v3_kernel<320>
                          : want == 384 ? (const void *) hfg_estep_v3_kernel<384> : want == 512 ? (const void *) hfg_estep_v3_kernel<512>
                                                                                                : (const void *) hfg_estep_v3_kernel<640>;
        }
        ctx->segs_per_cta = ctx->threads;
        ctx->smem_bytes = (size_t) optin - 1024; /* (less static shared memory) the hot-key table takes whatever the rest leaves:
                                                    sized per run in hfg_set_chunks */
        if (smem_base_v3(cfg->n_regions, G, ctx->threads) + 16 * 1024 > (size_t) optin) {
            fail(NULL, HFG_ERR_INVALID, "model too large for shared memory (%d regions x %d components)", cfg->n_regions, G);
            free(ctx);
            return HFG_ERR_INVALID;
        }
    } else if (ctx->quad == 2) {
        ctx->threads = HFGQ_THREADS;
        ctx->segs_per_cta = HFGQ_THREADS / 4;
        ctx->kernel = nb ? (const void *) hfg_estep_quad_kernel<HFGQ_THREADS, true> : (const void *) hfg_estep_quad_kernel<HFGQ_THREADS>;
        ctx->smem_bytes = smem_bytes_quad(cfg->n_regions, G, ctx->threads, HFGQ_LAB_SMAX); /* upper bound; per run in hfg_set_chunks */
    } else {
        ctx->threads = HFG_THREADS_MAX;
        ctx->kernel = nb ? (const void *) hfg_estep_kernel<HFG_THREADS_MAX, true> : (const void *) hfg_estep_kernel<HFG_THREADS_MAX>;
        ctx->smem_bytes = smem_bytes_for(cfg->n_regions, G, max_tasks(cfg), ctx->threads);
        /* HFG_THREADS=256: the 256-thread instantiation for every model (half the statistics area in shared memory) */
        const int force_small = getenv("HFG_THREADS") && atoi(getenv("HFG_THREADS")) == HFG_THREADS_MIN;
        if (force_small || ctx->smem_bytes > (size_t) optin) { /* many components / regions: halve the CTA, halving the statistics area */
            ctx->threads = HFG_THREADS_MIN;
            ctx->kernel = nb ? (const void *) hfg_estep_kernel<HFG_THREADS_MIN, true> : (const void *) hfg_estep_kernel<HFG_THREADS_MIN>;
            ctx->smem_bytes = smem_bytes_for(cfg->n_regions, G, max_tasks(cfg), ctx->threads);
        }
        ctx->segs_per_cta = ctx->threads;
    }
    if (!ctx->quad && max_tasks(cfg) > HFG_MAX_TASKS) {
        fail(NULL, HFG_ERR_INVALID, "too many mixture components (%d component evaluations per window, limit %d)", max_tasks(cfg), HFG_MAX_TASKS);
        free(ctx);
        return HFG_ERR_INVALID;
    }
    if (ctx->smem_bytes > (size_t) prop.sharedMemPerBlockOptin) {
        fail(NULL, HFG_ERR_INVALID, "model too large for shared memory: %zu bytes needed for %d regions x %d components, %zu available",
             ctx->smem_bytes, cfg->n_regions, G, (size_t) prop.sharedMemPerBlockOptin);
        free(ctx);
        return HFG_ERR_INVALID;
    }
    int coop = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, cfg->device);
    e = cudaFuncSetAttribute(ctx->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) ctx->smem_bytes);
    int per_sm = 0;
    if (e == cudaSuccess)
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ctx->kernel, ctx->threads, ctx->smem_bytes);
    if (e != cudaSuccess || !coop || per_sm < 1) {
        fail(NULL, HFG_ERR_CUDA, "E-step kernel cannot be launched cooperatively on device %d (%s; coop=%d, blocks/SM=%d)",
             cfg->device, cudaGetErrorString(e), coop, per_sm);
        free(ctx);
        return HFG_ERR_CUDA;
    }
    ctx->max_blocks = ctx->num_sms; /* one persistent 512-thread CTA per SM, all co-resident (cooperative launch) */
    ctx->n_ranks = 1;
    /* from here on a failure releases whatever exists already through hfg_destroy (it skips what is still NULL) */
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess ||
        cudaEventCreate(&ctx->ev2) != cudaSuccess || cudaEventCreate(&ctx->ev3) != cudaSuccess) {
        fail(NULL, HFG_ERR_CUDA, "stream/event creation failed: %s", cudaGetErrorString(cudaGetLastError()));
        hfg_destroy(ctx);
        return HFG_ERR_CUDA;
    }
    const size_t pbytes = sizeof(hfg_region_params) * (size_t) cfg->n_regions;
    {
        /* one pinned and one device block for the ring of parameter staging slots */
        char *hp = NULL, *dp = NULL;
        if (cudaMallocHost((void **) &hp, pbytes * STAGE_SLOTS) != cudaSuccess ||
            cudaMalloc((void **) &dp, pbytes * STAGE_SLOTS) != cudaSuccess) {
            fail(NULL, HFG_ERR_CUDA, "parameter staging allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
            if (hp) cudaFreeHost(hp);
            hfg_destroy(ctx);
            return HFG_ERR_CUDA;
        }
        for (int i = 0; i < STAGE_SLOTS; i++) {
            ctx->h_params[i] = (hfg_region_params *) (hp + pbytes * i);
            ctx->d_params[i] = (hfg_region_params *) (dp + pbytes * i);
        }
        for (int i = 0; i < STAGE_SLOTS; i++)
            if (cudaEventCreateWithFlags(&ctx->stage_ev[i], cudaEventDisableTiming) != cudaSuccess) {
                fail(NULL, HFG_ERR_CUDA, "event creation failed: %s", cudaGetErrorString(cudaGetLastError()));
                hfg_destroy(ctx);
                return HFG_ERR_CUDA;
            }
    }
    ctx->last_params = (hfg_region_params *) malloc(pbytes);
    if (!ctx->last_params) {
        fail(NULL, HFG_ERR_NOMEM, "hfg_create: out of memory");
        hfg_destroy(ctx);
        return HFG_ERR_NOMEM;
    }
    ctx->dbg = getenv("HFG_DBG") ? atoi(getenv("HFG_DBG")) : 0;
    ctx->timing = getenv("HFG_NO_FAST_BLOCKING") != NULL; /* A/B switch: the blocking calls always take the graph path */
    ctx->graph_disabled = getenv("HFG_NO_GRAPH") != NULL; /* A/B switch: plain stream launches instead of graph replay */
    if (nb) {
        const size_t tb = sizeof(double) * (size_t) cfg->n_regions * 4 * HFG_NB_XSTRIDE;
        /* (pinned: the table, then the histogram a blocking call reads back -- the same [R][4][256] shape) */
        if (cudaMalloc((void **) &ctx->d_nb_table, tb) != cudaSuccess || cudaMallocHost((void **) &ctx->h_nb_table, 2 * tb) != cudaSuccess) {
            fail(NULL, HFG_ERR_CUDA, "negative-binomial table allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
            hfg_destroy(ctx);
            return HFG_ERR_CUDA;
        }
        memset(ctx->h_nb_table, 0, 2 * tb);
        ctx->h_nb_hist = ctx->h_nb_table + (size_t) cfg->n_regions * 4 * HFG_NB_XSTRIDE;
    }
    *out = ctx;
    return HFG_OK;
}

/* One released arena per process is kept for the next context on the same device: cudaMalloc of the ~130 MB arena of a
 * 3 Gbp job costs 0.4 ms on a quiet driver but was measured at 6-60 ms on a freshly booted box, more than the job's upload.
 * hfg_release_cached_memory() returns it to the driver; HFG_NO_ARENA_CACHE=1 disables the cache. */
static pthread_mutex_t g_cache_mu = PTHREAD_MUTEX_INITIALIZER;
#define HFG_CACHE_SLOTS 4 /* kinds: 0 per-run arena, 1 key block (device); 2 result block, 3 label staging (pinned host) */
static struct { void *ptr; size_t bytes; int device; } g_cache[HFG_CACHE_SLOTS] = {{NULL, 0, -1}, {NULL, 0, -1}, {NULL, 0, -1},
                                                                                     {NULL, 0, -1}};
static void cache_free(void *p, int kind) {
    if (kind >= 2) cudaFreeHost(p); else cudaFree(p);
}

/* kind 0: the per-run arena; kind 1: the key block.  One cached block of each kind; a newly released block replaces the
 * cached one of its kind (the next job most likely looks like the last one). */
static void arena_release(void *ptr, size_t bytes, int device, int kind) {
    if (!ptr) return;
    void *drop = ptr;
    int drop_dev = device;
    if (!getenv("HFG_NO_ARENA_CACHE")) {
        pthread_mutex_lock(&g_cache_mu);
        drop = g_cache[kind].ptr;
        drop_dev = g_cache[kind].device;
        g_cache[kind].ptr = ptr;
        g_cache[kind].bytes = bytes;
        g_cache[kind].device = device;
        pthread_mutex_unlock(&g_cache_mu);
    }
    if (drop) {
        cudaSetDevice(drop_dev);
        cache_free(drop, kind);
    }
    cudaSetDevice(device);
}

static void *arena_acquire(size_t bytes, int device, size_t *got, int kind) {
    void *ptr = NULL;
    pthread_mutex_lock(&g_cache_mu);
    if (g_cache[kind].ptr && g_cache[kind].device == device && g_cache[kind].bytes >= bytes &&
        g_cache[kind].bytes <= 2 * bytes + (1u << 20)) {
        ptr = g_cache[kind].ptr;
        *got = g_cache[kind].bytes;
        g_cache[kind].ptr = NULL;
        g_cache[kind].bytes = 0;
        g_cache[kind].device = -1;
    }
    pthread_mutex_unlock(&g_cache_mu);
    if (!ptr) {
        if ((kind >= 2 ? cudaMallocHost(&ptr, bytes) : cudaMalloc(&ptr, bytes)) != cudaSuccess) return NULL;
        *got = bytes;
    }
    return ptr;
}

extern "C" int hfg_device_warmup(int device) {
    if (cudaSetDevice(device) != cudaSuccess || cudaFree(0) != cudaSuccess)
        return fail(NULL, HFG_ERR_CUDA, "cannot use CUDA device %d: %s (libhfg has no CPU fallback)", device,
                    cudaGetErrorString(cudaGetLastError()));
    return HFG_OK;
}

extern "C" void hfg_release_cached_memory(void) {
    pthread_mutex_lock(&g_cache_mu);
    for (int i = 0; i < HFG_CACHE_SLOTS; i++)
        if (g_cache[i].ptr) {
            cudaSetDevice(g_cache[i].device);
            cache_free(g_cache[i].ptr, i);
            g_cache[i].ptr = NULL;
            g_cache[i].bytes = 0;
            g_cache[i].device = -1;
        }
    pthread_mutex_unlock(&g_cache_mu);
}

static void free_device(hfg_ctx *ctx) {
    if (ctx->gexec) { /* the captured graph holds pointers into the buffers released below */
        cudaGraphExecDestroy(ctx->gexec);
        ctx->gexec = NULL;
    }
    arena_release(ctx->d_arena, ctx->arena_bytes, ctx->device, 0);
    ctx->arena_bytes = 0;
    arena_release(ctx->d_keyblock, ctx->keyblock_bytes, ctx->device, 1);
    ctx->d_keyblock = NULL;
    ctx->keyblock_bytes = 0;
    cudaFree(ctx->d_post);
    arena_release(ctx->h_out, ctx->h_out_bytes, ctx->device, 2);
    arena_release(ctx->h_labels, ctx->h_labels_bytes, ctx->device, 3);
    ctx->d_arena = NULL;
    ctx->d_wkeyT = ctx->d_kdesc = ctx->d_wposT = ctx->d_wkeyH = ctx->d_khot = NULL;
    ctx->d_hot_key = NULL;
    ctx->d_scan_stash = NULL;
    ctx->d_tabMT = NULL;
    ctx->d_seg_start = ctx->d_seg_len = ctx->d_block_reset = ctx->d_err = ctx->d_ticket = NULL;
    ctx->d_klist = ctx->d_tile_key = ctx->d_tile_begin = ctx->d_tile_cnt = ctx->d_region_tile_begin = NULL;
    ctx->d_kbeta = ctx->d_tabM = ctx->d_scrFT = ctx->d_scrXB = ctx->d_block_tot = ctx->d_partials = NULL;
    ctx->d_out = ctx->d_seg_loglik = ctx->d_post = NULL;
    ctx->d_labels = NULL;
    ctx->d_phase_clock = NULL;
    ctx->d_em_params = NULL; ctx->d_em_state = NULL; ctx->d_em_logliks = NULL;
    ctx->em_active = 0;
    ctx->h_out = NULL;
    ctx->h_labels = NULL;
    if (ctx->d_nb_tile_col) cudaFree(ctx->d_nb_tile_col);
    if (ctx->h_nb_tile_col) cudaFreeHost(ctx->h_nb_tile_col);
    free(ctx->h_tile_key);
    free(ctx->h_kdesc);
    ctx->d_nb_tile_col = ctx->h_nb_tile_col = NULL;
    if (ctx->d_nb_bins) cudaFree(ctx->d_nb_bins);
    if (ctx->d_nb_lgx1) cudaFree(ctx->d_nb_lgx1);
    ctx->d_nb_bins = NULL;
    ctx->d_nb_lgx1 = NULL;
    ctx->h_tile_key = NULL;
    ctx->h_kdesc = NULL;
    hfg_layout_free(&ctx->lay);
    ctx->have_chunks = 0;
}

/* unmaps every peer mailbox opened by hfg_peer_connect (the own one is a plain allocation) */
static void close_peers(hfg_ctx *ctx) {
    for (int p = 0; p < HFG_MAX_PEERS; p++) {
        if (ctx->peer_box[p] && ctx->peer_box[p] != ctx->d_mailbox) cudaIpcCloseMemHandle(ctx->peer_box[p]);
        ctx->peer_box[p] = NULL;
    }
    ctx->n_ranks = 1;
    ctx->rank = 0;
}

extern "C" void hfg_destroy(hfg_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    free_device(ctx);
    if (ctx->h_params[0]) cudaFreeHost(ctx->h_params[0]);
    if (ctx->d_params[0]) cudaFree(ctx->d_params[0]);
    for (int i = 0; i < STAGE_SLOTS; i++)
        if (ctx->stage_ev[i]) cudaEventDestroy(ctx->stage_ev[i]);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->ev2) cudaEventDestroy(ctx->ev2);
    if (ctx->ev3) cudaEventDestroy(ctx->ev3);
    close_peers(ctx);
    cudaFree(ctx->d_mailbox);
    cudaFree(ctx->d_epoch);
    cudaFree(ctx->d_flush);
    if (ctx->d_nb_table) cudaFree(ctx->d_nb_table);
    if (ctx->h_nb_table) cudaFreeHost(ctx->h_nb_table);
    for (int i = 0; i < 2 * ctx->em_ev_cap; i++) cudaEventDestroy(ctx->em_ev[i]);
    free(ctx->em_ev);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    free(ctx->last_params);
    cudaGetLastError(); /* a context torn down half-built must not leave a sticky-looking error behind */
    free(ctx);
}

/* page-locked host memory: result buffers allocated here are written by the device directly (no staging copy) */
/* the buffers handed out by hfg_host_alloc: the blocking calls recognise them without asking the driver every time */
#define HFG_HOST_REG 64
static struct { void *p; size_t bytes; } g_host_reg[HFG_HOST_REG];
static pthread_mutex_t g_host_reg_lock = PTHREAD_MUTEX_INITIALIZER;

extern "C" void *hfg_host_alloc(size_t bytes) {
    void *p = NULL;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
        cudaGetLastError();
        return NULL;
    }
    pthread_mutex_lock(&g_host_reg_lock);
    for (int i = 0; i < HFG_HOST_REG; i++)
        if (!g_host_reg[i].p) {
            g_host_reg[i].p = p;
            g_host_reg[i].bytes = bytes ? bytes : 1;
            break;
        }
    pthread_mutex_unlock(&g_host_reg_lock);
    return p;
}
extern "C" void hfg_host_free(void *p) {
    if (!p) return;
    pthread_mutex_lock(&g_host_reg_lock);
    for (int i = 0; i < HFG_HOST_REG; i++)
        if (g_host_reg[i].p == p) g_host_reg[i].p = NULL;
    pthread_mutex_unlock(&g_host_reg_lock);
    cudaFreeHost(p);
}
/* 1 when [p, p + bytes) lies inside a live hfg_host_alloc buffer (page-locked and, under unified addressing, device-visible at
 * the same address) */
static int host_reg_contains(const void *p, size_t bytes) {
    int hit = 0;
    pthread_mutex_lock(&g_host_reg_lock);
    for (int i = 0; i < HFG_HOST_REG && !hit; i++) {
        const char *b = (const char *) g_host_reg[i].p;
        if (b && (const char *) p >= b && (const char *) p + bytes <= b + g_host_reg[i].bytes) hit = 1;
    }
    pthread_mutex_unlock(&g_host_reg_lock);
    return hit;
}

extern "C" int64_t hfg_num_windows(const hfg_ctx *ctx) { return ctx && ctx->have_chunks ? ctx->lay.n_windows : 0; }
extern "C" int64_t hfg_kernel_launches(const hfg_ctx *ctx) { return ctx ? ctx->launches : 0; }

extern "C" size_t hfg_stats_device_bytes(const hfg_ctx *ctx) {
    return ctx ? ((size_t) ctx->cfg.n_regions * STATS_DOUBLES + 2) * sizeof(double) : 0;
}

static double wall_ms() {
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return 1e3 * t.tv_sec + 1e-6 * t.tv_nsec;
}

extern "C" int hfg_set_chunks(hfg_ctx *ctx, int32_t n_chunks, const hfg_chunk_desc *chunks, const uint16_t *cov,
                              const uint16_t *cov_high_mapq, const uint16_t *cov_high_clip, const uint8_t *region) {
    if (!ctx) return HFG_ERR_INVALID;
    if (n_chunks < 1 || !chunks || !cov || !cov_high_mapq || !cov_high_clip || !region)
        return fail(ctx, HFG_ERR_INVALID, "hfg_set_chunks: NULL or empty input");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    free_device(ctx);
    const bool timing = getenv("HFG_TIMING") != NULL; /* where the set-up time goes, on stderr */
    /* HFG_HOST_LAYOUT=1: build keys, lists and tiles with the host builder (hfg_layout.c) instead of on the device */
    const bool host_layout = getenv("HFG_HOST_LAYOUT") != NULL;
    double tm[6] = {0};
    tm[0] = wall_ms();
    int64_t W = 0;
    for (int32_t c = 0; c < n_chunks; c++) W += chunks[c].n_windows;
    /* persistent grid: at most one CTA per SM.  The layout picks the shortest segment length that fits the full grid and
     * then keeps only the CTAs those segments fill (profiles/grid_sweep.py: idle CTAs cost more at the grid barriers than
     * they save); HFG_MIN_WPT = minimum windows per thread before another CTA is added (default 1). */
    int wpt = 1;
    if (getenv("HFG_MIN_WPT")) wpt = atoi(getenv("HFG_MIN_WPT")) > 0 ? atoi(getenv("HFG_MIN_WPT")) : 1;
    const int spc = ctx->segs_per_cta;
    int64_t blocks = (W + (int64_t) spc * wpt - 1) / ((int64_t) spc * wpt);
    if (blocks < 1) blocks = 1;
    if (blocks > ctx->max_blocks) blocks = ctx->max_blocks;
    ctx->capacity_arg = (int32_t) blocks * spc;
    hfg_layout_tile_div = ctx->quad == 3 ? 4 : 1; /* v3: four lanes fold the statistics of a tile */
    int rc = hfg_layout_build_ex(&ctx->cfg, n_chunks, chunks, cov, cov_high_mapq, cov_high_clip, region, ctx->capacity_arg,
                                 spc, host_layout ? 0 : 1, &ctx->lay, ctx->err, sizeof(ctx->err));
    if (rc != HFG_OK) return rc;
    ctx->grid = ctx->lay.capacity / spc;
    const int32_t cap = ctx->lay.capacity;
    tm[1] = wall_ms();
    hfg_layout *l = &ctx->lay;
    const size_t slots = (size_t) l->smax * cap;
    const int R = ctx->cfg.n_regions, G = total_gauss_comps(&ctx->cfg);
    const size_t out_doubles = (size_t) R * STATS_DOUBLES + 2;
    const size_t w = (size_t) W;
    /* device-side build: raw inputs + temporaries share the region that holds the per-E-step scratch afterwards */
    const size_t raw_bytes = 3 * hfgl::align256(2 * w) + hfgl::align256(w) + hfgl::align256(sizeof(hfg_chunk_desc) * (size_t) n_chunks) +
                             2 * hfgl::align256(4 * (size_t) n_chunks) + hfgl::align256(4 * (size_t) cap);
    const size_t build_bytes = host_layout ? 0 : raw_bytes + hfgl::temp_bytes(W, NULL);
    /* forward scratch: segment-transposed [smax][4][capacity] (first generation) or window-major [W][4] (quad kernel) */
    /* (v3: the scan stash of a thread, 32 doubles, lives in the thread's own column of the forward scratch, rows 0..31, and is
     * consumed before the thread writes its first forward vector there: one buffer, 24 MB less to keep in L2) */
    size_t ft_rows = (slots > w ? slots : w) * 4;
    if (ctx->quad == 3 && ft_rows < (size_t) 32 * cap) ft_rows = (size_t) 32 * cap;
    const size_t ft_bytes = hfgl::align256(ft_rows * sizeof(double));
    const size_t scratch_bytes = ft_bytes + hfgl::align256(w * 8 * sizeof(double));
    const size_t max_tiles = w / HFG_TILE + w + 1;
    size_t o_scr = 0;
    /* one device allocation carved into the per-run buffers (cudaMalloc is the slow part of a short job) */
    {
        size_t off = 0;
#define CARVE(bytes) (off = (off + 255) & ~(size_t) 255, off += (bytes), off - (bytes))
        const size_t o_wk = CARVE(slots * sizeof(uint32_t));
        const size_t o_wp = CARVE(slots * sizeof(uint32_t));
        const size_t o_wh = CARVE(slots * sizeof(uint32_t));
        const size_t o_hk = CARVE((size_t) (4096 + 3 * HFG_MAX_REGIONS + 1024) * sizeof(int32_t)); /* hot slots (more than any shared
                                                                                                    memory holds), ranges, tile split */

        const size_t o_ss = CARVE(cap * sizeof(int32_t)), o_sl = CARVE(cap * sizeof(int32_t));
        const size_t o_kl = CARVE((w > 0 ? w : 1) * sizeof(int32_t));
        const size_t o_tk = CARVE(max_tiles * sizeof(int32_t)), o_tb = CARVE(max_tiles * sizeof(int32_t));
        const size_t o_tc = CARVE(max_tiles * sizeof(int32_t));
        const size_t o_rt = CARVE((HFG_MAX_REGIONS + 1) * sizeof(int32_t));
        o_scr = CARVE(scratch_bytes > build_bytes ? scratch_bytes : build_bytes);
        const size_t o_bt = CARVE((size_t) ctx->grid * 16 * sizeof(double));
        const size_t o_br = CARVE((size_t) ctx->grid * sizeof(int32_t));
        const size_t o_pa = CARVE((size_t) ctx->grid * R * hfg_nstat(G) * sizeof(double));
        const size_t o_out = CARVE(out_doubles * sizeof(double));
        const size_t o_ll = CARVE(cap * sizeof(double));
        const size_t o_lab = CARVE(w);
        const size_t o_err = CARVE(2 * sizeof(int32_t)); /* error flags, arrival ticket of the grid reduction */
        const size_t o_pc = CARVE((size_t) (ctx->grid + 1) * HFG_PC_STRIDE * sizeof(long long));
        const size_t o_emp = CARVE(sizeof(hfg_region_params) * (size_t) R), o_ems = CARVE(4 * sizeof(int32_t));
        const size_t o_eml = CARVE(sizeof(double) * HFG_EM_LOGLIK_SLOTS);
#undef CARVE
        ctx->d_arena = arena_acquire(off, ctx->device, &ctx->arena_bytes, 0);
        if (!ctx->d_arena) {
            cudaGetLastError();
            return fail(ctx, HFG_ERR_NOMEM, "hfg_set_chunks: cannot allocate %zu bytes of device memory", off);
        }
        if (timing) fprintf(stderr, "[hfg] arena %.1f MB; %lld windows\n", off / 1e6, (long long) W);
        char *base = (char *) ctx->d_arena;
        ctx->d_wkeyT = (uint32_t *) (base + o_wk);
        ctx->d_wposT = (uint32_t *) (base + o_wp);
        ctx->d_wkeyH = (uint32_t *) (base + o_wh);
        ctx->d_hot_key = (int32_t *) (base + o_hk);
        ctx->d_hot_range = ctx->d_hot_key + 4096;
        ctx->d_seg_start = (int32_t *) (base + o_ss);
        ctx->d_seg_len = (int32_t *) (base + o_sl);
        ctx->d_klist = (int32_t *) (base + o_kl);
        ctx->d_tile_key = (int32_t *) (base + o_tk);
        ctx->d_tile_begin = (int32_t *) (base + o_tb);
        ctx->d_tile_cnt = (int32_t *) (base + o_tc);
        ctx->d_region_tile_begin = (int32_t *) (base + o_rt);
        ctx->d_scrFT = (double *) (base + o_scr);
        ctx->d_scan_stash = ctx->d_scrFT;
        ctx->d_scrXB = (double *) (base + o_scr + ft_bytes);
        ctx->d_block_tot = (double *) (base + o_bt);
        ctx->d_block_reset = (int32_t *) (base + o_br);
        ctx->d_partials = (double *) (base + o_pa);
        ctx->d_out = (double *) (base + o_out);
        ctx->d_seg_loglik = (double *) (base + o_ll);
        ctx->d_labels = (int8_t *) (base + o_lab);
        ctx->d_err = (int32_t *) (base + o_err);
        ctx->d_ticket = ctx->d_err + 1;
        ctx->d_phase_clock = (long long *) (base + o_pc);
        ctx->d_em_params = (hfg_region_params *) (base + o_emp);
        ctx->d_em_state = (int32_t *) (base + o_ems);
        ctx->d_em_logliks = (double *) (base + o_eml);
        ctx->em_max = HFG_EM_LOGLIK_SLOTS;
    }
    tm[2] = wall_ms();
    ctx->h_out = (double *) arena_acquire((out_doubles + 2) * sizeof(double), ctx->device, &ctx->h_out_bytes, 2); /* + the completion word */
    ctx->h_labels = (int8_t *) arena_acquire(w, ctx->device, &ctx->h_labels_bytes, 3);
    if (!ctx->h_out || !ctx->h_labels) {
        cudaGetLastError();
        return fail(ctx, HFG_ERR_NOMEM, "hfg_set_chunks: cannot allocate page-locked host memory");
    }
    /* the completion word of the blocking fast path: the block may be a recycled one that still holds an old sequence number */
    *reinterpret_cast<volatile unsigned long long *>(ctx->h_out + out_doubles) = 0ull;
    CU(cudaMemcpyAsync(ctx->d_seg_start, l->seg_start, cap * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_seg_len, l->seg_len, cap * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemsetAsync(ctx->d_labels, 0xff, w, ctx->stream));
    CU(cudaMemsetAsync(ctx->d_err, 0, 2 * sizeof(int32_t), ctx->stream)); /* every launch leaves flags and ticket cleared */

    /* the key block: key table, key descriptors, betas -- sized by the number of keys */
    auto acquire_keyblock = [&](int32_t n_keys) -> int {
        const size_t P = (size_t) n_keys;
        const size_t o_M = 0, o_MT = hfgl::align256(P * 16 * sizeof(double)), o_kd = 2 * o_MT, o_kb = o_kd + hfgl::align256(P * sizeof(uint32_t));
        const size_t o_kh = o_kb + hfgl::align256(P * 3 * sizeof(double));
        const size_t total = o_kh + hfgl::align256(P * sizeof(uint32_t));
        ctx->d_keyblock = arena_acquire(total, ctx->device, &ctx->keyblock_bytes, 1);
        if (!ctx->d_keyblock) {
            cudaGetLastError();
            return fail(ctx, HFG_ERR_NOMEM, "hfg_set_chunks: cannot allocate %zu bytes of device memory for %d keys", total, n_keys);
        }
        char *kb = (char *) ctx->d_keyblock;
        ctx->d_tabM = (double *) (kb + o_M);
        ctx->d_tabMT = (double *) (kb + o_MT);
        ctx->d_kdesc = (uint32_t *) (kb + o_kd);
        ctx->d_kbeta = (double *) (kb + o_kb);
        ctx->d_khot = (uint32_t *) (kb + o_kh);
        return HFG_OK;
    };

    if (host_layout) {
        if ((rc = acquire_keyblock(l->n_keys)) != HFG_OK) return rc;
        CU(cudaMemcpyAsync(ctx->d_wkeyT, l->wkeyT, slots * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(ctx->d_kdesc, l->kdesc, (size_t) l->n_keys * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(ctx->d_kbeta, l->kbeta, (size_t) l->n_keys * 3 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        if (l->n_list > 0)
            CU(cudaMemcpyAsync(ctx->d_klist, l->klist, (size_t) l->n_list * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
        if (l->n_tiles > 0) {
            const size_t tb = (size_t) l->n_tiles * sizeof(int32_t);
            CU(cudaMemcpyAsync(ctx->d_tile_key, l->tile_key, tb, cudaMemcpyHostToDevice, ctx->stream));
            CU(cudaMemcpyAsync(ctx->d_tile_begin, l->tile_begin, tb, cudaMemcpyHostToDevice, ctx->stream));
            CU(cudaMemcpyAsync(ctx->d_tile_cnt, l->tile_cnt, tb, cudaMemcpyHostToDevice, ctx->stream));
        }
        CU(cudaMemcpyAsync(ctx->d_region_tile_begin, l->region_tile_begin, (HFG_MAX_REGIONS + 1) * sizeof(int32_t),
                           cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        tm[3] = tm[4] = wall_ms();
    } else {
        /* raw inputs -> device, then the build (hfg_layout_dev.cuh) */
        char *rb = (char *) ctx->d_arena + o_scr;
        size_t ro = 0;
#define RAW(type, name, bytes) type *name = (type *) (rb + ro); ro += hfgl::align256(bytes)
        RAW(uint16_t, d_cov, 2 * w);
        RAW(uint16_t, d_mapq, 2 * w);
        RAW(uint16_t, d_clip, 2 * w);
        RAW(uint8_t, d_region, w);
        RAW(hfg_chunk_desc, d_chunks, sizeof(hfg_chunk_desc) * (size_t) n_chunks);
        RAW(int32_t, d_head, 4 * (size_t) n_chunks);
        RAW(int32_t, d_tail, 4 * (size_t) n_chunks);
        RAW(int32_t, d_seg_chunk, 4 * (size_t) cap);
#undef RAW
        CU(cudaMemcpyAsync(d_cov, cov, 2 * w, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(d_mapq, cov_high_mapq, 2 * w, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(d_clip, cov_high_clip, 2 * w, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(d_region, region, w, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(d_chunks, chunks, sizeof(hfg_chunk_desc) * (size_t) n_chunks, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(d_head, l->edge_head, 4 * (size_t) n_chunks, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(d_tail, l->edge_tail, 4 * (size_t) n_chunks, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(d_seg_chunk, l->seg_chunk, 4 * (size_t) cap, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemsetAsync(ctx->d_wkeyT, 0, slots * sizeof(uint32_t), ctx->stream));
        tm[3] = wall_ms();
        hfgl::PackArgs pa;
        memset(&pa, 0, sizeof(pa));
        pa.cov = d_cov; pa.mapq = d_mapq; pa.clip = d_clip; pa.region = d_region;
        pa.chunks = d_chunks; pa.edge_head = d_head; pa.edge_tail = d_tail;
        pa.seg_start = ctx->d_seg_start; pa.seg_len = ctx->d_seg_len; pa.seg_chunk = d_seg_chunk;
        pa.n_seg = l->n_seg; pa.capacity = cap;
        pa.adjust_contig_ends = ctx->cfg.adjust_contig_ends; pa.mean_read_length = ctx->cfg.mean_read_length;
        pa.min_read_fraction_at_ends = ctx->cfg.min_read_fraction_at_ends;
        pa.max_high_mapq_ratio = ctx->cfg.max_high_mapq_ratio; pa.min_high_mapq_ratio = ctx->cfg.min_high_mapq_ratio;
        pa.min_highly_clipped_ratio = ctx->cfg.min_highly_clipped_ratio;
        hfgl::DeviceLayoutTemp tmp;
        memset(&tmp, 0, sizeof(tmp));
        CU(hfgl::layout_build_device_keys(pa, W, rb + raw_bytes, ctx->stream, &tmp));
        const double t_keys = wall_ms();
        if ((rc = acquire_keyblock(tmp.P)) != HFG_OK) return rc;
        const double t_kb = wall_ms();
        hfgl::DeviceLayoutOut lo;
        lo.wkeyT = ctx->d_wkeyT; lo.klist = ctx->d_klist;
        lo.tile_key = ctx->d_tile_key; lo.tile_begin = ctx->d_tile_begin; lo.tile_cnt = ctx->d_tile_cnt;
        lo.region_tile_begin = ctx->d_region_tile_begin; lo.kdesc = ctx->d_kdesc; lo.kbeta = ctx->d_kbeta;
        int64_t n_list = 0;
        int32_t n_tiles = 0, tile_len = 0;
        CU(hfgl::layout_build_device_lists(&tmp, l->beta0, lo, ctx->stream, &n_list, &n_tiles, &tile_len));
        CU(cudaStreamSynchronize(ctx->stream));
        l->n_keys = tmp.P;
        l->n_list = n_list;
        l->n_tiles = n_tiles;
        l->tile_len = tile_len;
        ctx->launches += 12; /* the layout kernels (CUB's own passes not counted) */
        tm[4] = wall_ms();
        if (timing)
            fprintf(stderr, "[hfg] device build: pack + sorts + run heads %.2f ms, key block %.2f ms, numbering + lists + tiles %.2f ms\n",
                    t_keys - tm[3], t_kb - t_keys, tm[4] - t_kb);
    }
    if (timing)
        fprintf(stderr, "[hfg] set_chunks (%s keys): segments%s %.2f ms, arena %.2f ms, uploads queued %.2f ms, %s %.2f ms; %d keys, "
                "%d tiles of <= %d windows\n", host_layout ? "host" : "device", host_layout ? " + keys" : "", tm[1] - tm[0],
                tm[2] - tm[1], tm[3] - tm[2], host_layout ? "copies" : "device build", tm[4] - tm[3], l->n_keys, l->n_tiles,
                l->tile_len);
    if (ctx->quad >= 2) {
        /* quad kernel: where every window's statistics record goes -- its position in the key lists -- stored like the key
         * words (segment-transposed); two passes: window -> slot, list position -> slot.  The window -> slot map borrows the
         * forward scratch. */
        uint32_t *wslot = (uint32_t *) ctx->d_scrFT;
        CU(cudaMemsetAsync(ctx->d_wposT, 0xff, slots * sizeof(uint32_t), ctx->stream));
        hfg_wslot_kernel<<<(l->n_seg + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_seg_start, ctx->d_seg_len, l->n_seg, cap, wslot);
        if (l->n_list > 0)
            hfg_wpos_kernel<<<(unsigned) ((l->n_list + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_klist, l->n_list, wslot, ctx->d_wposT);
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(ctx->stream));
        ctx->launches += 2;
        if (ctx->quad == 2) ctx->smem_bytes = smem_bytes_quad(R, G, ctx->threads, l->smax);
    }
    if (ctx->quad == 3) {
        /* the shared-memory matrix table: labels are staged when a segment is short enough, the hot keys get the rest.  The
         * slots go to the regions in proportion to their statistics tiles (~ windows), hottest keys first inside a region;
         * the windows' key words are rewritten (into a copy) to carry the slot instead of the key id. */
        const size_t base = smem_base_v3(R, G, ctx->threads);
        ctx->lab_bytes = l->smax <= HFGQ_LAB_SMAX ? (int32_t) ((((size_t) ctx->threads * l->smax + 16) + 15) & ~(size_t) 15) : 0;
        /* 1 KB: static shared memory, launch reserve; 16 KB are left to the L1 (prefetched cold matrices, key words) */
        const size_t room = ctx->smem_optin - 1024 - 16384 - base - (size_t) ctx->lab_bytes;
        int32_t hot_cap = (int32_t) (room / 128);
        if (hot_cap > 4096) hot_cap = 4096;
        if (getenv("HFG_HOT")) hot_cap = atoi(getenv("HFG_HOT")) < hot_cap ? atoi(getenv("HFG_HOT")) : hot_cap; /* A/B: smaller table */
        ctx->n_hot = hot_cap < l->n_keys ? hot_cap : l->n_keys;
        ctx->smem_bytes = base + (size_t) ctx->lab_bytes + (size_t) ctx->n_hot * 128;
        {
            /* the M-step of the device-resident loop works on shared-memory copies behind the region tables */
            size_t need = ((((size_t) R * QRT_STRIDE(G) + 1) & ~(size_t) 1) * sizeof(double)) + mstep_work_bytes(ctx->threads, 1);
            /* negative binomial: the tail's histogram / estimator scratch (hfg_nb_dev.cuh) */
            if (ctx->nb) {
                int np = 0;
                for (int st = 0; st < HFG_NUM_STATES; st++) np += ctx->cfg.n_comps[st];
                need += (size_t) HFG_NB_TAIL_DOUBLES(np) * sizeof(double);
            }
            if (ctx->smem_bytes < need) ctx->smem_bytes = need;
        }
        CU(cudaMemsetAsync(ctx->d_khot, 0xff, (size_t) l->n_keys * sizeof(uint32_t), ctx->stream));
        hfg_hot_assign_kernel<<<1, 1024, 0, ctx->stream>>>(ctx->d_kdesc, l->n_keys, ctx->d_region_tile_begin, R, ctx->n_hot,
                                                            ctx->d_hot_key, ctx->d_khot, ctx->d_hot_range);
        hfg_hot_rewrite_kernel<<<(unsigned) ((slots + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_wkeyT, ctx->d_khot, slots, ctx->d_wkeyH);
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(ctx->stream));
        ctx->launches += 2;
        if (timing) fprintf(stderr, "[hfg] v3: %d of %d keys in shared memory (%zu bytes), %d bytes of staged labels\n", ctx->n_hot,
                            l->n_keys, (size_t) ctx->n_hot * 128, ctx->lab_bytes);
    }
    if (ctx->nb) {
        /* per-tile pair masses come back to the host, which folds them into the (region, state, x) histogram: it needs
         * every tile's key and every key's observation word */
        const size_t nt = (size_t) (l->n_tiles > 0 ? l->n_tiles : 1), nk = (size_t) (l->n_keys > 0 ? l->n_keys : 1);
        ctx->h_tile_key = (int32_t *) malloc(sizeof(int32_t) * nt);
        ctx->h_kdesc = (uint32_t *) malloc(sizeof(uint32_t) * nk);
        if (!ctx->h_tile_key || !ctx->h_kdesc) return fail(ctx, HFG_ERR_NOMEM, "out of host memory");
        CU(cudaMalloc((void **) &ctx->d_nb_tile_col, sizeof(double) * 4 * nt));
        CU(cudaMallocHost((void **) &ctx->h_nb_tile_col, sizeof(double) * 4 * nt));
        if (l->n_tiles > 0) CU(cudaMemcpy(ctx->h_tile_key, ctx->d_tile_key, sizeof(int32_t) * (size_t) l->n_tiles, cudaMemcpyDeviceToHost));
        if (l->n_keys > 0) CU(cudaMemcpy(ctx->h_kdesc, ctx->d_kdesc, sizeof(uint32_t) * (size_t) l->n_keys, cudaMemcpyDeviceToHost));
        /* device-resident loop (hfg_nb_dev.cuh): the tiles of every (region, coverage bin), in tile order -- the order in which
         * run_blocking_nb folds them -- and lgamma(x + 1) from libm */
        {
            const size_t nbins = (size_t) R * HFG_NB_BINS;
            int32_t *bins = (int32_t *) calloc(nbins + 1 + nt, sizeof(int32_t));
            if (!bins) return fail(ctx, HFG_ERR_NOMEM, "out of host memory");
            int32_t *begin = bins, *tiles = bins + nbins + 1;
            for (int32_t t = 0; t < l->n_tiles; t++) {
                const uint32_t w = ctx->h_kdesc[ctx->h_tile_key[t]];
                const int x = (int) HFG_OBS_X(w), region = (int) HFG_OBS_REGION(w);
                begin[(size_t) region * HFG_NB_BINS + (x < HFG_NB_BINS ? x : HFG_NB_BINS - 1) + 1]++;
            }
            for (size_t b = 0; b < nbins; b++) begin[b + 1] += begin[b];
            int32_t *fill = (int32_t *) malloc(sizeof(int32_t) * (nbins + 1));
            if (!fill) {
                free(bins);
                return fail(ctx, HFG_ERR_NOMEM, "out of host memory");
            }
            memcpy(fill, begin, sizeof(int32_t) * (nbins + 1));
            for (int32_t t = 0; t < l->n_tiles; t++) {
                const uint32_t w = ctx->h_kdesc[ctx->h_tile_key[t]];
                const int x = (int) HFG_OBS_X(w), region = (int) HFG_OBS_REGION(w);
                tiles[fill[(size_t) region * HFG_NB_BINS + (x < HFG_NB_BINS ? x : HFG_NB_BINS - 1)]++] = t;
            }
            free(fill);
            cudaError_t e = cudaMalloc((void **) &ctx->d_nb_bins, sizeof(int32_t) * (nbins + 1 + nt));
            if (e == cudaSuccess) e = cudaMemcpy(ctx->d_nb_bins, bins, sizeof(int32_t) * (nbins + 1 + nt), cudaMemcpyHostToDevice);
            free(bins);
            double lgx1[HFG_NB_TABLE_X];
            for (int x = 0; x < HFG_NB_TABLE_X; x++) lgx1[x] = lgamma(x + 1);
            /* [256] lgamma(x + 1), then the histogram [R][4][256] the grid folds the tile masses into */
            if (e == cudaSuccess) e = cudaMalloc((void **) &ctx->d_nb_lgx1, sizeof(double) * (256 + (size_t) R * 4 * 256));
            if (e == cudaSuccess) e = cudaMemcpy(ctx->d_nb_lgx1, lgx1, sizeof(lgx1), cudaMemcpyHostToDevice);
            if (e != cudaSuccess) return fail(ctx, HFG_ERR_CUDA, "negative-binomial bin lists: %s", cudaGetErrorString(e));
        }
    }
    /* the per-window tables now live on the device */
    free(ctx->lay.obsT); ctx->lay.obsT = NULL;
    free(ctx->lay.wkeyT); ctx->lay.wkeyT = NULL;
    free(ctx->lay.klist); ctx->lay.klist = NULL;
    ctx->have_chunks = 1;
    ctx->have_last = 0;
    return HFG_OK;
}

/* Enqueue one E-step (or forward pass) on `stream`; results land in out_dev = [stats | loglik | error flags]. */
/* kernel arguments of one E-step (or forward pass) reading its parameters from staging slot `slot` */
static void build_args(hfg_ctx *ctx, const double *alpha, double *out_dev, double *post_dev, int forward_only, int slot,
                       EstepArgs *out_args) {
    const hfg_config *cfg = &ctx->cfg;
    const hfg_layout *l = &ctx->lay;
    const int R = cfg->n_regions;
    static const double zero_alpha[16] = {0};
    if (ctx->nb) alpha = zero_alpha; /* the model has no previous-window dependency (hmm_utils.c:1409-1417) */
    hfg_classes cl;
    hfg_classes_build(cfg, alpha, &cl);
    EstepArgs &a = *out_args;
    memset(&a, 0, sizeof(a));
    a.wkeyT = ctx->d_wkeyT;
    a.seg_start = ctx->d_seg_start;
    a.seg_len = ctx->d_seg_len;
    a.n_keys = l->n_keys;
    a.kdesc = ctx->d_kdesc;
    a.kbeta = ctx->d_kbeta;
    a.klist = ctx->d_klist;
    a.tile_key = ctx->d_tile_key;
    a.tile_begin = ctx->d_tile_begin;
    a.tile_cnt = ctx->d_tile_cnt;
    a.region_tile_begin = ctx->d_region_tile_begin;
    a.capacity = l->capacity;
    a.smax = l->smax;
    a.n_regions = R;
    a.beta0 = l->beta0;
    a.n_classes = cl.n_classes;
    int g = 0;
    for (int s = 0; s < HFG_NS; s++) {
        a.is_gauss[s] = cl.is_gaussian[s];
        a.ncomp[s] = cfg->n_comps[s];
        a.gbase[s] = g;
        if (cl.is_gaussian[s]) g += cfg->n_comps[s];
        for (int pre = 0; pre < HFG_NS; pre++) {
            a.cls[pre][s] = cl.cls[pre][s];
            a.alpha[pre][s] = cl.is_gaussian[s] ? alpha[pre * HFG_NS + s] : 0.0;
        }
    }
    a.G = g;
    a.n_tasks = 0;
    a.texp_slot = -1;
    a.slots_start = 0xfu;
    a.slots_other = 0;
    for (int s = 0; s < HFG_NS; s++) {
        for (int pre = 0; pre < HFG_NS; pre++) {
            a.slots_other |= 1u << cl.cls[pre][s];
            a.inv_one_minus_alpha[pre][s] = 1.0 / (1.0 - a.alpha[pre][s]);
            int first = pre;
            for (int q = pre - 1; q >= 0; q--)
                if (cl.cls[q][s] == cl.cls[pre][s]) first = q;
            a.first_pre_of_class[pre][s] = first;
        }
    }
    for (int d = 0; d < HFG_MAX_CLASSES; d++) {
        a.class_state[d] = cl.class_state[d];
        a.class_alpha[d] = cl.class_alpha[d];
    }
    for (int d = 0; d < cl.n_classes; d++) {
        if (!((a.slots_other >> d) & 1u)) continue;
        const int s = cl.class_state[d];
        if (!cl.is_gaussian[s]) {
            a.texp_slot = d;
            continue;
        }
        for (int c = 0; c < cfg->n_comps[s]; c++) {
            a.task_class[a.n_tasks] = (uint8_t) d;
            a.task_comp[a.n_tasks] = (uint8_t) c;
            a.n_tasks++;
        }
    }
    if (a.n_tasks == 0) { /* keep the table non-empty */
        a.task_class[0] = 0;
        a.task_comp[0] = 0;
    }
    a.params = ctx->d_params[slot];
    a.tabM = ctx->d_tabM;
    a.scrFT = ctx->d_scrFT;
    a.scrXB = ctx->d_scrXB;
    a.block_tot = ctx->d_block_tot;
    a.block_reset = ctx->d_block_reset;
    a.partials = ctx->d_partials;
    a.out = out_dev;
    a.out_host = (out_dev == ctx->d_out) ? ctx->h_out : NULL; /* the blocking calls read their results from pinned memory */
    a.seg_loglik = ctx->d_seg_loglik;
    a.labels = ctx->d_labels;
    a.labels_host = NULL;
    a.n_seg = l->n_seg;
    a.n_windows = l->n_windows;
    a.posteriors = post_dev;
    a.err_flags = ctx->d_err;
    a.ticket = ctx->d_ticket;
    a.dbg = ctx->dbg;
    a.tabMT = ctx->d_tabMT;
    a.wposT = ctx->d_wposT;
    if (ctx->quad >= 2)
        a.work_doubles = (int32_t) (ctx->smem_bytes / sizeof(double) - ((((size_t) R * QRT_STRIDE(total_gauss_comps(cfg)) + 1) & ~(size_t) 1)));
    if (ctx->quad == 3) {
        a.wkeyT = ctx->d_wkeyH;
        a.hot_key = ctx->d_hot_key;
        a.hot_range = ctx->d_hot_range;
        a.n_hot = ctx->n_hot;
        a.lab_bytes = ctx->lab_bytes;
        a.scan_stash = ctx->d_scan_stash;
    }
    a.forward_only = forward_only;
    a.phase_clock = ctx->d_phase_clock;
    a.out_doubles = R * STATS_DOUBLES + 2;
    /* the posterior re-run (post_dev != NULL) is local to this rank: no exchange */
    a.n_ranks = (ctx->n_ranks > 1 && post_dev == NULL) ? ctx->n_ranks : 1;
    a.rank = ctx->rank;
    for (int p = 0; p < HFG_MAX_PEERS; p++) a.peer_box[p] = ctx->peer_box[p];
    a.epoch = ctx->d_epoch;
    a.model_type = cfg->model_type;
    a.nb_table = ctx->d_nb_table;
    a.nb_tile_col = ctx->d_nb_tile_col;
    a.nb_bin_begin = ctx->d_nb_bins;
    a.nb_bin_tiles = ctx->d_nb_bins ? ctx->d_nb_bins + (size_t) ctx->cfg.n_regions * HFG_NB_BINS + 1 : NULL;
    a.nb_lgx1 = ctx->d_nb_lgx1;
    a.nb_hist = ctx->d_nb_lgx1 ? ctx->d_nb_lgx1 + 256 : NULL;
}

/* negative binomial: the pmf of every (region, state, x) for these parameters (host, libm: hfg_nb.c) -> pinned staging
 * -> device table, queued on `stream`.  The staging buffer is reused, so the stream is drained first. */
static int nb_upload_table(hfg_ctx *ctx, const hfg_region_params *params, cudaStream_t stream) {
    const int R = ctx->cfg.n_regions;
    double *tmp = (double *) malloc(sizeof(double) * (size_t) R * 4 * HFG_NB_TABLE_X);
    if (!tmp) return fail(ctx, HFG_ERR_NOMEM, "out of host memory");
    const int rc = hfg_nb_emission_table(&ctx->cfg, params, tmp);
    if (rc != HFG_OK) {
        free(tmp);
        return fail(ctx, rc, rc == HFG_ERR_NAN ? "prob is NAN (a negative-binomial pmf evaluated to NaN; hmm_utils.c:507-510)"
                                               : "hfg_nb_emission_table: invalid arguments");
    }
    cudaError_t e = cudaStreamSynchronize(stream);
    for (int q = 0; q < R * 4; q++)
        memcpy(ctx->h_nb_table + (size_t) q * HFG_NB_XSTRIDE, tmp + (size_t) q * HFG_NB_TABLE_X, sizeof(double) * HFG_NB_TABLE_X);
    free(tmp);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(ctx->d_nb_table, ctx->h_nb_table, sizeof(double) * (size_t) R * 4 * HFG_NB_XSTRIDE,
                            cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) return fail(ctx, HFG_ERR_CUDA, "negative-binomial table upload failed: %s", cudaGetErrorString(e));
    return HFG_OK;
}

/* Enqueue one E-step (or forward pass) on `stream`; results land in out_dev = [stats | loglik | error flags]. */
static int enqueue_estep(hfg_ctx *ctx, const double *alpha, const hfg_region_params *params, double *out_dev,
                         double *post_dev, int forward_only, cudaStream_t stream, int timed) {
    if (!ctx->have_chunks) return fail(ctx, HFG_ERR_INVALID, "hfg_set_chunks must be called first");
    if (!alpha || !params) return fail(ctx, HFG_ERR_INVALID, "NULL alpha/params");
    CU(cudaSetDevice(ctx->device));
    const int R = ctx->cfg.n_regions;

    /* stage the parameters (ring of pinned buffers so that back-to-back asynchronous calls never race) */
    const int slot = ctx->stage_next;
    ctx->stage_next = (slot + 1) % STAGE_SLOTS;
    CU(cudaEventSynchronize(ctx->stage_ev[slot]));
    memcpy(ctx->h_params[slot], params, sizeof(hfg_region_params) * (size_t) R);
    CU(cudaMemcpyAsync(ctx->d_params[slot], ctx->h_params[slot], sizeof(hfg_region_params) * (size_t) R,
                       cudaMemcpyHostToDevice, stream));

    EstepArgs a;
    build_args(ctx, alpha, out_dev, post_dev, forward_only, slot, &a);
    if (ctx->nb) {
        const int rc_nb = nb_upload_table(ctx, params, stream);
        if (rc_nb != HFG_OK) return rc_nb;
    }

    void *kargs[] = {(void *) &a};
    if (timed) CU(cudaEventRecord(ctx->ev0, stream));
    CU(cudaLaunchCooperativeKernel(ctx->kernel, dim3(ctx->grid), dim3(ctx->threads), kargs,
                                   ctx->smem_bytes, stream));
    if (timed) {
        CU(cudaEventRecord(ctx->ev1, stream));
        ctx->ev_valid = 1;
    }
    CU(cudaEventRecord(ctx->stage_ev[slot], stream));
    ctx->launches += 1;

    memcpy(ctx->last_alpha, alpha, sizeof(double) * 16);
    memcpy(ctx->last_params, params, sizeof(hfg_region_params) * (size_t) R);
    ctx->have_last = 1;
    return HFG_OK;
}

/* translate [stats | loglik | flags] sitting in the pinned output block */
static int parse_out(hfg_ctx *ctx, hfg_region_stats *stats, double *loglik) {
    const int R = ctx->cfg.n_regions;
    const size_t out_doubles = (size_t) R * STATS_DOUBLES + 2;
    const int flags = (int) ctx->h_out[out_doubles - 1];
    if (flags & 1)
        return fail(ctx, HFG_ERR_SCALE_UNDERFLOW, "scale is very low! (a forward scale fell below 1e-50; hmm.c:412-415)");
    if (flags & 2) return fail(ctx, HFG_ERR_NAN, "[Error] prob is NAN (an emission pdf evaluated to NaN; hmm_utils.c:782-786)");
    if (flags & 4) return fail(ctx, HFG_ERR_CUDA, "peer all-reduce timed out: a rank did not reach the exchange");
    if (stats) memcpy(stats, ctx->h_out, sizeof(hfg_region_stats) * (size_t) R);
    if (loglik) *loglik = ctx->h_out[out_doubles - 2];
    return HFG_OK;
}

/* (re)capture  upload params -> kernel  as one graph (the kernel delivers its results into pinned memory itself) */
static int capture_graph(hfg_ctx *ctx, const EstepArgs *a) {
    const int R = ctx->cfg.n_regions;
    if (ctx->gexec) {
        cudaGraphExecDestroy(ctx->gexec);
        ctx->gexec = NULL;
    }
    cudaGraph_t graph = NULL;
    cudaError_t e = cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal);
    if (e != cudaSuccess) return 0;
    EstepArgs copy = *a;
    void *kargs[] = {(void *) &copy};
    e = cudaMemcpyAsync(ctx->d_params[0], ctx->h_params[0], sizeof(hfg_region_params) * (size_t) R, cudaMemcpyHostToDevice,
                        ctx->stream);
    if (e == cudaSuccess) e = cudaEventRecordWithFlags(ctx->ev0, ctx->stream, cudaEventRecordExternal); /* event-record node */
    if (e == cudaSuccess)
        e = cudaLaunchCooperativeKernel(ctx->kernel, dim3(ctx->grid), dim3(ctx->threads), kargs, ctx->smem_bytes,
                                        ctx->stream);
    if (e == cudaSuccess) e = cudaEventRecordWithFlags(ctx->ev1, ctx->stream, cudaEventRecordExternal);
    /* no read-back nodes: statistics, log-likelihood and flags arrive through the kernel's own stores into pinned memory,
     * and so do the labels (EstepArgs.labels_host: the caller's buffer when it is page-locked, hfg_host_alloc, else the
     * pinned staging buffer) */
    cudaError_t e2 = cudaStreamEndCapture(ctx->stream, &graph);
    if (e == cudaSuccess) e = e2;
    if (e == cudaSuccess) e = cudaGraphInstantiate(&ctx->gexec, graph, 0);
    if (graph) cudaGraphDestroy(graph);
    if (e != cudaSuccess) {
        cudaGetLastError(); /* clear; the direct launch path is used instead */
        ctx->gexec = NULL;
        return 0;
    }
    ctx->graph_args = *a;
    return 1;
}

/* the device-side address of p when p is page-locked host memory the device can write directly, else NULL */
static int8_t *pinned_device_alias(const void *p, size_t bytes) {
    if (host_reg_contains(p, bytes)) return (int8_t *) p; /* hfg_host_alloc: no driver query */
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return NULL;
    }
    return at.type == cudaMemoryTypeHost ? (int8_t *) at.devicePointer : NULL;
}

/* blocking E-step of the negative-binomial model: emission table from the host,
 * the NB kernel instantiation, per-tile pair masses back to the host, folded into the (region, state, x) histogram in
 * tile order and turned into the estimator sums by hfg_nb_stats_from_histogram (hmm.c:615-617,643-649). */
static int run_blocking_nb(hfg_ctx *ctx, const hfg_region_params *params, int forward_only, hfg_region_stats *stats,
                           double *loglik, int8_t *labels) {
    const int with_labels = labels != NULL;
    const int R = ctx->cfg.n_regions;
    const hfg_layout *l = &ctx->lay;
    CU(cudaSetDevice(ctx->device));
    int8_t *label_alias = with_labels ? pinned_device_alias(labels, (size_t) ctx->lay.n_windows) : NULL;
    int8_t *label_dst = !with_labels ? NULL : (label_alias ? label_alias : ctx->h_labels);
    static const double zero_alpha[16] = {0};
    EstepArgs a;
    build_args(ctx, zero_alpha, ctx->d_out, NULL, forward_only, 0, &a);
    a.labels_host = label_dst;
    CU(cudaEventSynchronize(ctx->stage_ev[0]));
    memcpy(ctx->h_params[0], params, sizeof(hfg_region_params) * (size_t) R);
    int rc = nb_upload_table(ctx, params, ctx->stream);
    if (rc != HFG_OK) return rc;
    void *kargs[] = {(void *) &a};
    CU(cudaEventRecord(ctx->ev2, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_params[0], ctx->h_params[0], sizeof(hfg_region_params) * (size_t) R, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaEventRecord(ctx->ev0, ctx->stream));
    CU(cudaLaunchCooperativeKernel(ctx->kernel, dim3(ctx->grid), dim3(ctx->threads), kargs, ctx->smem_bytes, ctx->stream));
    CU(cudaEventRecord(ctx->ev1, ctx->stream));
    /* what comes back: the (region, state, coverage bin) histogram folded by the grid (grid_fold_histogram), [R][4][256]
     * doubles = 8 KB per region, instead of four doubles per statistics tile (585 KB for config 2) and a host fold */
    const size_t hist_doubles = (size_t) R * 4 * 256;
    const int device_fold = ctx->quad == 3; /* the earlier kernel generations (HFG_KERNEL=quad / v1) leave the fold to the host */
    if (!forward_only && l->n_tiles > 0) {
        if (device_fold)
            CU(cudaMemcpyAsync(ctx->h_nb_hist, ctx->d_nb_lgx1 + 256, sizeof(double) * hist_doubles, cudaMemcpyDeviceToHost, ctx->stream));
        else
            CU(cudaMemcpyAsync(ctx->h_nb_tile_col, ctx->d_nb_tile_col, sizeof(double) * 4 * (size_t) l->n_tiles, cudaMemcpyDeviceToHost,
                               ctx->stream));
    }
    CU(cudaEventRecord(ctx->ev3, ctx->stream));
    CU(cudaEventRecord(ctx->stage_ev[0], ctx->stream));
    ctx->ev_valid = ctx->span_valid = 1;
    ctx->launches += 1;
    memcpy(ctx->last_alpha, zero_alpha, sizeof(double) * 16);
    memcpy(ctx->last_params, params, sizeof(hfg_region_params) * (size_t) R);
    ctx->have_last = 1;
    CU(cudaStreamSynchronize(ctx->stream));
    if (with_labels && !label_alias) memcpy(labels, ctx->h_labels, (size_t) l->n_windows);
    rc = parse_out(ctx, stats, loglik); /* transition counts and log-likelihood; the emission slots are zero */
    if (rc != HFG_OK || !stats || forward_only) return rc;
    double *hist = (double *) calloc((size_t) R * 4 * HFG_NB_BINS, sizeof(double));
    if (!hist) return fail(ctx, HFG_ERR_NOMEM, "out of host memory");
    if (device_fold) {
        if (l->n_tiles > 0)
            for (int q = 0; q < R * 4; q++) /* device rows are 256 wide, the estimator's HFG_NB_BINS */
                memcpy(hist + (size_t) q * HFG_NB_BINS, ctx->h_nb_hist + (size_t) q * 256, sizeof(double) * HFG_NB_BINS);
    } else {
        for (int32_t t = 0; t < l->n_tiles; t++) { /* tile order: the same sum on every run */
            const uint32_t w = ctx->h_kdesc[ctx->h_tile_key[t]];
            const int x = (int) HFG_OBS_X(w), region = (int) HFG_OBS_REGION(w);
            const int bin = x < HFG_NB_BINS ? x : HFG_NB_BINS - 1; /* count_data.c:56-64 */
            for (int s = 0; s < 4; s++) hist[((size_t) region * 4 + s) * HFG_NB_BINS + bin] += ctx->h_nb_tile_col[(size_t) t * 4 + s];
        }
    }
    rc = hfg_nb_stats_from_histogram(&ctx->cfg, params, hist, stats);
    free(hist);
    if (rc != HFG_OK) return fail(ctx, rc, "prob is NAN (negative-binomial estimator update)");
    return HFG_OK;
}

/* blocking E-step with host buffers: graph replay when possible, plain launches otherwise (same kernel either way) */
static int run_blocking(hfg_ctx *ctx, const double *alpha, const hfg_region_params *params, int forward_only,
                        hfg_region_stats *stats, double *loglik, int8_t *labels) {
    if (!ctx->have_chunks) return fail(ctx, HFG_ERR_INVALID, "hfg_set_chunks must be called first");
    if (!alpha || !params) return fail(ctx, HFG_ERR_INVALID, "NULL alpha/params");
    if (ctx->nb) return run_blocking_nb(ctx, params, forward_only, stats, loglik, labels);
    const int with_labels = labels != NULL;
    CU(cudaSetDevice(ctx->device));
    /* where the kernel streams the labels: the caller's buffer when it is page-locked, else the pinned staging buffer */
    int8_t *label_alias = with_labels ? pinned_device_alias(labels, (size_t) ctx->lay.n_windows) : NULL;
    int8_t *label_dst = !with_labels ? NULL : (label_alias ? label_alias : ctx->h_labels);
    const int R = ctx->cfg.n_regions;
    EstepArgs a;
    build_args(ctx, alpha, ctx->d_out, NULL, forward_only, 0, &a);
    a.labels_host = label_dst;
    if (ctx->quad == 3 && R == 1 && !ctx->timing) {
        /* Fast path (single-region models, the default kernel): ONE launch and nothing else on the stream.  The parameters
         * travel in the kernel arguments (no upload node, no staging slot), no events are recorded, and the host does not
         * synchronise the stream: the kernel's tail stores a sequence number behind the results it writes into the pinned
         * block (after a system-wide fence that also covers the labels every CTA streamed to host memory), and the host
         * polls that word.  The stream is queried now and then so that a failed launch ends the wait. */
        const size_t out_doubles = (size_t) R * STATS_DOUBLES + 2;
        volatile unsigned long long *flag = reinterpret_cast<volatile unsigned long long *>(ctx->h_out + out_doubles);
        static unsigned long long g_done_seq = 0; /* process-wide: a recycled block never sees a number twice */
        const unsigned long long seq = ctx->done_seq = __atomic_add_fetch(&g_done_seq, 1ull, __ATOMIC_RELAXED);
        a.params_inline = 1;
        a.inl_params = params[0];
        a.done_flag = const_cast<unsigned long long *>(flag);
        a.done_seq = seq;
        void *kargs[] = {(void *) &a};
        CU(cudaLaunchCooperativeKernel(ctx->kernel, dim3(ctx->grid), dim3(ctx->threads), kargs, ctx->smem_bytes, ctx->stream));
        ctx->ev_valid = ctx->span_valid = 0;
        ctx->launches += 1;
        memcpy(ctx->last_alpha, alpha, sizeof(double) * 16);
        memcpy(ctx->last_params, params, sizeof(hfg_region_params) * (size_t) R);
        ctx->have_last = 1;
        for (unsigned spins = 1; *flag != seq; spins++) {
            if ((spins & 0x3fffu) == 0 && cudaStreamQuery(ctx->stream) != cudaErrorNotReady) break;
            __builtin_ia32_pause();
        }
        if (*flag != seq) { /* the stream is idle or broken and the word never came */
            CU(cudaStreamSynchronize(ctx->stream));
            if (*flag != seq) return fail(ctx, HFG_ERR_CUDA, "the E-step kernel ended without delivering its results");
        }
        __atomic_thread_fence(__ATOMIC_ACQUIRE); /* the results are read after the word that announces them */
        if (with_labels && !label_alias) memcpy(labels, ctx->h_labels, (size_t) ctx->lay.n_windows);
        return parse_out(ctx, stats, loglik);
    }
    CU(cudaEventSynchronize(ctx->stage_ev[0])); /* slot 0 may still feed an asynchronous device-variant call */
    memcpy(ctx->h_params[0], params, sizeof(hfg_region_params) * (size_t) R);
    if (!ctx->graph_disabled) {
        if (!ctx->gexec || memcmp(&a, &ctx->graph_args, sizeof(a)) != 0) {
            if (!capture_graph(ctx, &a)) ctx->graph_disabled = 1;
        }
    }
    CU(cudaEventRecord(ctx->ev2, ctx->stream));
    if (ctx->gexec && !ctx->graph_disabled) {
        CU(cudaGraphLaunch(ctx->gexec, ctx->stream));
    } else {
        /* plain launches of the same sequence */
        void *kargs[] = {(void *) &a};
        CU(cudaMemcpyAsync(ctx->d_params[0], ctx->h_params[0], sizeof(hfg_region_params) * (size_t) R, cudaMemcpyHostToDevice,
                           ctx->stream));
        CU(cudaEventRecord(ctx->ev0, ctx->stream));
        CU(cudaLaunchCooperativeKernel(ctx->kernel, dim3(ctx->grid), dim3(ctx->threads), kargs, ctx->smem_bytes, ctx->stream));
        CU(cudaEventRecord(ctx->ev1, ctx->stream));
    }
    CU(cudaEventRecord(ctx->ev3, ctx->stream));
    CU(cudaEventRecord(ctx->stage_ev[0], ctx->stream));
    ctx->ev_valid = ctx->span_valid = 1;
    ctx->launches += 1;
    memcpy(ctx->last_alpha, alpha, sizeof(double) * 16);
    memcpy(ctx->last_params, params, sizeof(hfg_region_params) * (size_t) R);
    ctx->have_last = 1;
    CU(cudaStreamSynchronize(ctx->stream));
    if (with_labels && !label_alias) memcpy(labels, ctx->h_labels, (size_t) ctx->lay.n_windows);
    return parse_out(ctx, stats, loglik);
}

extern "C" int hfg_em_iteration(hfg_ctx *ctx, const double *alpha, const hfg_region_params *params,
                                hfg_region_stats *stats, double *loglik, int8_t *labels) {
    if (!ctx) return HFG_ERR_INVALID;
    return run_blocking(ctx, alpha, params, 0, stats, loglik, labels);
}

extern "C" int hfg_forward_only(hfg_ctx *ctx, const double *alpha, const hfg_region_params *params, double *loglik) {
    if (!ctx) return HFG_ERR_INVALID;
    return run_blocking(ctx, alpha, params, 1, NULL, loglik, NULL);
}

extern "C" int hfg_em_iteration_device(hfg_ctx *ctx, const double *alpha, const hfg_region_params *params,
                                       void *stats_dev, void *stream) {
    if (!ctx) return HFG_ERR_INVALID;
    if (!stats_dev) return fail(ctx, HFG_ERR_INVALID, "hfg_em_iteration_device: NULL stats_dev");
    if (ctx->nb) return fail(ctx, HFG_ERR_INVALID, "hfg_em_iteration_device: not available for the negative-binomial model");
    cudaStream_t s = stream ? (cudaStream_t) stream : ctx->stream;
    return enqueue_estep(ctx, alpha, params, (double *) stats_dev, NULL, 0, s, 1);
}

extern "C" int hfg_get_labels(hfg_ctx *ctx, int8_t *labels) {
    if (!ctx || !labels) return HFG_ERR_INVALID;
    if (!ctx->have_last) return fail(ctx, HFG_ERR_INVALID, "no E-step has run yet");
    CU(cudaSetDevice(ctx->device));
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(labels, ctx->d_labels, (size_t) ctx->lay.n_windows, cudaMemcpyDeviceToHost));
    return HFG_OK;
}

extern "C" int hfg_get_posteriors(hfg_ctx *ctx, double *posteriors) {
    if (!ctx || !posteriors) return HFG_ERR_INVALID;
    if (!ctx->have_last) return fail(ctx, HFG_ERR_INVALID, "no E-step has run yet");
    CU(cudaSetDevice(ctx->device));
    const size_t bytes = (size_t) ctx->lay.n_windows * 4 * sizeof(double);
    if (!ctx->d_post) CU(cudaMalloc((void **) &ctx->d_post, bytes));
    /* the E-step is deterministic: re-running it with the same parameters reproduces the same f^, b^, scales */
    hfg_region_params *p = (hfg_region_params *) malloc(sizeof(hfg_region_params) * (size_t) ctx->cfg.n_regions);
    if (!p) return fail(ctx, HFG_ERR_NOMEM, "out of host memory");
    double al[16];
    memcpy(p, ctx->last_params, sizeof(hfg_region_params) * (size_t) ctx->cfg.n_regions);
    memcpy(al, ctx->last_alpha, sizeof(al));
    int rc = enqueue_estep(ctx, al, p, ctx->d_out, ctx->d_post, 0, ctx->stream, 0);
    free(p);
    if (rc != HFG_OK) return rc;
    CU(cudaMemcpyAsync(posteriors, ctx->d_post, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return HFG_OK;
}

extern "C" int hfg_get_chunk_logliks(hfg_ctx *ctx, double *logliks) {
    if (!ctx || !logliks) return HFG_ERR_INVALID;
    if (!ctx->have_last) return fail(ctx, HFG_ERR_INVALID, "no E-step has run yet");
    CU(cudaSetDevice(ctx->device));
    const hfg_layout *l = &ctx->lay;
    double *seg = (double *) malloc(sizeof(double) * (size_t) l->capacity);
    if (!seg) return fail(ctx, HFG_ERR_NOMEM, "out of host memory");
    CU(cudaDeviceSynchronize());
    cudaError_t e = cudaMemcpy(seg, ctx->d_seg_loglik, sizeof(double) * (size_t) l->capacity, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) {
        free(seg);
        return fail(ctx, HFG_ERR_CUDA, "cudaMemcpy failed: %s", cudaGetErrorString(e));
    }
    for (int c = 0; c < l->n_chunks; c++) logliks[c] = 0.0;
    for (int j = 0; j < l->n_seg; j++) logliks[l->seg_chunk[j]] += seg[j]; /* segments are in window order */
    free(seg);
    return HFG_OK;
}

extern "C" double hfg_last_call_device_ms(hfg_ctx *ctx) {
    if (!ctx || !ctx->span_valid) return -1.0;
    float ms = 0.f;
    if (cudaEventSynchronize(ctx->ev3) != cudaSuccess) return -1.0;
    if (cudaEventElapsedTime(&ms, ctx->ev2, ctx->ev3) != cudaSuccess) return -1.0;
    return (double) ms;
}

extern "C" double hfg_last_estep_kernel_ms(hfg_ctx *ctx) {
    if (!ctx || !ctx->ev_valid) return -1.0;
    float ms = 0.f;
    if (cudaEventSynchronize(ctx->ev1) != cudaSuccess) return -1.0;
    if (cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1) != cudaSuccess) return -1.0;
    return (double) ms;
}

/* ---- device-resident EM loop ------------------------------------------------------------------------------------------ */

extern "C" int hfg_em_begin(hfg_ctx *ctx, const double *alpha, const hfg_region_params *params, double convergence_tol,
                            int max_esteps) {
    if (!ctx) return HFG_ERR_INVALID;
    if (!ctx->have_chunks) return fail(ctx, HFG_ERR_INVALID, "hfg_set_chunks must be called first");
    if (!alpha || !params || max_esteps < 1) return fail(ctx, HFG_ERR_INVALID, "hfg_em_begin: bad argument");
    if (ctx->nb && ctx->quad != 3)
        return fail(ctx, HFG_ERR_INVALID, "hfg_em_begin: the negative-binomial model needs the default kernel for the device-resident loop");
    CU(cudaSetDevice(ctx->device));
    const int R = ctx->cfg.n_regions;
    const size_t pb = sizeof(hfg_region_params) * (size_t) R;
    if (max_esteps > HFG_EM_LOGLIK_SLOTS)
        return fail(ctx, HFG_ERR_INVALID, "hfg_em_begin: at most %d E-steps per loop", HFG_EM_LOGLIK_SLOTS);
    if (ctx->em_ev_cap < max_esteps) {
        cudaEvent_t *ne = (cudaEvent_t *) realloc(ctx->em_ev, sizeof(cudaEvent_t) * 2 * (size_t) max_esteps);
        if (!ne) return fail(ctx, HFG_ERR_NOMEM, "out of host memory");
        ctx->em_ev = ne;
        while (ctx->em_ev_cap < max_esteps) { /* pair by pair, so that a failure leaves only complete pairs to destroy */
            const int i = 2 * ctx->em_ev_cap;
            CU(cudaEventCreate(&ctx->em_ev[i]));
            cudaError_t e = cudaEventCreate(&ctx->em_ev[i + 1]);
            if (e != cudaSuccess) {
                cudaEventDestroy(ctx->em_ev[i]);
                return fail(ctx, HFG_ERR_CUDA, "cudaEventCreate failed: %s", cudaGetErrorString(e));
            }
            ctx->em_ev_cap++;
        }
    }
    CU(cudaEventSynchronize(ctx->stage_ev[0]));
    memcpy(ctx->h_params[0], params, pb);
    CU(cudaMemcpyAsync(ctx->d_em_params, ctx->h_params[0], pb, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaEventRecord(ctx->stage_ev[0], ctx->stream));
    CU(cudaMemsetAsync(ctx->d_em_state, 0, 4 * sizeof(int32_t), ctx->stream));
    CU(cudaMemsetAsync(ctx->d_err, 0, sizeof(int32_t), ctx->stream));
    memcpy(ctx->em_alpha, alpha, sizeof(double) * 16);
    ctx->em_tol = convergence_tol;
    ctx->em_limit = max_esteps;
    ctx->em_enqueued = 0;
    ctx->em_active = 1;
    return HFG_OK;
}

extern "C" int hfg_em_enqueue(hfg_ctx *ctx, int final_pass) {
    if (!ctx) return HFG_ERR_INVALID;
    if (!ctx->em_active) return fail(ctx, HFG_ERR_INVALID, "hfg_em_enqueue: call hfg_em_begin first");
    if (ctx->em_enqueued >= ctx->em_limit) return fail(ctx, HFG_ERR_INVALID, "hfg_em_enqueue: more than max_esteps iterations");
    CU(cudaSetDevice(ctx->device));
    EstepArgs a;
    build_args(ctx, ctx->em_alpha, ctx->d_out, NULL, 0, 0, &a);
    a.params = ctx->d_em_params;
    a.em_params = ctx->d_em_params;
    a.out_host = NULL; /* the loop's results are fetched once, by hfg_em_finish */
    a.em_mode = final_pass ? 2 : 1;
    a.em_tol = ctx->em_tol;
    a.em_state = ctx->d_em_state;
    a.em_logliks = ctx->d_em_logliks;
    a.em_max_logliks = ctx->em_limit;
    void *kargs[] = {(void *) &a};
    const int i = ctx->em_enqueued;
    CU(cudaEventRecord(ctx->em_ev[2 * i], ctx->stream));
    CU(cudaLaunchCooperativeKernel(ctx->kernel, dim3(ctx->grid), dim3(ctx->threads), kargs, ctx->smem_bytes, ctx->stream));
    CU(cudaEventRecord(ctx->em_ev[2 * i + 1], ctx->stream));
    ctx->em_enqueued = i + 1;
    ctx->launches += 1;
    return HFG_OK;
}

extern "C" int hfg_em_finish(hfg_ctx *ctx, hfg_region_params *params, double *logliks, int *n_esteps, int *converged,
                             int8_t *labels) {
    if (!ctx) return HFG_ERR_INVALID;
    if (!ctx->em_active) return fail(ctx, HFG_ERR_INVALID, "hfg_em_finish: call hfg_em_begin first");
    CU(cudaSetDevice(ctx->device));
    const int R = ctx->cfg.n_regions;
    const size_t pb = sizeof(hfg_region_params) * (size_t) R;
    int32_t state[4] = {0, 0, 0, 0};
    if (labels) CU(cudaMemcpyAsync(ctx->h_labels, ctx->d_labels, (size_t) ctx->lay.n_windows, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(ctx->h_params[0], ctx->d_em_params, pb, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaMemcpy(state, ctx->d_em_state, sizeof(state), cudaMemcpyDeviceToHost));
    ctx->em_active = 0;
    const int n = state[1] < ctx->em_limit ? state[1] : ctx->em_limit;
    if (logliks && n > 0) CU(cudaMemcpy(logliks, ctx->d_em_logliks, sizeof(double) * (size_t) n, cudaMemcpyDeviceToHost));
    if (n_esteps) *n_esteps = n;
    if (converged) *converged = state[3];
    if (params) memcpy(params, ctx->h_params[0], pb);
    if (labels) memcpy(labels, ctx->h_labels, (size_t) ctx->lay.n_windows);
    /* the parameters of the last E-step that ran, for hfg_get_posteriors */
    memcpy(ctx->last_alpha, ctx->em_alpha, sizeof(double) * 16);
    memcpy(ctx->last_params, ctx->h_params[0], pb);
    ctx->have_last = 1;
    if (state[2] & 1)
        return fail(ctx, HFG_ERR_SCALE_UNDERFLOW, "scale is very low! (a forward scale fell below 1e-50; hmm.c:412-415)");
    if (state[2] & 2) return fail(ctx, HFG_ERR_NAN, "[Error] prob is NAN (an emission pdf evaluated to NaN; hmm_utils.c:782-786)");
    if (state[2] & 4) return fail(ctx, HFG_ERR_CUDA, "peer all-reduce timed out: a rank did not reach the exchange");
    return HFG_OK;
}

extern "C" double hfg_em_enqueued_ms(hfg_ctx *ctx, int i) {
    if (!ctx || i < 0 || i >= ctx->em_enqueued) return -1.0;
    float ms = 0.f;
    if (cudaEventSynchronize(ctx->em_ev[2 * i + 1]) != cudaSuccess) return -1.0;
    if (cudaEventElapsedTime(&ms, ctx->em_ev[2 * i], ctx->em_ev[2 * i + 1]) != cudaSuccess) return -1.0;
    return (double) ms;
}

extern "C" int hfg_debug_set_timing(hfg_ctx *ctx, int on) {
    if (!ctx) return HFG_ERR_INVALID;
    ctx->timing = on != 0;
    return HFG_OK;
}

extern "C" int hfg_debug_l2_flush(hfg_ctx *ctx, size_t bytes) {
    if (!ctx || bytes == 0) return HFG_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    if (ctx->flush_bytes < bytes) {
        cudaFree(ctx->d_flush);
        ctx->d_flush = NULL;
        ctx->flush_bytes = 0;
        CU(cudaMalloc(&ctx->d_flush, bytes));
        ctx->flush_bytes = bytes;
    }
    CU(cudaMemsetAsync(ctx->d_flush, 0x5a, bytes, ctx->stream));
    return HFG_OK;
}

extern "C" int hfg_debug_blocking_steps(hfg_ctx *ctx, const double *alpha, hfg_region_params *params, hfg_region_stats *stats,
                                        int8_t *labels, int n_steps, size_t flush_bytes, double convergence_tol,
                                        double *step_seconds, double *logliks) {
    if (!ctx || !alpha || !params || !stats || !step_seconds || !logliks || n_steps < 0) return HFG_ERR_INVALID;
    for (int i = 0; i < n_steps; i++) {
        if (flush_bytes) {
            int rc = hfg_debug_l2_flush(ctx, flush_bytes);
            if (rc != HFG_OK) return rc;
            CU(cudaStreamSynchronize(ctx->stream));
        }
        struct timespec t0, t1;
        clock_gettime(CLOCK_MONOTONIC, &t0);
        int rc = hfg_em_iteration(ctx, alpha, params, stats, &logliks[i], labels);
        int converged = 0;
        if (rc == HFG_OK) rc = hfg_mstep(&ctx->cfg, params, stats, convergence_tol, &converged);
        clock_gettime(CLOCK_MONOTONIC, &t1);
        if (rc != HFG_OK) return rc;
        step_seconds[i] = (double) (t1.tv_sec - t0.tv_sec) + 1e-9 * (double) (t1.tv_nsec - t0.tv_nsec);
    }
    return HFG_OK;
}

/* while (iter <= numberOfIterations && converged == false) { E-step; M-step }  (src/hmm_flagger.c:337-431), then the final
 * inference with the final parameters (:464) -- all of it queued at once on the device (hfg_em_*). */
extern "C" int hfg_run_em(hfg_ctx *ctx, const double *alpha, hfg_region_params *params, int max_iterations,
                          double convergence_tol, double *logliks, int *n_esteps, int8_t *labels) {
    if (!ctx || !params || !logliks || !n_esteps) return HFG_ERR_INVALID;
    if (max_iterations < 0) max_iterations = 0;
    if (max_iterations + 1 > HFG_EM_LOGLIK_SLOTS || (ctx->nb && getenv("HFG_NB_HOST_LOOP"))) {
        /* longer than the device loop records (or HFG_NB_HOST_LOOP=1: the negative-binomial model with the host's libm /
         * long-double estimator update between the iterations, the A/B partner of the device-resident loop): the same loop
         * with the host between the iterations */
        const int R = ctx->cfg.n_regions;
        hfg_region_stats *stats = (hfg_region_stats *) malloc(sizeof(hfg_region_stats) * (size_t) R);
        if (!stats) return fail(ctx, HFG_ERR_NOMEM, "out of host memory");
        int iter = 1, converged = 0, k = 0, rc = HFG_OK;
        while (iter <= max_iterations && !converged) {
            if ((rc = hfg_em_iteration(ctx, alpha, params, stats, &logliks[k], NULL)) != HFG_OK) break;
            k++;
            if ((rc = hfg_mstep(&ctx->cfg, params, stats, convergence_tol, &converged)) != HFG_OK) break;
            iter++;
        }
        if (rc == HFG_OK && (rc = hfg_em_iteration(ctx, alpha, params, stats, &logliks[k], labels)) == HFG_OK) k++;
        *n_esteps = k;
        free(stats);
        return rc;
    }
    int rc = hfg_em_begin(ctx, alpha, params, convergence_tol, max_iterations + 1);
    for (int it = 0; it < max_iterations && rc == HFG_OK; it++) rc = hfg_em_enqueue(ctx, 0);
    if (rc == HFG_OK) rc = hfg_em_enqueue(ctx, 1);
    if (rc != HFG_OK) {
        ctx->em_active = 0;
        cudaStreamSynchronize(ctx->stream);
        return rc;
    }
    return hfg_em_finish(ctx, params, logliks, n_esteps, NULL, labels);
}

/* ---- batched EM runs (include/hfg.h) ------------------------------------------------------------------------------------ */

struct hfg_batch {
    int n_lanes;
    hfg_ctx *lane[16];
    char err[512];
};

extern "C" int hfg_set_max_blocks(hfg_ctx *ctx, int max_blocks) {
    if (!ctx || max_blocks < 0) return HFG_ERR_INVALID;
    ctx->max_blocks = max_blocks == 0 || max_blocks > ctx->num_sms ? ctx->num_sms : max_blocks;
    return HFG_OK;
}

extern "C" const char *hfg_batch_last_error(const hfg_batch *b) { return b ? b->err : g_create_err; }

static int batch_fail(hfg_batch *b, int rc, const hfg_ctx *from) {
    snprintf(b->err, sizeof(b->err), "%s", hfg_last_error(from));
    return rc;
}

extern "C" void hfg_batch_destroy(hfg_batch *b) {
    if (!b) return;
    for (int i = 0; i < b->n_lanes; i++) hfg_destroy(b->lane[i]);
    free(b);
}

extern "C" int hfg_batch_create(hfg_batch **out, const hfg_config *cfg, int n_lanes) {
    if (!out || !cfg) return fail(NULL, HFG_ERR_INVALID, "hfg_batch_create: NULL argument");
    *out = NULL;
    if (n_lanes < 1 || n_lanes > 16) return fail(NULL, HFG_ERR_INVALID, "hfg_batch_create: n_lanes %d outside 1..16", n_lanes);
    if (cfg->model_type == HFG_MODEL_NEGATIVE_BINOMIAL)
        return fail(NULL, HFG_ERR_INVALID, "hfg_batch_create: the negative-binomial model has no device-resident loop to batch");
    hfg_batch *b = (hfg_batch *) calloc(1, sizeof(hfg_batch));
    if (!b) return fail(NULL, HFG_ERR_NOMEM, "hfg_batch_create: out of memory");
    for (int i = 0; i < n_lanes; i++) {
        const int rc = hfg_create(&b->lane[i], cfg);
        if (rc != HFG_OK) {
            hfg_batch_destroy(b);
            return rc;
        }
        b->n_lanes = i + 1;
        const int per = b->lane[i]->num_sms / n_lanes;
        hfg_set_max_blocks(b->lane[i], per < 1 ? 1 : per);
    }
    *out = b;
    return HFG_OK;
}

extern "C" int hfg_batch_set_chunks(hfg_batch *b, int32_t n_chunks, const hfg_chunk_desc *chunks, const uint16_t *cov,
                                    const uint16_t *cov_high_mapq, const uint16_t *cov_high_clip, const uint8_t *region) {
    if (!b) return HFG_ERR_INVALID;
    for (int i = 0; i < b->n_lanes; i++) {
        const int rc = hfg_set_chunks(b->lane[i], n_chunks, chunks, cov, cov_high_mapq, cov_high_clip, region);
        if (rc != HFG_OK) return batch_fail(b, rc, b->lane[i]);
    }
    return HFG_OK;
}

extern "C" int hfg_batch_run_em(hfg_batch *b, int n_runs, const double *alphas, hfg_region_params *params, int max_iterations,
                                double convergence_tol, double *logliks, int *n_esteps, int8_t *labels) {
    if (!b || n_runs < 0 || !alphas || !params || !logliks || !n_esteps) return HFG_ERR_INVALID;
    if (max_iterations < 0) max_iterations = 0;
    if (max_iterations + 1 > HFG_EM_LOGLIK_SLOTS) {
        snprintf(b->err, sizeof(b->err), "hfg_batch_run_em: at most %d iterations per run", HFG_EM_LOGLIK_SLOTS - 1);
        return HFG_ERR_INVALID;
    }
    const int R = b->lane[0]->cfg.n_regions;
    const int64_t W = b->lane[0]->lay.n_windows;
    int rc_all = HFG_OK;
    /* waves of n_lanes runs: every lane's whole loop is queued on its stream before any lane is waited for */
    for (int r0 = 0; r0 < n_runs; r0 += b->n_lanes) {
        const int n = n_runs - r0 < b->n_lanes ? n_runs - r0 : b->n_lanes;
        int rc_lane[16];
        for (int l = 0; l < n; l++) {
            hfg_ctx *ctx = b->lane[l];
            const int r = r0 + l;
            int rc = hfg_em_begin(ctx, alphas + (size_t) r * 16, params + (size_t) r * R, convergence_tol, max_iterations + 1);
            for (int it = 0; it < max_iterations && rc == HFG_OK; it++) rc = hfg_em_enqueue(ctx, 0);
            if (rc == HFG_OK) rc = hfg_em_enqueue(ctx, 1);
            rc_lane[l] = rc;
        }
        for (int l = 0; l < n; l++) {
            hfg_ctx *ctx = b->lane[l];
            const int r = r0 + l;
            int rc = rc_lane[l];
            if (rc != HFG_OK) {
                ctx->em_active = 0;
                cudaStreamSynchronize(ctx->stream);
            } else {
                rc = hfg_em_finish(ctx, params + (size_t) r * R, logliks + (size_t) r * (max_iterations + 1), &n_esteps[r], NULL,
                                   labels ? labels + (size_t) r * W : NULL);
            }
            if (rc != HFG_OK && rc_all == HFG_OK) rc_all = batch_fail(b, rc, ctx);
        }
    }
    return rc_all;
}

/* One outer iteration of the `acceleration` branch of runHMMFlagger (src/hmm_flagger.c:344-416) up to, and excluding,
 * its closing M-step: E(p0) -> M -> E(p1) -> M -> p' by SquareAccelerator_getModelPrime (hmm.c:885-918; forward-only
 * passes choose the step length) -> E(p').  In: params = p0.  Out: params = p', stats = statistics of E(p'),
 * *loglik0 = log-likelihood of p0, *alpha_rate = accepted step. */
extern "C" int hfg_squarem_iteration(hfg_ctx *ctx, const double *alpha, hfg_region_params *params, hfg_region_stats *stats,
                                     double convergence_tol, double *loglik0, double *alpha_rate) {
    if (!ctx || !params || !stats || !loglik0) return HFG_ERR_INVALID;
    const int R = ctx->cfg.n_regions;
    const size_t pb = sizeof(hfg_region_params) * (size_t) R;
    hfg_region_params *p0 = (hfg_region_params *) malloc(4 * pb);
    if (!p0) return fail(ctx, HFG_ERR_NOMEM, "out of host memory");
    hfg_region_params *p1 = p0 + R, *p2 = p1 + R, *prime = p2 + R;
    int rc = HFG_OK, ignored = 0, weights_failed = 0;
    double ll0 = 0.0, llp = 0.0, rate = -1.0;
    do {
        if ((rc = hfg_em_iteration(ctx, alpha, params, stats, &ll0, NULL)) != HFG_OK) break;
        memcpy(p0, params, pb);
        memcpy(p1, params, pb);
        if ((rc = hfg_mstep(&ctx->cfg, p1, stats, convergence_tol, &ignored)) != HFG_OK) break;
        if ((rc = hfg_em_iteration(ctx, alpha, p1, stats, &llp, NULL)) != HFG_OK) break;
        memcpy(p2, p1, pb);
        if ((rc = hfg_mstep(&ctx->cfg, p2, stats, convergence_tol, &ignored)) != HFG_OK) break;
        rate = hfg_squarem_alpha_rate(&ctx->cfg, p0, p1, p2);
        rc = hfg_squarem_prime(&ctx->cfg, p0, p1, p2, rate, prime);
        while (rc == HFG_OK && !hfg_params_feasible(&ctx->cfg, prime))
            rc = hfg_squarem_shrink(&ctx->cfg, p0, p1, p2, 1e-2, &rate, prime);
        if (rc != HFG_OK) { weights_failed = 1; break; }
        if ((rc = hfg_forward_only(ctx, alpha, prime, &llp)) != HFG_OK) break;
        while (llp < ll0) { /* shrink the step until the likelihood does not drop below p0's */
            do {
                rc = hfg_squarem_shrink(&ctx->cfg, p0, p1, p2, 1e-2, &rate, prime);
            } while (rc == HFG_OK && !hfg_params_feasible(&ctx->cfg, prime));
            if (rc != HFG_OK) { weights_failed = 1; break; }
            if ((rc = hfg_forward_only(ctx, alpha, prime, &llp)) != HFG_OK) break;
        }
        if (rc != HFG_OK) break;
        if ((rc = hfg_em_iteration(ctx, alpha, prime, stats, &llp, NULL)) != HFG_OK) break;
        memcpy(params, prime, pb);
    } while (0);
    if (weights_failed) fail(ctx, rc, "SQUAREM: the extrapolated mixture weights do not sum to > 0");
    *loglik0 = ll0;
    if (alpha_rate) *alpha_rate = rate;
    free(p0);
    return rc;
}

/* The accelerated EM loop: while (iter <= n && !converged) { squarem iteration; M-step }, then the final inference. */
extern "C" int hfg_run_em_accelerated(hfg_ctx *ctx, const double *alpha, hfg_region_params *params, int max_iterations,
                                      double convergence_tol, double *logliks, double *alpha_rates, int *n_outer,
                                      int8_t *labels) {
    if (!ctx || !params || !logliks || !n_outer) return HFG_ERR_INVALID;
    hfg_region_stats *stats = (hfg_region_stats *) malloc(sizeof(hfg_region_stats) * (size_t) ctx->cfg.n_regions);
    if (!stats) return fail(ctx, HFG_ERR_NOMEM, "out of host memory");
    int iter = 1, converged = 0, k = 0, rc = HFG_OK;
    while (iter <= max_iterations && !converged) {
        double rate = 0.0;
        if ((rc = hfg_squarem_iteration(ctx, alpha, params, stats, convergence_tol, &logliks[k], &rate)) != HFG_OK) break;
        if (alpha_rates) alpha_rates[k] = rate;
        k++;
        if ((rc = hfg_mstep(&ctx->cfg, params, stats, convergence_tol, &converged)) != HFG_OK) break;
        iter++;
    }
    if (rc == HFG_OK) rc = hfg_em_iteration(ctx, alpha, params, stats, &logliks[k], labels); /* final inference (:464) */
    *n_outer = k;
    free(stats);
    return rc;
}

/* ---- instrumentation / test hooks (not part of the reference-facing surface) ------------------------------------- */

/* clock64() of thread 0 of every CTA at the six phase boundaries of the last E-step kernel: start, end of phase A,
 * arrival at the grid barrier, release, end of C1 (thread 0 only), end of phase C, arrival at the second barrier (block
 * reduction done), its release, and the SM id.  out: [grid][12]; slots 9 and 10 of block 0: totals reduced, M-step done. */
extern "C" int hfg_debug_phase_clocks(hfg_ctx *ctx, long long *out, int *grid) {
    if (!ctx || !out || !grid || !ctx->have_chunks) return HFG_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(out, ctx->d_phase_clock, (size_t) (ctx->grid + 1) * HFG_PC_STRIDE * sizeof(long long), cudaMemcpyDeviceToHost));
    *grid = ctx->grid;
    return HFG_OK;
}

__global__ void hfg_debug_exp_kernel(const double *in, double *out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = hfgk::exp_nonpos(in[i]);
}

/* Applies the kernel's exp_nonpos() to n host values (accuracy test of the custom exponential). */
extern "C" int hfg_debug_exp(hfg_ctx *ctx, const double *in, double *out, int n) {
    if (!ctx || !in || !out || n < 1) return HFG_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    double *d_in = NULL, *d_out = NULL;
    const size_t bytes = sizeof(double) * (size_t) n;
    cudaError_t e = cudaMalloc((void **) &d_in, bytes);
    if (e == cudaSuccess) e = cudaMalloc((void **) &d_out, bytes);
    if (e == cudaSuccess) e = cudaMemcpy(d_in, in, bytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        hfg_debug_exp_kernel<<<(n + 255) / 256, 256>>>(d_in, d_out, n);
        ctx->launches += 1;
        e = cudaMemcpy(out, d_out, bytes, cudaMemcpyDeviceToHost);
    }
    cudaFree(d_in);
    cudaFree(d_out);
    if (e != cudaSuccess) return fail(ctx, HFG_ERR_CUDA, "hfg_debug_exp: %s", cudaGetErrorString(e));
    return HFG_OK;
}

/* Test hook: rebuilds keys, lists and tiles with the HOST builder (hfg_layout.c) for the inputs this context was given and
 * compares them, bit for bit, with what lives on the device (normally built there, hfg_layout_dev.cuh).  Returns HFG_OK
 * when identical; otherwise HFG_ERR_INVALID with the first differing table named in hfg_last_error(). */
extern "C" int hfg_debug_layout_compare(hfg_ctx *ctx, int32_t n_chunks, const hfg_chunk_desc *chunks, const uint16_t *cov,
                                        const uint16_t *cov_high_mapq, const uint16_t *cov_high_clip, const uint8_t *region) {
    if (!ctx || !ctx->have_chunks) return HFG_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    CU(cudaDeviceSynchronize());
    hfg_layout h;
    char err[256];
    hfg_layout_tile_div = ctx->quad == 3 ? 4 : 1;
    int rc = hfg_layout_build_ex(&ctx->cfg, n_chunks, chunks, cov, cov_high_mapq, cov_high_clip, region, ctx->capacity_arg,
                                 ctx->segs_per_cta, 0, &h, err, sizeof(err));
    if (rc != HFG_OK) return fail(ctx, rc, "host layout build failed: %s", err);
    const hfg_layout *d = &ctx->lay;
    const char *bad = NULL;
    if (h.capacity != d->capacity || h.smax != d->smax || h.n_seg != d->n_seg || h.n_windows != d->n_windows) bad = "segmentation";
    else if (h.n_keys != d->n_keys) bad = "number of keys";
    else if (h.n_list != d->n_list) bad = "list length";
    else if (h.n_tiles != d->n_tiles || h.tile_len != d->tile_len) bad = "tile count / length";
    if (!bad) {
        const size_t slots = (size_t) h.smax * h.capacity, NT = (size_t) h.n_tiles, P = (size_t) h.n_keys;
        size_t cap_bytes = slots * 4;
        if (P * 24 > cap_bytes) cap_bytes = P * 24;
        if ((size_t) h.n_list * 4 > cap_bytes) cap_bytes = (size_t) h.n_list * 4;
        if (NT * 4 > cap_bytes) cap_bytes = NT * 4;
        if (cap_bytes < 1024) cap_bytes = 1024;
        void *buf = malloc(cap_bytes);
        if (!buf) {
            hfg_layout_free(&h);
            return fail(ctx, HFG_ERR_NOMEM, "out of host memory");
        }
        struct { const char *name; const void *dev; const void *host; size_t bytes; } tabs[] = {
            {"wkeyT", ctx->d_wkeyT, h.wkeyT, slots * 4},
            {"kdesc", ctx->d_kdesc, h.kdesc, P * 4},
            {"kbeta", ctx->d_kbeta, h.kbeta, P * 24},
            {"klist", ctx->d_klist, h.klist, (size_t) h.n_list * 4},
            {"tile_key", ctx->d_tile_key, h.tile_key, NT * 4},
            {"tile_begin", ctx->d_tile_begin, h.tile_begin, NT * 4},
            {"tile_cnt", ctx->d_tile_cnt, h.tile_cnt, NT * 4},
            {"region_tile_begin", ctx->d_region_tile_begin, h.region_tile_begin, (HFG_MAX_REGIONS + 1) * 4},
        };
        for (size_t i = 0; i < sizeof(tabs) / sizeof(tabs[0]) && !bad; i++) {
            if (tabs[i].bytes == 0) continue;
            if (cudaMemcpy(buf, tabs[i].dev, tabs[i].bytes, cudaMemcpyDeviceToHost) != cudaSuccess) {
                bad = "cudaMemcpy";
                cudaGetLastError();
            } else if (memcmp(buf, tabs[i].host, tabs[i].bytes) != 0) {
                bad = tabs[i].name;
            }
        }
        free(buf);
    }
    if (bad)
        fail(ctx, HFG_ERR_INVALID, "device layout differs from the host builder's: %s (device: %d keys, %lld listed, %d tiles of %d; "
             "host: %d keys, %lld listed, %d tiles of %d)", bad, d->n_keys, (long long) d->n_list, d->n_tiles, d->tile_len, h.n_keys,
             (long long) h.n_list, h.n_tiles, h.tile_len);
    hfg_layout_free(&h);
    return bad ? HFG_ERR_INVALID : HFG_OK;
}

/* ---- multi-GPU: peer exchange set-up ------------------------------------------------------------------------------- */

static size_t mailbox_bytes(const hfg_ctx *ctx) {
    const size_t n = (size_t) ctx->cfg.n_regions * STATS_DOUBLES + 2;
    /* slots, arrival counters, barrier counters (first-generation protocol: data, fence, flag), then the flagged slots of the
     * current one: [2 epochs][HFG_MAX_PEERS][n][2] 64-bit words, each 32 bits of data + the 32-bit epoch (hfg_estep_tail) */
    return (2 * HFG_MAX_PEERS * n + 3 * HFG_MAX_PEERS + 4 * HFG_MAX_PEERS * n) * sizeof(double);
}

extern "C" size_t hfg_peer_handle_bytes(void) { return sizeof(cudaIpcMemHandle_t); }

extern "C" int hfg_peer_export(hfg_ctx *ctx, void *handle_out) {
    if (!ctx || !handle_out) return HFG_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    if (!ctx->d_mailbox) {
        CU(cudaMalloc((void **) &ctx->d_mailbox, mailbox_bytes(ctx)));
        CU(cudaMemset(ctx->d_mailbox, 0, mailbox_bytes(ctx)));
    }
    if (!ctx->d_epoch) {
        CU(cudaMalloc((void **) &ctx->d_epoch, 2 * sizeof(unsigned long long))); /* exchanges done, barriers done */
        CU(cudaMemset(ctx->d_epoch, 0, 2 * sizeof(unsigned long long)));
    }
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, ctx->d_mailbox));
    memcpy(handle_out, &h, sizeof(h));
    return HFG_OK;
}

/* Device-side rendezvous of the ranks on the context's stream: every rank bumps a counter in every peer's mailbox and waits
 * for the others' bumps in its own.  Benchmarks put it in front of a timed E-step so that the interval starts at the same
 * moment on every rank (the all-reduce inside the kernel would otherwise charge the fastest rank with the others' lateness). */
__global__ void hfg_peer_barrier_kernel(EstepArgs A, size_t bar_base) {
    const int p = threadIdx.x;
    unsigned long long *mine_epoch = A.epoch + 1;
    const unsigned long long e = *mine_epoch + 1;
    if (p < A.n_ranks) {
        volatile unsigned long long *c = (volatile unsigned long long *) (A.peer_box[p] + bar_base) + A.rank;
        *c = e;
        __threadfence_system();
        volatile unsigned long long *mine = (volatile unsigned long long *) (A.peer_box[A.rank] + bar_base) + p;
        const long long t0 = clock64();
        while (*mine < e)
            if (clock64() - t0 > 4000000000LL) break; /* ~2 s: a lost peer must not hang the stream */
    }
    __syncthreads();
    if (p == 0) *mine_epoch = e;
}

extern "C" int hfg_peer_barrier(hfg_ctx *ctx) {
    if (!ctx) return HFG_ERR_INVALID;
    if (ctx->n_ranks < 2) return HFG_OK;
    CU(cudaSetDevice(ctx->device));
    EstepArgs a;
    memset(&a, 0, sizeof(a));
    a.n_ranks = ctx->n_ranks;
    a.rank = ctx->rank;
    a.epoch = ctx->d_epoch;
    for (int p = 0; p < HFG_MAX_PEERS; p++) a.peer_box[p] = ctx->peer_box[p];
    const size_t n = (size_t) ctx->cfg.n_regions * STATS_DOUBLES + 2;
    hfg_peer_barrier_kernel<<<1, 32, 0, ctx->stream>>>(a, 2 * HFG_MAX_PEERS * n + 2 * HFG_MAX_PEERS);
    CU(cudaGetLastError());
    ctx->launches += 1;
    return HFG_OK;
}

extern "C" int hfg_peer_connect(hfg_ctx *ctx, int n_ranks, int rank, const void *handles) {
    if (!ctx || !handles) return HFG_ERR_INVALID;
    if (n_ranks < 1 || n_ranks > HFG_MAX_PEERS || rank < 0 || rank >= n_ranks)
        return fail(ctx, HFG_ERR_INVALID, "hfg_peer_connect: %d ranks (limit %d), rank %d", n_ranks, HFG_MAX_PEERS, rank);
    if (!ctx->d_mailbox) return fail(ctx, HFG_ERR_INVALID, "hfg_peer_connect: call hfg_peer_export first");
    if (ctx->nb && n_ranks > 1) return fail(ctx, HFG_ERR_INVALID, "hfg_peer_connect: the negative-binomial model is single-GPU so far");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    close_peers(ctx); /* a second connect replaces the first; a failure below leaves the context single-rank */
    if (ctx->gexec) { /* the captured graph was built for the previous set of peers */
        cudaGraphExecDestroy(ctx->gexec);
        ctx->gexec = NULL;
    }
    for (int p = 0; p < n_ranks; p++) {
        if (p == rank) {
            ctx->peer_box[p] = ctx->d_mailbox;
            continue;
        }
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char *) handles + (size_t) p * sizeof(h), sizeof(h));
        void *ptr = NULL;
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            close_peers(ctx);
            return fail(ctx, HFG_ERR_CUDA, "cudaIpcOpenMemHandle for rank %d failed: %s", p, cudaGetErrorString(e));
        }
        ctx->peer_box[p] = (double *) ptr;
    }
    ctx->n_ranks = n_ranks;
    ctx->rank = rank;
    return HFG_OK;
}
