"""Python host side of libhfg: a thin ctypes binding over the C-ABI of include/hfg.h.

Mirrors the reference's operator surface for this path (names and meaning follow
submodules/hmm/hmm.h): `em_iteration` == EM_runOneIterationForList, `forward_only` ==
EM_runForwardForList, `mstep` == HMM_estimateParameters, `model_init` == createModel/HMM_construct.
There is no fallback of any kind: if libhfg.so is missing or CUDA is unusable, this raises.
"""
import ctypes as C
import os

import numpy as np

from . import _abi
from ._abi import ptr

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libhfg.so")
_lib = None


class HfgError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"libhfg error {code}: {message}")
        self.code = code


def lib():
    """Load libhfg.so (built in-tree by __graft_entry__.build() / flagger_b200/csrc/Makefile)."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise HfgError(_abi.ERR_CUDA, f"{_LIB_PATH} is missing: build it with `make -C flagger_b200/csrc` "
                                          "(there is no CPU fallback)")
        L = C.CDLL(_LIB_PATH)
        L.hfg_last_error.restype = C.c_char_p
        L.hfg_last_error.argtypes = [C.c_void_p]
        L.hfg_batch_last_error.restype = C.c_char_p
        L.hfg_batch_last_error.argtypes = [C.c_void_p]
        L.hfg_batch_destroy.restype = None
        L.hfg_batch_destroy.argtypes = [C.c_void_p]
        L.hfg_num_windows.restype = C.c_int64
        L.hfg_kernel_launches.restype = C.c_int64
        L.hfg_last_estep_kernel_ms.restype = C.c_double
        L.hfg_last_call_device_ms.restype = C.c_double
        L.hfg_stats_device_bytes.restype = C.c_size_t
        for name in ("hfg_num_windows", "hfg_kernel_launches", "hfg_last_estep_kernel_ms", "hfg_last_call_device_ms",
                     "hfg_stats_device_bytes",
                     "hfg_destroy"):
            getattr(L, name).argtypes = [C.c_void_p]
        L.hfg_destroy.restype = None
        _lib = L
    return _lib


EXPORTED_SYMBOLS = (
    "hfg_create", "hfg_destroy", "hfg_last_error", "hfg_set_chunks", "hfg_num_windows", "hfg_em_iteration",
    "hfg_forward_only", "hfg_get_posteriors", "hfg_get_chunk_logliks", "hfg_em_iteration_device",
    "hfg_stats_device_bytes", "hfg_get_labels", "hfg_best_num_collapsed_comps", "hfg_model_init", "hfg_mstep",
    "hfg_host_alloc", "hfg_host_free", "hfg_run_em", "hfg_em_begin", "hfg_em_enqueue", "hfg_em_finish",
    "hfg_em_enqueued_ms", "hfg_debug_l2_flush", "hfg_debug_set_timing", "hfg_debug_blocking_steps", "hfg_kernel_launches", "hfg_last_estep_kernel_ms",
    "hfg_last_call_device_ms", "hfg_debug_phase_clocks", "hfg_debug_exp", "hfg_debug_layout_check", "hfg_debug_beta",
    "hfg_debug_layout_compare", "hfg_peer_handle_bytes", "hfg_peer_export", "hfg_peer_connect", "hfg_peer_barrier", "hfg_read_cov",
    "hfg_read_bin", "hfg_cov_free", "hfg_write_summary_tsv", "hfg_benchmark_scores", "hfg_params_feasible", "hfg_squarem_alpha_rate",
    "hfg_squarem_prime", "hfg_squarem_shrink", "hfg_squarem_iteration", "hfg_run_em_accelerated",
    "hfg_release_cached_memory", "hfg_device_warmup", "hfg_batch_create", "hfg_batch_set_chunks",
    "hfg_batch_run_em", "hfg_batch_last_error", "hfg_batch_destroy", "hfg_set_max_blocks", "hfg_nb_emission_table", "hfg_nb_stats_from_histogram", "hfg_digammal", "hfg_debug_gunzip",
)


def layout_check(cfg, wl, capacity):
    """Host-only self-check of the segment layout and of the observation keys (returns (ok, summary[n_seg, smax, n_edge,
    windows, keys, tiles]))."""
    summary = np.zeros(6, np.int64)
    chunks = np.ascontiguousarray(wl.chunks)
    rc = lib().hfg_debug_layout_check(ptr(np.ascontiguousarray(cfg)), C.c_int32(len(chunks)), ptr(chunks),
                                      ptr(np.ascontiguousarray(wl.cov, np.uint16)),
                                      ptr(np.ascontiguousarray(wl.cov_high_mapq, np.uint16)),
                                      ptr(np.ascontiguousarray(wl.cov_high_clip, np.uint16)),
                                      ptr(np.ascontiguousarray(wl.region, np.uint8)), C.c_int32(capacity), ptr(summary))
    return rc == 0, summary


def beta(cfg, chunk_desc, window):
    f = lib().hfg_debug_beta
    f.restype = C.c_double
    return float(f(ptr(np.ascontiguousarray(cfg)), ptr(np.ascontiguousarray(chunk_desc)), C.c_int(window)))


class PinnedArray:
    """A numpy view (`.array`) over page-locked host memory from hfg_host_alloc: result buffers the device writes
    directly.  Freed with .free() or on garbage collection."""

    def __init__(self, n, dtype=np.int8):
        L = lib()
        L.hfg_host_alloc.restype = C.c_void_p
        nbytes = int(n) * np.dtype(dtype).itemsize
        self._p = L.hfg_host_alloc(C.c_size_t(nbytes))
        if not self._p:
            raise HfgError(_abi.ERR_NOMEM, "hfg_host_alloc failed (no usable CUDA device?)")
        self.array = np.frombuffer((C.c_char * nbytes).from_address(self._p), dtype=dtype, count=int(n))

    def free(self):
        if self._p:
            self.array = None
            lib().hfg_host_free(C.c_void_p(self._p))
            self._p = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def best_num_collapsed_comps(max_coverage, region_coverages):
    rc = np.ascontiguousarray(region_coverages, np.int32)
    return int(lib().hfg_best_num_collapsed_comps(C.c_int(int(max_coverage)), ptr(rc), C.c_int(len(rc))))


def model_init(cfg, region_coverages, window_len, start_only=False):
    params = np.zeros(int(cfg["n_regions"][0]), dtype=_abi.region_params_dtype)
    rc_ = np.ascontiguousarray(region_coverages, np.int32)
    rc = lib().hfg_model_init(ptr(cfg), ptr(rc_), C.c_int(int(window_len)), C.c_int(1 if start_only else 0),
                              ptr(params))
    if rc != 0:
        raise HfgError(rc, "hfg_model_init: invalid configuration")
    return params


def mstep(cfg, params, stats, tol=1e-3):
    params = params.copy()
    conv = C.c_int(0)
    rc = lib().hfg_mstep(ptr(cfg), ptr(params), ptr(stats), C.c_double(tol), C.byref(conv))
    if rc != 0:
        raise HfgError(rc, "hfg_mstep: invalid arguments")
    return params, bool(conv.value)


def params_feasible(cfg, params):
    """HMM_isFeasible (hmm.c:80-87)."""
    return bool(lib().hfg_params_feasible(ptr(cfg), ptr(np.ascontiguousarray(params))))


def squarem(cfg, p0, p1, p2, n_shrinks=0, margin=1e-2):
    """The SQUAREM candidate of SquareAccelerator (hmm.c:820-1098): step length from (p0, p1, p2), extrapolated and
    renormalised parameters, then `n_shrinks` step halvings.  Returns (prime, alpha_rate, feasible)."""
    L = lib()
    L.hfg_squarem_alpha_rate.restype = C.c_double
    p0, p1, p2 = (np.ascontiguousarray(p) for p in (p0, p1, p2))
    rate = C.c_double(L.hfg_squarem_alpha_rate(ptr(cfg), ptr(p0), ptr(p1), ptr(p2)))
    prime = np.zeros_like(p0)
    rc = L.hfg_squarem_prime(ptr(cfg), ptr(p0), ptr(p1), ptr(p2), rate, ptr(prime))
    for _ in range(n_shrinks):
        if rc != 0:
            break
        rc = L.hfg_squarem_shrink(ptr(cfg), ptr(p0), ptr(p1), ptr(p2), C.c_double(margin), C.byref(rate), ptr(prime))
    if rc != 0:
        raise HfgError(rc, "hfg_squarem: the extrapolated mixture weights do not sum to > 0")
    return prime, rate.value, params_feasible(cfg, prime)


def nb_emission_table(cfg, params):
    """Negative-binomial pmf of every (region, state, x): [R, 4, 251] (hfg_nb_emission_table; host only)."""
    R = int(cfg["n_regions"][0])
    table = np.zeros((R, 4, _abi.NB_TABLE_X))
    rc = lib().hfg_nb_emission_table(ptr(cfg), ptr(np.ascontiguousarray(params)), ptr(table))
    if rc != 0:
        raise HfgError(rc, "hfg_nb_emission_table: invalid arguments or a NaN pmf")
    return table


def nb_stats_from_histogram(cfg, params, histogram, stats=None):
    """theta / lambda / weight estimator sums from the [R, 4, 250] pair-mass histogram (hfg_nb_stats_from_histogram)."""
    R = int(cfg["n_regions"][0])
    hist = np.ascontiguousarray(histogram, np.float64)
    assert hist.shape == (R, 4, _abi.NB_BINS)
    stats = np.zeros(R, dtype=_abi.region_stats_dtype) if stats is None else stats
    rc = lib().hfg_nb_stats_from_histogram(ptr(cfg), ptr(np.ascontiguousarray(params)), ptr(hist), ptr(stats))
    if rc != 0:
        raise HfgError(rc, "hfg_nb_stats_from_histogram: invalid arguments or a NaN pmf")
    return stats


def digamma(x):
    """hfg_digammal(x) as (hi, lo) doubles with hi + lo the exact long-double result."""
    f = lib().hfg_digammal
    f.restype, f.argtypes = C.c_longdouble, [C.c_longdouble]
    v = np.longdouble(f(np.longdouble(x)))
    hi = np.float64(v)
    return float(hi), float(v - np.longdouble(hi))


class HmmFlaggerBatch:
    """hfg_batch: `n_lanes` contexts over the same chunks on one GPU, each on num_SMs / n_lanes CTAs; run_em fits many
    (alpha, start parameters) candidates, n_lanes at a time (the alpha tuner's inner loop, tune_alpha_hmm_flagger.py:243-267)."""

    def __init__(self, cfg, workload, n_lanes=8):
        self.cfg = np.ascontiguousarray(cfg)
        self._h = C.c_void_p()
        rc = lib().hfg_batch_create(C.byref(self._h), ptr(self.cfg), C.c_int(int(n_lanes)))
        if rc != 0:
            raise HfgError(rc, lib().hfg_last_error(None).decode())
        self.n_lanes = int(n_lanes)
        self.n_regions = int(self.cfg["n_regions"][0])
        chunks = np.ascontiguousarray(workload.chunks)
        self._check(lib().hfg_batch_set_chunks(self._h, C.c_int32(len(chunks)), ptr(chunks),
                                               ptr(np.ascontiguousarray(workload.cov, np.uint16)),
                                               ptr(np.ascontiguousarray(workload.cov_high_mapq, np.uint16)),
                                               ptr(np.ascontiguousarray(workload.cov_high_clip, np.uint16)),
                                               ptr(np.ascontiguousarray(workload.region, np.uint8))))
        self.n_windows = int(workload.n_windows)

    def _check(self, rc):
        if rc != 0:
            raise HfgError(rc, lib().hfg_batch_last_error(self._h).decode())

    def run_em(self, alphas, params, max_iterations, tol=1e-3, want_labels=True):
        """alphas [n][4][4]; params: one start set for all runs, or a list of n.  Returns (list of params, list of logliks,
        labels [n][W] or None), run r exactly what HmmFlaggerGPU.run_em(alphas[r], params[r], ...) returns."""
        alphas = np.ascontiguousarray(np.asarray(alphas, np.float64).reshape(-1, 16))
        n = alphas.shape[0]
        plist = list(params) if isinstance(params, (list, tuple)) else [params] * n
        pall = np.ascontiguousarray(np.concatenate([np.ascontiguousarray(p).reshape(-1) for p in plist]))
        logliks = np.zeros((n, max_iterations + 1), np.float64)
        n_esteps = np.zeros(n, np.int32)
        labels = np.empty((n, self.n_windows), np.int8) if want_labels else None
        self._check(lib().hfg_batch_run_em(self._h, C.c_int(n), ptr(alphas), ptr(pall), C.c_int(int(max_iterations)),
                                           C.c_double(tol), ptr(logliks), ptr(n_esteps), ptr(labels)))
        R = self.n_regions
        return ([pall[r * R:(r + 1) * R].copy() for r in range(n)], [logliks[r, :n_esteps[r]].copy() for r in range(n)], labels)

    def close(self):
        if self._h:
            lib().hfg_batch_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class HmmFlaggerGPU:
    """One libhfg context: a set of chunks resident on one GPU."""

    def __init__(self, cfg, workload=None, timing=False):
        """timing=True: the blocking calls record CUDA events (last_estep_kernel_ms / last_call_device_ms) and take the
        graph path; by default a single-region model takes the one-launch fast path (hfg_api.cu::run_blocking)."""
        self.cfg = np.ascontiguousarray(cfg)
        self._h = C.c_void_p()
        rc = lib().hfg_create(C.byref(self._h), ptr(self.cfg))
        if rc != 0:
            raise HfgError(rc, lib().hfg_last_error(None).decode())
        if timing:
            lib().hfg_debug_set_timing(self._h, C.c_int(1))
        self.n_regions = int(self.cfg["n_regions"][0])
        self.n_windows = 0
        self.n_chunks = 0
        if workload is not None:
            self.set_chunks(workload)

    def _check(self, rc):
        if rc != 0:
            raise HfgError(rc, lib().hfg_last_error(self._h).decode())

    def close(self):
        if self._h:
            lib().hfg_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_chunks(self, wl):
        chunks = np.ascontiguousarray(wl.chunks)
        self._check(lib().hfg_set_chunks(self._h, C.c_int32(len(chunks)), ptr(chunks),
                                         ptr(np.ascontiguousarray(wl.cov, np.uint16)),
                                         ptr(np.ascontiguousarray(wl.cov_high_mapq, np.uint16)),
                                         ptr(np.ascontiguousarray(wl.cov_high_clip, np.uint16)),
                                         ptr(np.ascontiguousarray(wl.region, np.uint8))))
        self.n_windows = int(lib().hfg_num_windows(self._h))
        self.n_chunks = len(chunks)

    def layout_matches_host(self, wl):
        """Are the device-built keys / lists / tiles bit-identical to the host builder's for this workload?  (ok, message)"""
        chunks = np.ascontiguousarray(wl.chunks)
        rc = lib().hfg_debug_layout_compare(self._h, C.c_int32(len(chunks)), ptr(chunks),
                                            ptr(np.ascontiguousarray(wl.cov, np.uint16)),
                                            ptr(np.ascontiguousarray(wl.cov_high_mapq, np.uint16)),
                                            ptr(np.ascontiguousarray(wl.cov_high_clip, np.uint16)),
                                            ptr(np.ascontiguousarray(wl.region, np.uint8)))
        return rc == 0, lib().hfg_last_error(self._h).decode()

    def em_iteration(self, alpha, params, want_labels=True, stats=None, labels=None):
        """EM_runOneIterationForList: returns (stats, loglik, labels)."""
        alpha = np.ascontiguousarray(alpha, np.float64)
        if stats is None:
            stats = np.zeros(self.n_regions, dtype=_abi.region_stats_dtype)
        if labels is None and want_labels:
            labels = np.empty(self.n_windows, np.int8)
        ll = C.c_double(0.0)
        self._check(lib().hfg_em_iteration(self._h, ptr(alpha), ptr(params), ptr(stats), C.byref(ll),
                                           ptr(labels) if want_labels else None))
        return stats, ll.value, labels

    def forward_only(self, alpha, params):
        """EM_runForwardForList: returns the total log-likelihood."""
        alpha = np.ascontiguousarray(alpha, np.float64)
        ll = C.c_double(0.0)
        self._check(lib().hfg_forward_only(self._h, ptr(alpha), ptr(params), C.byref(ll)))
        return ll.value

    def em_iteration_device(self, alpha, params, stats_dev_ptr, stream_ptr=0):
        alpha = np.ascontiguousarray(alpha, np.float64)
        self._check(lib().hfg_em_iteration_device(self._h, ptr(alpha), ptr(params), C.c_void_p(stats_dev_ptr),
                                                  C.c_void_p(stream_ptr)))

    def peer_connect(self, dist):
        """Wire this context to the contexts of all other ranks of a torch.distributed group (one process per GPU):
        exchanges the CUDA-IPC mailbox handles with an all-gather, after which every E-step call returns statistics
        and log-likelihood summed over the ranks by the kernel itself."""
        import torch
        L = lib()
        L.hfg_peer_handle_bytes.restype = C.c_size_t
        nb = int(L.hfg_peer_handle_bytes())
        mine = np.zeros(nb, np.uint8)
        self._check(L.hfg_peer_export(self._h, ptr(mine)))
        world, rank = dist.get_world_size(), dist.get_rank()
        dev = torch.device("cuda", int(self.cfg["device"][0])) if dist.get_backend() == "nccl" else torch.device("cpu")
        gathered = [torch.zeros(nb, dtype=torch.uint8, device=dev) for _ in range(world)]
        dist.all_gather(gathered, torch.from_numpy(mine).to(dev))
        handles = np.ascontiguousarray(np.concatenate([g.cpu().numpy() for g in gathered]))
        self._check(L.hfg_peer_connect(self._h, C.c_int(world), C.c_int(rank), ptr(handles)))
        dist.barrier()

    def peer_barrier(self):
        """Device-side rendezvous of all ranks on the library's stream (hfg_peer_barrier)."""
        self._check(lib().hfg_peer_barrier(self._h))

    def stats_device_bytes(self):
        return int(lib().hfg_stats_device_bytes(self._h))

    def labels(self):
        out = np.empty(self.n_windows, np.int8)
        self._check(lib().hfg_get_labels(self._h, ptr(out)))
        return out

    def posteriors(self):
        out = np.empty((self.n_windows, 4), np.float64)
        self._check(lib().hfg_get_posteriors(self._h, ptr(out)))
        return out

    def chunk_logliks(self):
        out = np.empty(self.n_chunks, np.float64)
        self._check(lib().hfg_get_chunk_logliks(self._h, ptr(out)))
        return out

    def run_em(self, alpha, params, max_iterations, tol=1e-3, want_labels=True):
        """The EM loop of runHMMFlagger (src/hmm_flagger.c:337-467): returns (params, logliks, labels)."""
        alpha = np.ascontiguousarray(alpha, np.float64)
        params = params.copy()
        logliks = np.zeros(max_iterations + 1, np.float64)
        n = C.c_int(0)
        labels = np.empty(self.n_windows, np.int8) if want_labels else None
        self._check(lib().hfg_run_em(self._h, ptr(alpha), ptr(params), C.c_int(max_iterations), C.c_double(tol),
                                     ptr(logliks), C.byref(n), ptr(labels)))
        return params, logliks[:n.value].copy(), labels

    # ---- device-resident EM loop (hfg_em_begin / hfg_em_enqueue / hfg_em_finish) ----
    def em_begin(self, alpha, params, tol=1e-3, max_esteps=64):
        alpha = np.ascontiguousarray(alpha, np.float64)
        self._em_max = int(max_esteps)
        self._check(lib().hfg_em_begin(self._h, ptr(alpha), ptr(np.ascontiguousarray(params)), C.c_double(tol),
                                       C.c_int(int(max_esteps))))

    def em_enqueue(self, final_pass=False):
        """Queues one iteration (E-step + device M-step, or the final inference pass) without waiting."""
        self._check(lib().hfg_em_enqueue(self._h, C.c_int(1 if final_pass else 0)))

    def em_finish(self, want_labels=True):
        """Waits for the queued iterations: returns (params, logliks of the E-steps that ran, converged, labels)."""
        params = np.zeros(self.n_regions, dtype=_abi.region_params_dtype)
        logliks = np.zeros(self._em_max, np.float64)
        n, conv = C.c_int(0), C.c_int(0)
        labels = np.empty(self.n_windows, np.int8) if want_labels else None
        self._check(lib().hfg_em_finish(self._h, ptr(params), ptr(logliks), C.byref(n), C.byref(conv), ptr(labels)))
        return params, logliks[:n.value].copy(), bool(conv.value), labels

    def em_enqueued_ms(self, i):
        f = lib().hfg_em_enqueued_ms
        f.restype = C.c_double
        return float(f(self._h, C.c_int(int(i))))

    def blocking_steps(self, alpha, params, n_steps, stats=None, labels=None, flush_bytes=0, tol=1e-12):
        """n_steps x [hfg_em_iteration + hfg_mstep] driven from C, as the drop-in binding does: returns (params, stats,
        logliks, labels, seconds per step)."""
        alpha = np.ascontiguousarray(alpha, np.float64)
        params = params.copy()
        if stats is None:
            stats = np.zeros(self.n_regions, dtype=_abi.region_stats_dtype)
        if labels is None:
            labels = np.empty(self.n_windows, np.int8)
        secs, ll = np.zeros(n_steps, np.float64), np.zeros(n_steps, np.float64)
        self._check(lib().hfg_debug_blocking_steps(self._h, ptr(alpha), ptr(params), ptr(stats), ptr(labels), C.c_int(int(n_steps)),
                                                   C.c_size_t(int(flush_bytes)), C.c_double(tol), ptr(secs), ptr(ll)))
        return params, stats, ll, labels, secs

    def l2_flush(self, nbytes=256 << 20):
        self._check(lib().hfg_debug_l2_flush(self._h, C.c_size_t(int(nbytes))))

    def run_em_accelerated(self, alpha, params, max_iterations, tol=1e-3, want_labels=True):
        """The EM loop of runHMMFlagger with --accelerate (SQUAREM): returns (params, logliks, alpha_rates, labels)."""
        alpha = np.ascontiguousarray(alpha, np.float64)
        params = params.copy()
        logliks = np.zeros(max_iterations + 1, np.float64)
        rates = np.zeros(max_iterations + 1, np.float64)
        n = C.c_int(0)
        labels = np.empty(self.n_windows, np.int8) if want_labels else None
        self._check(lib().hfg_run_em_accelerated(self._h, ptr(alpha), ptr(params), C.c_int(max_iterations),
                                                 C.c_double(tol), ptr(logliks), ptr(rates), C.byref(n), ptr(labels)))
        return params, logliks[:n.value + 1].copy(), rates[:n.value].copy(), labels

    def debug_exp(self, q):
        q = np.ascontiguousarray(q, np.float64)
        out = np.empty_like(q)
        self._check(lib().hfg_debug_exp(self._h, ptr(q), ptr(out), C.c_int(q.size)))
        return out

    def debug_phase_clocks(self):
        buf = np.zeros((4096, 16), np.int64)
        g = C.c_int(0)
        self._check(lib().hfg_debug_phase_clocks(self._h, ptr(buf), C.byref(g)))
        return buf[: g.value + 1].copy()  # last row: the tail clocks of the quad kernel

    def kernel_launches(self):
        return int(lib().hfg_kernel_launches(self._h))

    def last_call_device_ms(self):
        return float(lib().hfg_last_call_device_ms(self._h))

    def last_estep_kernel_ms(self):
        return float(lib().hfg_last_estep_kernel_ms(self._h))
