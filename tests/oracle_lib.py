"""ctypes access to the CHECKERS: oracle/liborc.so (C restatement) and oracle/_ref/libref_harness.so
(unmodified reference objects).  Test infrastructure only -- never imported by flagger_b200/."""
import ctypes as C
import os
import subprocess

import numpy as np

from flagger_b200 import _abi
from flagger_b200._abi import ptr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_BIN = os.path.join(ORACLE_DIR, "_ref", "hmm_flagger_ref")


def build_oracle():
    """Compile the restatement (and the reference build when /root/reference is mounted)."""
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True, stdout=subprocess.DEVNULL)


def _load(path):
    return C.CDLL(path) if os.path.exists(path) else None


class _Checker:
    """Common numpy-facing API over orc_* / ref_* entry points."""

    def __init__(self, lib, prefix, threads=None):
        self.lib = lib
        self.prefix = prefix
        self.threads = threads

    def _fn(self, name):
        f = getattr(self.lib, f"{self.prefix}_{name}")
        f.restype = C.c_int
        return f

    def estep(self, cfg, wl, alpha, params, forward_only=False, want_fb=False):
        W, Cn, R = wl.n_windows, wl.n_chunks, int(cfg["n_regions"][0])
        stats = np.zeros(R, dtype=_abi.region_stats_dtype)
        loglik = C.c_double(0.0)
        chunk_ll = np.zeros(Cn, np.float64)
        labels = np.full(W, -1, np.int8)
        post = np.zeros((W, 4), np.float64)
        fwd = np.zeros((W, 4), np.float64) if want_fb else None
        bwd = np.zeros((W, 4), np.float64) if want_fb else None
        scales = np.zeros(W, np.float64) if want_fb else None
        alpha = np.ascontiguousarray(alpha, np.float64)
        args = [ptr(cfg), C.c_int(Cn), ptr(wl.chunks), ptr(wl.cov), ptr(wl.cov_high_mapq), ptr(wl.cov_high_clip),
                ptr(wl.region), ptr(alpha), ptr(params), ptr(stats), C.byref(loglik), ptr(chunk_ll), ptr(labels),
                ptr(post), ptr(fwd), ptr(bwd), ptr(scales), C.c_int(1 if forward_only else 0)]
        if self.threads is not None:
            args.append(C.c_int(self.threads))
        rc = self._fn("estep")(*args)
        return dict(rc=rc, stats=stats, loglik=loglik.value, chunk_logliks=chunk_ll, labels=labels, posteriors=post,
                    fwd=fwd, bwd=bwd, scales=scales)

    def model_init(self, cfg, region_coverages, window_len, start_only=False):
        params = np.zeros(int(cfg["n_regions"][0]), dtype=_abi.region_params_dtype)
        rc_ = np.ascontiguousarray(region_coverages, np.int32)
        rc = self._fn("model_init")(ptr(cfg), ptr(rc_), C.c_int(window_len), C.c_int(1 if start_only else 0),
                                    ptr(params))
        assert rc == 0
        return params

    def mstep(self, cfg, params, stats, tol=1e-3):
        params = params.copy()
        conv = C.c_int(0)
        rc = self._fn("mstep")(ptr(cfg), ptr(params), ptr(stats), C.c_double(tol), C.byref(conv))
        assert rc == 0
        return params, bool(conv.value)

    def run_em(self, cfg, wl, alpha, params, max_iterations, tol=1e-12, threads=None):
        params = params.copy()
        logliks = np.zeros(max_iterations + 1, np.float64)
        n = C.c_int(0)
        labels = np.full(wl.n_windows, -1, np.int8)
        alpha = np.ascontiguousarray(alpha, np.float64)
        args = [ptr(cfg), C.c_int(wl.n_chunks), ptr(wl.chunks), ptr(wl.cov), ptr(wl.cov_high_mapq),
                ptr(wl.cov_high_clip), ptr(wl.region), ptr(alpha), ptr(params), C.c_int(max_iterations),
                C.c_double(tol), ptr(logliks), C.byref(n), ptr(labels)]
        secs = C.c_double(0.0)
        if self.threads is not None:
            args += [C.c_int(threads or self.threads), C.byref(secs)]
        rc = self._fn("run_em")(*args)
        return dict(rc=rc, params=params, logliks=logliks[:n.value].copy(), labels=labels, estep_seconds=secs.value)

    def squarem(self, cfg, p0, p1, p2, n_shrinks=0, margin=1e-2):
        """SquareAccelerator: computeRates + computeValuesForModelPrime + n_shrinks step halvings."""
        p0, p1, p2 = (np.ascontiguousarray(p) for p in (p0, p1, p2))
        prime = np.zeros_like(p0)
        rate = C.c_double(0.0)
        feasible = self._fn("squarem")(ptr(cfg), ptr(p0), ptr(p1), ptr(p2), C.c_int(n_shrinks), C.c_double(margin),
                                       ptr(prime), C.byref(rate))
        return prime, rate.value, bool(feasible)

    def run_em_accelerated(self, cfg, wl, alpha, params, max_iterations, tol=1e-12):
        params = params.copy()
        logliks = np.zeros(max_iterations + 1, np.float64)
        rates = np.zeros(max_iterations + 1, np.float64)
        n = C.c_int(0)
        labels = np.full(wl.n_windows, -1, np.int8)
        alpha = np.ascontiguousarray(alpha, np.float64)
        args = [ptr(cfg), C.c_int(wl.n_chunks), ptr(wl.chunks), ptr(wl.cov), ptr(wl.cov_high_mapq),
                ptr(wl.cov_high_clip), ptr(wl.region), ptr(alpha), ptr(params), C.c_int(max_iterations),
                C.c_double(tol), ptr(logliks), ptr(rates), C.byref(n), ptr(labels)]
        if self.threads is not None:
            args.append(C.c_int(self.threads))
        rc = self._fn("run_em_accelerated")(*args)
        return dict(rc=rc, params=params, logliks=logliks[:n.value + 1].copy(), alpha_rates=rates[:n.value].copy(),
                    labels=labels)

    def nb_histogram(self, cfg, wl, alpha, params):
        """(E-step result, [R, 4, 250] histogram of the pair mass over the coverage value summed over chunks): the
        negative-binomial model's CountData (oracle only)."""
        assert self.prefix == "orc"
        hist = np.zeros((int(cfg["n_regions"][0]), 4, 250))
        self.lib.orc_nb_histogram_out.restype = None
        self.lib.orc_nb_histogram_out(ptr(hist))
        try:
            e = self.estep(cfg, wl, alpha, params)
        finally:
            self.lib.orc_nb_histogram_out(None)
        return e, hist

    def digamma(self, x):
        """digamma of a double argument in long double, returned as (hi, lo) doubles with hi + lo exact -- the routine
        behind the negative-binomial model's table (orc_digammal / the reference's digammal)."""
        f = getattr(self.lib, "orc_digammal" if self.prefix == "orc" else "digammal")
        f.restype, f.argtypes = C.c_longdouble, [C.c_longdouble]
        v = np.longdouble(f(np.longdouble(x)))
        hi = np.float64(v)
        return float(hi), float(v - np.longdouble(hi))

    def best_num_collapsed_comps(self, max_cov, region_coverages):
        rc_ = np.ascontiguousarray(region_coverages, np.int32)
        return self._fn("best_num_collapsed_comps")(C.c_int(int(max_cov)), ptr(rc_), C.c_int(len(rc_)))


def oracle():
    """The C restatement (always available once built)."""
    path = os.path.join(ORACLE_DIR, "liborc.so")
    if not os.path.exists(path):
        build_oracle()
    return _Checker(C.CDLL(path), "orc")


def reference(threads=1):
    """The unmodified reference behind ref_harness.c, or None when oracle/_ref has not been built."""
    lib = _load(os.path.join(ORACLE_DIR, "_ref", "libref_harness.so"))
    return _Checker(lib, "ref", threads=threads) if lib is not None else None


class CpuRun:
    """Persistent CPU run (reference objects when available, else the restatement) for timing whole EM iterations.
    Used only by bench.py's cpu_baseline / `--impl reference` legs."""

    def __init__(self, cfg, wl, alpha, params, threads):
        ref_lib = _load(os.path.join(ORACLE_DIR, "_ref", "libref_harness.so"))
        if ref_lib is not None:
            self.lib, self.prefix, self.kind, self.cores = ref_lib, "ref", "reference", int(threads)
        else:
            self.lib, self.prefix, self.kind, self.cores = oracle().lib, "orc", "port", 1
        self.wl = wl  # keep the arrays alive
        self.cfg = np.ascontiguousarray(cfg)
        self.alpha = np.ascontiguousarray(alpha, np.float64)
        self.params = np.ascontiguousarray(params)
        fn = getattr(self.lib, f"{self.prefix}_open")
        fn.restype = C.c_void_p
        self.h = C.c_void_p(fn(ptr(self.cfg), C.c_int(wl.n_chunks), ptr(wl.chunks), ptr(wl.cov), ptr(wl.cov_high_mapq),
                               ptr(wl.cov_high_clip), ptr(wl.region), ptr(self.alpha), ptr(self.params)))
        self._step = getattr(self.lib, f"{self.prefix}_step")
        self._step.restype = C.c_double

    def step(self, tol=1e-12):
        ll = C.c_double(0.0)
        secs = self._step(self.h, C.c_int(self.cores), C.c_double(tol), C.byref(ll))
        return float(secs), ll.value

    def close(self):
        if self.h:
            getattr(self.lib, f"{self.prefix}_close")(self.h)
            self.h = None


def reference_parse_cov(path, chunk_len, window_len, threads=2):
    """The reference's own chunk builder on a .cov/.cov.gz (writes <path>.index next to it!).  None when oracle/_ref is
    not built.  Returns dict(chunks, names, cov, mapq, clip, flags, truth, prediction, region_coverages, header)."""
    lib = _load(os.path.join(ORACLE_DIR, "_ref", "libref_harness.so"))
    if lib is None:
        return None
    lib.ref_cov_open.restype = C.c_void_p
    h = C.c_void_p(lib.ref_cov_open(str(path).encode(), C.c_int(chunk_len), C.c_int(window_len), C.c_int(threads)))
    n_chunks, n_windows = C.c_int32(0), C.c_int64(0)
    header = np.zeros(8, np.int32)
    lib.ref_cov_counts(h, C.byref(n_chunks), C.byref(n_windows), ptr(header))
    Cn, W = n_chunks.value, n_windows.value
    out = dict(chunks=np.zeros(Cn, _abi.chunk_desc_dtype), names=np.zeros(Cn * 200, np.uint8), cov=np.zeros(W, np.uint16),
               mapq=np.zeros(W, np.uint16), clip=np.zeros(W, np.uint16), flags=np.zeros(W, np.uint64),
               truth=np.zeros(W, np.int8), prediction=np.zeros(W, np.int8),
               region_coverages=np.zeros(max(int(header[1]), 1), np.int32))
    lib.ref_cov_fill(h, ptr(out["chunks"]), ptr(out["names"]), ptr(out["cov"]), ptr(out["mapq"]), ptr(out["clip"]),
                     ptr(out["flags"]), ptr(out["truth"]), ptr(out["prediction"]), ptr(out["region_coverages"]))
    lib.ref_cov_close(h)
    raw = out["names"].tobytes()
    out["names"] = [raw[i * 200:(i + 1) * 200].split(b"\0", 1)[0].decode() for i in range(Cn)]
    out["region_coverages"] = out["region_coverages"][: int(header[1])]
    out["header"] = header
    return out
