#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ from the UNMODIFIED reference (oracle/_ref/libref_harness.so,
built by oracle/Makefile from /root/reference).  Run in the build container, where the reference tree is mounted:

    python tests/golden/make_golden.py

Each fixture is a small .npz holding the exact inputs (window arrays, chunk table, configuration, alpha, parameters)
and the reference's outputs at full double precision: one E-step (statistics, log-likelihoods, labels, posteriors,
forward/backward/scales), the M-step that follows, and a 5-iteration EM run (log-likelihood trajectory, final
parameters, final labels).  The reference ships no such vectors itself."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from flagger_b200 import _abi, synth  # noqa: E402
import oracle_lib  # noqa: E402

CASES = {
    # name: (workload factory, n_regions, model_type, adjust_ends, alpha)
    "mixed_r3_hifi": (lambda: synth.small_mixed(n_regions=3, seed=101), 3, _abi.MODEL_TRUNC_EXP_GAUSSIAN, True, "hifi"),
    "mixed_r1_zero_noadjust": (lambda: synth.small_mixed(n_regions=1, seed=102), 1, _abi.MODEL_TRUNC_EXP_GAUSSIAN, False,
                               "zero"),
    "cfg1_250w": (lambda: synth.config1(seed=103), 1, _abi.MODEL_TRUNC_EXP_GAUSSIAN, True, "hifi"),
    "mixed_r1_gaussian_model": (lambda: synth.small_mixed(n_regions=1, seed=104), 1, _abi.MODEL_GAUSSIAN, True, "hifi"),
}


def main():
    ref = oracle_lib.reference(threads=2)
    if ref is None:
        raise SystemExit("oracle/_ref/libref_harness.so is missing: run `make -C oracle` with /root/reference mounted")
    for name, (factory, R, model_type, adjust, alpha_name) in CASES.items():
        wl = factory()
        K = ref.best_num_collapsed_comps(int(wl.cov.max()), wl.region_coverages)
        cfg = _abi.make_config(n_regions=R, n_col_comps=K, model_type=model_type, adjust_contig_ends=adjust,
                               mean_read_length=wl.avg_alignment_len)
        alpha = synth.HIFI_ALPHA if alpha_name == "hifi" else np.zeros((4, 4))
        p0 = ref.model_init(cfg, wl.region_coverages, wl.window_len)
        e = ref.estep(cfg, wl, alpha, p0, want_fb=True)
        p1, conv1 = ref.mstep(cfg, p0, e["stats"], tol=1e-3)
        em = ref.run_em(cfg, wl, alpha, p0, 5, tol=1e-12)
        fwd = ref.estep(cfg, wl, alpha, p1, forward_only=True)
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"),
            cfg=cfg, chunks=wl.chunks, cov=wl.cov, cov_high_mapq=wl.cov_high_mapq, cov_high_clip=wl.cov_high_clip,
            region=wl.region, region_coverages=wl.region_coverages, window_len=wl.window_len, alpha=alpha, K=K,
            params0=p0, stats=e["stats"], loglik=e["loglik"], chunk_logliks=e["chunk_logliks"], labels=e["labels"],
            posteriors=e["posteriors"], fwd=e["fwd"], bwd=e["bwd"], scales=e["scales"], params1=p1, converged1=conv1,
            em_logliks=em["logliks"], em_params=em["params"], em_labels=em["labels"], fwd_only_loglik=fwd["loglik"])
        print(name, wl.n_windows, "windows", os.path.getsize(os.path.join(HERE, name + ".npz")), "bytes")


if __name__ == "__main__":
    main()
