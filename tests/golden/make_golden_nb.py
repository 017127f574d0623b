#!/usr/bin/env python
"""Golden vectors for the negative-binomial model (`--modelType negative_binomial`) from the UNMODIFIED reference
(oracle/_ref/libref_harness.so):

    python tests/golden/make_golden_nb.py

The model is restated by the oracle only (oracle/hmm_oracle.h, ORC_MODEL_NEGATIVE_BINOMIAL); these fixtures pin that
restatement.  Each <name>.nb.npz holds the inputs and, at full double precision, one E-step (statistics = the theta /
lambda / weight estimators filled from the per-state count histograms, log-likelihoods, labels, posteriors, forward /
backward / scales), the M-step that follows, a 5-iteration EM run, SQUAREM candidates for three successive parameter sets
and a 2-outer-iteration accelerated run.  digamma.nb.npz holds the reference's long-double digamma at 600 arguments as
(hi, lo) double pairs."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from flagger_b200 import _abi, synth  # noqa: E402
import oracle_lib  # noqa: E402

NB = 2  # ORC_MODEL_NEGATIVE_BINOMIAL
CASES = {
    "mixed_r3": (lambda: synth.small_mixed(n_regions=3, seed=201), 3, True),
    "mixed_r1_noadjust": (lambda: synth.small_mixed(n_regions=1, seed=202), 1, False),
    "cfg1_250w": (lambda: synth.config1(seed=203), 1, True),
}


def main():
    ref = oracle_lib.reference(threads=2)
    if ref is None:
        raise SystemExit("oracle/_ref/libref_harness.so is missing: run `make -C oracle` with /root/reference mounted")
    rng = np.random.default_rng(7)
    xs = np.concatenate([rng.uniform(-6, 6, 200), rng.uniform(0, 2000, 380),
                         [1.0, 2.0, 3.0, 0.5, 1e-6, 2.9999999, 3.0000001, 1e4, 1.5, 2.5, 0.999999, 4.0, 6.0, 12.0, 96.0, 250.0,
                          -0.5, -1.5, -2.25, 1e-12]])
    dg = np.array([ref.digamma(x) for x in xs])
    np.savez_compressed(os.path.join(HERE, "digamma.nb.npz"), x=xs, hi=dg[:, 0], lo=dg[:, 1])
    for name, (factory, R, adjust) in CASES.items():
        wl = factory()
        K = ref.best_num_collapsed_comps(int(wl.cov.max()), wl.region_coverages)
        cfg = _abi.make_config(n_regions=R, n_col_comps=K, model_type=NB, adjust_contig_ends=adjust,
                               mean_read_length=wl.avg_alignment_len)
        alpha = np.zeros((4, 4))  # the reference's NB emission ignores alpha; its count histograms assume it is zero
        p0 = ref.model_init(cfg, wl.region_coverages, wl.window_len)
        e = ref.estep(cfg, wl, alpha, p0, want_fb=True)
        p1, conv1 = ref.mstep(cfg, p0, e["stats"], tol=1e-3)
        p2, _ = ref.mstep(cfg, p1, ref.estep(cfg, wl, alpha, p1)["stats"], tol=1e-3)
        em = ref.run_em(cfg, wl, alpha, p0, 5, tol=1e-12)
        fwd = ref.estep(cfg, wl, alpha, p1, forward_only=True)
        cands = [ref.squarem(cfg, p0, p1, p2, n) for n in range(4)]
        acc = ref.run_em_accelerated(cfg, wl, alpha, p0, 2, tol=1e-12)
        assert acc["rc"] == 0 and e["rc"] == 0
        out = os.path.join(HERE, name + ".nb.npz")
        np.savez_compressed(
            out, cfg=cfg, chunks=wl.chunks, cov=wl.cov, cov_high_mapq=wl.cov_high_mapq, cov_high_clip=wl.cov_high_clip,
            region=wl.region, region_coverages=wl.region_coverages, window_len=wl.window_len, alpha=alpha, K=K,
            params0=p0, stats=e["stats"], loglik=e["loglik"], chunk_logliks=e["chunk_logliks"], labels=e["labels"],
            posteriors=e["posteriors"], fwd=e["fwd"], bwd=e["bwd"], scales=e["scales"], params1=p1, converged1=conv1,
            params2=p2, em_logliks=em["logliks"], em_params=em["params"], em_labels=em["labels"],
            fwd_only_loglik=fwd["loglik"], cand_params=np.stack([c[0] for c in cands]),
            cand_rates=np.array([c[1] for c in cands]), cand_feasible=np.array([c[2] for c in cands]),
            acc_logliks=acc["logliks"], acc_rates=acc["alpha_rates"], acc_params=acc["params"], acc_labels=acc["labels"])
        print(name, wl.n_windows, "windows; EM logliks", em["logliks"], "labels", np.bincount(em["labels"], minlength=4),
              "acc rates", acc["alpha_rates"], os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
