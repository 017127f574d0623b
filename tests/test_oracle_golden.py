"""CPU: the oracle (oracle/hmm_oracle.c) against the golden vectors generated from the UNMODIFIED reference, and --
when oracle/_ref has been built in this container -- against the reference itself, bit for bit."""
import numpy as np
import pytest

import golden_util
from flagger_b200 import _abi, synth


def _flat(a):
    return a.view(np.float64).reshape(-1)


@pytest.mark.parametrize("name", golden_util.NAMES)
def test_oracle_reproduces_golden_bit_exact(orc, name):
    g, wl = golden_util.load(name)
    cfg, alpha = g["cfg"], g["alpha"]
    assert orc.best_num_collapsed_comps(int(wl.cov.max()), wl.region_coverages) == int(g["K"])
    p0 = orc.model_init(cfg, wl.region_coverages, wl.window_len)
    assert np.array_equal(_flat(p0), _flat(g["params0"]))
    e = orc.estep(cfg, wl, alpha, p0, want_fb=True)
    assert e["rc"] == 0
    assert e["loglik"] == float(g["loglik"])
    for key in ("chunk_logliks", "labels", "posteriors", "fwd", "bwd", "scales"):
        assert np.array_equal(e[key], g[key]), key
    assert np.array_equal(_flat(e["stats"]), _flat(g["stats"]))
    p1, conv = orc.mstep(cfg, p0, e["stats"], tol=1e-3)
    assert np.array_equal(_flat(p1), _flat(g["params1"])) and conv == bool(g["converged1"])
    em = orc.run_em(cfg, wl, alpha, p0, 5, tol=1e-12)
    assert np.array_equal(em["logliks"], g["em_logliks"])
    assert np.array_equal(em["labels"], g["em_labels"])
    assert np.array_equal(_flat(em["params"]), _flat(g["em_params"]))
    f = orc.estep(cfg, wl, alpha, p1, forward_only=True)
    assert f["loglik"] == float(g["fwd_only_loglik"])


def test_oracle_matches_live_reference_when_built(orc, ref):
    if ref is None:
        pytest.skip("oracle/_ref not built (the reference tree is not mounted on this box)")
    for seed, R in ((31, 1), (32, 4), (33, 7)):
        wl = synth.small_mixed(n_regions=R, seed=seed)
        K = orc.best_num_collapsed_comps(int(wl.cov.max()), wl.region_coverages)
        assert K == ref.best_num_collapsed_comps(int(wl.cov.max()), wl.region_coverages)
        cfg = _abi.make_config(n_regions=R, n_col_comps=K)
        p = orc.model_init(cfg, wl.region_coverages, wl.window_len)
        assert np.array_equal(_flat(p), _flat(ref.model_init(cfg, wl.region_coverages, wl.window_len)))
        a, b = orc.estep(cfg, wl, synth.HIFI_ALPHA, p), ref.estep(cfg, wl, synth.HIFI_ALPHA, p)
        assert a["loglik"] == b["loglik"] and np.array_equal(a["labels"], b["labels"])
        assert np.array_equal(a["posteriors"], b["posteriors"])
        assert np.array_equal(_flat(a["stats"]), _flat(b["stats"]))
        ea, eb = orc.run_em(cfg, wl, synth.HIFI_ALPHA, p, 8), ref.run_em(cfg, wl, synth.HIFI_ALPHA, p, 8)
        assert np.array_equal(ea["logliks"], eb["logliks"]) and np.array_equal(ea["labels"], eb["labels"])
        assert np.array_equal(_flat(ea["params"]), _flat(eb["params"]))


def test_oracle_edge_cases(orc):
    """Ragged inputs: 1-, 2-, 3-window chunks (no pair statistics below 3 windows), region change at every window."""
    for L in (1, 2, 3, 4):
        wl = synth.make_workload([4000 * L], name=f"tiny{L}", seed=L)
        cfg = _abi.make_config(n_col_comps=3)
        p = orc.model_init(cfg, wl.region_coverages, wl.window_len)
        e = orc.estep(cfg, wl, synth.HIFI_ALPHA, p)
        assert e["rc"] == 0 and np.isfinite(e["loglik"]) and e["labels"].min() >= 0
        tot = _flat(e["stats"]).sum()
        assert (tot == 0.0) == (L < 3)  # pairs i -> i+1 exist only for i = 1 .. L-2
        assert np.allclose(e["posteriors"].sum(axis=1), 1.0)
    wl = synth.small_mixed(n_regions=3, seed=9)
    wl.region[:] = (np.arange(wl.n_windows) % 3).astype(np.uint8)  # transition is the constant 1/5 everywhere
    cfg = _abi.make_config(n_regions=3, n_col_comps=3)
    p = orc.model_init(cfg, wl.region_coverages, wl.window_len)
    e = orc.estep(cfg, wl, synth.HIFI_ALPHA, p)
    assert e["rc"] == 0 and np.isfinite(e["loglik"])
