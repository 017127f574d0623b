"""--accelerate (SQUAREM, hmm.c:820-1098; src/hmm_flagger.c:382-416).

CPU: the host arithmetic of libhfg (hfg_squarem_*, hfg_params_feasible) and the oracle's restatement against golden
vectors produced by the unmodified reference's SquareAccelerator (tests/golden/make_golden_squarem.py), bit for bit, and
the oracle's accelerated EM loop against the reference's.  GPU: hfg_run_em_accelerated against the same golden runs."""
import numpy as np
import pytest

import golden_util
from flagger_b200 import _abi, api, synth

SYN = ("09", "05", "01")


def _flat(a):
    return np.ascontiguousarray(a).view(np.float64).reshape(-1)


def _same(a, b):
    return np.array_equal(_flat(a), _flat(b), equal_nan=True)


@pytest.mark.parametrize("name", golden_util.NAMES)
def test_host_squarem_matches_reference_golden(orc, name):
    g, _ = golden_util.load(name)
    q = golden_util.load_squarem(name)
    cfg, p0, p1 = g["cfg"], g["params0"], q["p1"]
    triples = [(q["p2"], q["cand_params"], q["cand_rates"], q["cand_feasible"])]
    triples += [(q[f"syn{t}_p2"], q[f"syn{t}_params"], q[f"syn{t}_rates"], q[f"syn{t}_feasible"]) for t in SYN]
    for p2, want_params, want_rates, want_feasible in triples:
        for n in range(len(want_rates)):
            for impl in (api.squarem, orc.squarem):
                prime, rate, feasible = impl(cfg, p0, p1, p2, n)
                assert rate == want_rates[n], (impl, n)
                assert feasible == bool(want_feasible[n]), (impl, n)
                assert _same(prime, want_params[n]), (impl, n)
    assert api.params_feasible(cfg, p0) and api.params_feasible(cfg, p1)


def test_feasibility_rules():
    """hmm_utils.c:685-694,920-925,2130-2139: positive means/variances/rate/truncation point, weights and transitions in
    [0, 1]; a NaN fails an emission check but passes the transition check (the reference's comparisons, as written)."""
    cfg = _abi.make_config(n_col_comps=3)
    base = api.model_init(cfg, np.array([40], np.int32), 4000)
    assert api.params_feasible(cfg, base)
    for field, idx, value, ok in (("lambda", (), 0.0, False), ("trunc_point", (), -1.0, False),
                                  ("mean", (3, 1), 0.0, False), ("var", (2, 0), -1e-9, False),
                                  ("weight", (3, 2), 1.0000001, False), ("weight", (3, 2), -1e-12, False),
                                  ("weight", (3, 2), 0.0, True), ("trans", (1, 2), 1.0000001, False),
                                  ("trans", (0, 3), -1e-300, False), ("trans", (4, 0), 7.0, True),   # start row: unchecked
                                  ("trans", (0, 4), 7.0, True),                                      # end column: unchecked
                                  ("mean", (0, 0), -5.0, True),                                      # Err row unused here
                                  ("var", (1, 0), np.nan, False), ("trans", (2, 2), np.nan, True)):
        p = base.copy()
        if idx:
            p[field][0][idx] = value
        else:
            p[field][0] = value
        assert api.params_feasible(cfg, p) == ok, (field, idx, value)
    gauss = _abi.make_config(n_col_comps=3, model_type=_abi.MODEL_GAUSSIAN)
    p = api.model_init(gauss, np.array([40], np.int32), 4000)
    p["mean"][0][0, 0] = -5.0   # Err is a Gaussian in this model
    assert not api.params_feasible(gauss, p)


def test_squarem_against_live_reference(orc, ref):
    """Random slowly-contracting triples around real EM parameter sets, several regions, both model types."""
    if ref is None:
        pytest.skip("oracle/_ref not built (the reference tree is not mounted on this box)")
    rng = np.random.default_rng(7)
    for R, model_type in ((1, _abi.MODEL_TRUNC_EXP_GAUSSIAN), (4, _abi.MODEL_TRUNC_EXP_GAUSSIAN), (2, _abi.MODEL_GAUSSIAN)):
        wl = synth.small_mixed(n_regions=R, seed=40 + R)
        K = orc.best_num_collapsed_comps(int(wl.cov.max()), wl.region_coverages)
        cfg = _abi.make_config(n_regions=R, n_col_comps=K, model_type=model_type)
        p0 = orc.model_init(cfg, wl.region_coverages, wl.window_len)
        for _ in range(3):  # move away from the initial point
            p0, _c = orc.mstep(cfg, p0, orc.estep(cfg, wl, synth.HIFI_ALPHA, p0)["stats"])
        p1, _c = orc.mstep(cfg, p0, orc.estep(cfg, wl, synth.HIFI_ALPHA, p0)["stats"])
        p2, _c = orc.mstep(cfg, p1, orc.estep(cfg, wl, synth.HIFI_ALPHA, p1)["stats"])
        cases = [p2]
        for c in rng.uniform(0.0, 0.95, 6):
            q2 = p1.copy()
            q2.view(np.float64)[:] = p1.view(np.float64) + c * (p1.view(np.float64) - p0.view(np.float64))
            cases.append(q2)
        for q2 in cases:
            for n in (0, 1, 2, 5, 9):
                want = ref.squarem(cfg, p0, p1, q2, n)
                for impl in (api.squarem, orc.squarem):
                    got = impl(cfg, p0, p1, q2, n)
                    assert got[1] == want[1] and got[2] == want[2] and _same(got[0], want[0]), (impl, R, n)


@pytest.mark.parametrize("name", golden_util.NAMES)
def test_oracle_accelerated_em_matches_reference_golden(orc, name):
    g, wl = golden_util.load(name)
    q = golden_util.load_squarem(name)
    acc = orc.run_em_accelerated(g["cfg"], wl, g["alpha"], g["params0"], len(q["acc_rates"]), tol=1e-12)
    assert acc["rc"] == 0
    assert np.array_equal(acc["logliks"], q["acc_logliks"])
    assert np.array_equal(acc["alpha_rates"], q["acc_rates"])
    assert np.array_equal(acc["labels"], q["acc_labels"])
    assert _same(acc["params"], q["acc_params"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", golden_util.NAMES)
def test_gpu_accelerated_em_matches_reference_golden(name):
    g, wl = golden_util.load(name)
    q = golden_util.load_squarem(name)
    gpu = api.HmmFlaggerGPU(g["cfg"], wl)
    params, logliks, rates, labels = gpu.run_em_accelerated(g["alpha"], g["params0"], len(q["acc_rates"]), tol=1e-12)
    gpu.close()
    assert logliks.shape == q["acc_logliks"].shape
    assert np.allclose(logliks, q["acc_logliks"], rtol=1e-9, atol=0)   # bar: 1e-5 relative
    # the step length is sqrt(sum r^2 / sum v^2) with v a SECOND difference of successive parameter sets: near convergence
    # it amplifies the 1e-15-relative rounding differences between the two E-steps by |p| / |v| (observed: 9e-5 on the
    # third outer iteration of one fixture), while the parameters themselves barely move with it
    assert np.allclose(rates, q["acc_rates"], rtol=2e-3, atol=0)
    assert np.array_equal(labels, q["acc_labels"])                      # bar: bit-exact
    assert np.allclose(_flat(params), _flat(q["acc_params"]), rtol=1e-5, atol=1e-12)


@pytest.mark.gpu
def test_gpu_accelerated_em_fullsize_against_oracle(orc):
    """A run where the step length is actually < -1 most iterations: the oracle's loop on a mid-size workload."""
    wl = synth.small_mixed(n_regions=2, seed=77)
    K = api.best_num_collapsed_comps(int(wl.cov.max()), wl.region_coverages)
    cfg = _abi.make_config(n_regions=2, n_col_comps=K)
    p0 = api.model_init(cfg, wl.region_coverages, wl.window_len)
    want = orc.run_em_accelerated(cfg, wl, synth.HIFI_ALPHA, p0, 6, tol=1e-12)
    gpu = api.HmmFlaggerGPU(cfg, wl)
    params, logliks, rates, labels = gpu.run_em_accelerated(synth.HIFI_ALPHA, p0, 6, tol=1e-12)
    gpu.close()
    assert np.allclose(logliks, want["logliks"], rtol=1e-9, atol=0)
    assert np.allclose(rates, want["alpha_rates"], rtol=2e-3, atol=0)
    assert np.array_equal(labels, want["labels"])
    assert np.allclose(_flat(params), _flat(want["params"]), rtol=1e-5, atol=1e-12)
