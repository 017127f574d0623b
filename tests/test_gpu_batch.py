"""GPU: batched EM runs (hfg_batch_*, include/hfg.h) -- B (alpha, start parameters) candidates over the same windows, several
at a time on sub-grids of one GPU -- against the same candidates run one after the other on the whole GPU (hfg_run_em).

Bars: labels identical; log-likelihoods and fitted parameters equal to rounding (a smaller grid cuts the chains into other
segments, the sums associate differently); the same number of E-steps (same stopping rule)."""
import numpy as np
import pytest

from flagger_b200 import _abi, api, synth

pytestmark = pytest.mark.gpu


def candidates(n, seed=0):
    rng = np.random.default_rng(seed)
    out = [synth.HIFI_ALPHA.copy()]
    for _ in range(n - 1):
        a = synth.HIFI_ALPHA * rng.uniform(0.3, 1.6, (4, 4))
        out.append(np.clip(a, 0.0, 0.95))
    return np.array(out)


@pytest.mark.parametrize("kind,n_lanes,n_runs,iters", [("small3", 4, 6, 8), ("medium", 8, 8, 12), ("medium", 3, 7, 5)])
def test_batch_equals_sequential_runs(kind, n_lanes, n_runs, iters):
    wl = synth.small_mixed(n_regions=3, seed=14) if kind == "small3" else synth.config2(total_bp=300_000_000, seed=22)
    R = len(wl.region_coverages)
    K = api.best_num_collapsed_comps(int(wl.cov.max()), wl.region_coverages)
    cfg = _abi.make_config(n_regions=R, n_col_comps=K)
    p0 = api.model_init(cfg, wl.region_coverages, wl.window_len)
    alphas = candidates(n_runs, seed=n_lanes)
    one = api.HmmFlaggerGPU(cfg, wl)
    want = [one.run_em(a, p0, iters, tol=1e-3) for a in alphas]
    one.close()
    batch = api.HmmFlaggerBatch(cfg, wl, n_lanes=n_lanes)
    try:
        params, logliks, labels = batch.run_em(alphas, p0, iters, tol=1e-3)
        # a second call on the same batch (the tuner calls it once per round of proposals)
        params2, logliks2, labels2 = batch.run_em(alphas[::-1], p0, iters, tol=1e-3)
    finally:
        batch.close()
    for r, (wp, wll, wlab) in enumerate(want):
        assert len(logliks[r]) == len(wll), (r, len(logliks[r]), len(wll))
        assert np.all(np.abs(logliks[r] - wll) <= 1e-10 * np.abs(wll))
        assert np.array_equal(labels[r], wlab), (r, int((labels[r] != wlab).sum()))
        a, b = _abi.params_as_flat(params[r]), _abi.params_as_flat(wp)
        nz = np.abs(b) > 0
        assert np.all(np.abs(a[nz] - b[nz]) <= 1e-8 * np.abs(b[nz]))
        assert np.array_equal(labels2[n_runs - 1 - r], wlab)
        assert np.array_equal(logliks2[n_runs - 1 - r], logliks[r])  # same lane arithmetic whichever lane runs it


def test_batch_rejects_what_it_cannot_serve():
    wl = synth.small_mixed(n_regions=1, seed=12)
    K = api.best_num_collapsed_comps(int(wl.cov.max()), wl.region_coverages)
    cfg = _abi.make_config(n_regions=1, n_col_comps=K)
    with pytest.raises(api.HfgError):
        api.HmmFlaggerBatch(cfg, wl, n_lanes=0)
    with pytest.raises(api.HfgError):
        api.HmmFlaggerBatch(cfg, wl, n_lanes=17)
    nb = _abi.make_config(n_regions=1, n_col_comps=K, model_type=_abi.MODEL_NEGATIVE_BINOMIAL)
    with pytest.raises(api.HfgError):
        api.HmmFlaggerBatch(nb, wl, n_lanes=2)


def test_tuner_engine_batches_its_candidates(tmp_path):
    from flagger_b200 import binfmt, tune_alpha
    inp = str(tmp_path / "train.cov.gz")
    binfmt.write_random_rle_cov(inp, [310_000, 650_000, 2_400_000], seed=23, n_regions=1, with_truth=True)
    cov = binfmt.NativeCov(inp, 1_000_000, 4000)
    eng = tune_alpha.GpuEngine(cov, "trunc_exp_gaussian", 6, 1e-3, lanes=4)
    try:
        alphas = candidates(5, seed=3)
        got = eng.many(alphas)
        for a, lab in zip(alphas, got):
            assert np.array_equal(lab, eng(a))
        obj = tune_alpha.Objective([(cov, eng)])
        xs = [tune_alpha.alpha_to_x(a) for a in alphas]
        many = obj.score_many(xs)
        assert many == [obj.score(x) for x in xs]
    finally:
        eng.close()
        cov.close()
