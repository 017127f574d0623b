"""GPU parity: the CUDA E-step (through the C-ABI of include/hfg.h) against the CPU oracle on identical inputs.

Bars (BASELINE.json north_star): per-window labels bit-exact; forward log-likelihoods and posteriors within
1e-5 relative (we assert far tighter, see TOL_*); statistics and parameters after the M-step likewise.
"""
import numpy as np
import pytest

from flagger_b200 import _abi, api, synth

pytestmark = pytest.mark.gpu

TOL_LOGLIK = 1e-9   # relative; north_star allows 1e-5
TOL_POST = 1e-5     # relative, as north_star states
TOL_STATS = 1e-9    # relative to the largest statistic of the region


def _rel(a, b):
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-300)


def _check_estep(gpu, oracle_out, wl, alpha, params, label_exact=True):
    stats, ll, labels = gpu.em_iteration(alpha, params)
    assert oracle_out["rc"] == 0
    assert abs(ll - oracle_out["loglik"]) <= TOL_LOGLIK * abs(oracle_out["loglik"])
    cl = gpu.chunk_logliks()
    assert np.all(np.abs(cl - oracle_out["chunk_logliks"]) <= 1e-9 * np.abs(oracle_out["chunk_logliks"]) + 1e-9)
    mism = int((labels != oracle_out["labels"]).sum())
    if label_exact:
        assert mism == 0, f"{mism} label mismatches of {wl.n_windows}"
    post = gpu.posteriors()
    # relative agreement wherever the posterior is not vanishing; absolute below that
    big = oracle_out["posteriors"] > 1e-200
    assert np.all(_rel(post[big], oracle_out["posteriors"][big]) <= TOL_POST)
    assert np.all(post[~big] <= 1e-190)
    sg, so = _abi.stats_as_flat(stats), _abi.stats_as_flat(oracle_out["stats"])
    scale = np.abs(so).max()
    assert np.all(np.abs(sg - so) <= TOL_STATS * scale), np.abs(sg - so).max() / scale
    return stats, ll, labels, mism


@pytest.mark.parametrize("n_regions,adjust,alpha_name", [
    (1, True, "hifi"), (3, True, "hifi"), (3, False, "zero"), (1, True, "zero"), (7, True, "hifi"),
])
def test_estep_matches_oracle_small(orc, n_regions, adjust, alpha_name):
    wl = synth.small_mixed(n_regions=n_regions, seed=11 + n_regions)
    K = api.best_num_collapsed_comps(int(wl.cov.max()), wl.region_coverages)
    cfg = _abi.make_config(n_regions=n_regions, n_col_comps=K, adjust_contig_ends=adjust)
    alpha = synth.HIFI_ALPHA if alpha_name == "hifi" else np.zeros((4, 4))
    params = api.model_init(cfg, wl.region_coverages, wl.window_len)
    gpu = api.HmmFlaggerGPU(cfg, wl)
    out = orc.estep(cfg, wl, alpha, params)
    stats, ll, labels, _ = _check_estep(gpu, out, wl, alpha, params)
    # second iteration from the M-step of the first (parameters no longer symmetric)
    p2, _ = api.mstep(cfg, params, stats)
    p2o, _ = orc.mstep(cfg, params, out["stats"])
    assert np.allclose(_abi.params_as_flat(p2), _abi.params_as_flat(p2o), rtol=1e-9, atol=0)
    out2 = orc.estep(cfg, wl, alpha, p2o)
    _check_estep(gpu, out2, wl, alpha, p2o)
    gpu.close()


def test_forward_only_matches(orc):
    wl = synth.small_mixed(n_regions=3, seed=3)
    cfg = _abi.make_config(n_regions=3, n_col_comps=4)
    params = api.model_init(cfg, wl.region_coverages, wl.window_len)
    gpu = api.HmmFlaggerGPU(cfg, wl)
    ll = gpu.forward_only(synth.HIFI_ALPHA, params)
    out = orc.estep(cfg, wl, synth.HIFI_ALPHA, params, forward_only=True)
    assert abs(ll - out["loglik"]) <= TOL_LOGLIK * abs(out["loglik"])
    gpu.close()


def test_gaussian_model_type(orc):
    wl = synth.small_mixed(n_regions=1, seed=8)
    cfg = _abi.make_config(n_regions=1, n_col_comps=3, model_type=_abi.MODEL_GAUSSIAN)
    params = api.model_init(cfg, wl.region_coverages, wl.window_len)
    gpu = api.HmmFlaggerGPU(cfg, wl)
    out = orc.estep(cfg, wl, synth.HIFI_ALPHA, params)
    _check_estep(gpu, out, wl, synth.HIFI_ALPHA, params)
    gpu.close()


def test_config1_full_em_matches_oracle(orc):
    """BASELINE.json configs[0]: 1 Mbp contig, 250 windows, 5 EM iterations + final decode."""
    wl = synth.config1()
    K = api.best_num_collapsed_comps(int(wl.cov.max()), wl.region_coverages)
    cfg = _abi.make_config(n_regions=1, n_col_comps=K)
    params = api.model_init(cfg, wl.region_coverages, wl.window_len)
    gpu = api.HmmFlaggerGPU(cfg, wl)
    pg, llg, labg = gpu.run_em(synth.HIFI_ALPHA, params, 5, tol=1e-12)
    eo = orc.run_em(cfg, wl, synth.HIFI_ALPHA, params, 5, tol=1e-12)
    assert len(llg) == len(eo["logliks"]) == 6
    assert np.all(np.abs(llg - eo["logliks"]) <= TOL_LOGLIK * np.abs(eo["logliks"]))
    assert np.array_equal(labg, eo["labels"])
    assert np.allclose(_abi.params_as_flat(pg), _abi.params_as_flat(eo["params"]), rtol=1e-8, atol=0)
    gpu.close()


def test_medium_many_chunks_em(orc):
    """cfg3 shape at reduced size: 400 contigs x 75 windows, 3 EM iterations, labels bit-exact."""
    wl = synth.config3(n_contigs=400, seed=21)
    K = api.best_num_collapsed_comps(int(wl.cov.max()), wl.region_coverages)
    cfg = _abi.make_config(n_regions=1, n_col_comps=K)
    params = api.model_init(cfg, wl.region_coverages, wl.window_len)
    gpu = api.HmmFlaggerGPU(cfg, wl)
    pg, llg, labg = gpu.run_em(synth.HIFI_ALPHA, params, 3, tol=1e-12)
    eo = orc.run_em(cfg, wl, synth.HIFI_ALPHA, params, 3, tol=1e-12)
    assert np.all(np.abs(llg - eo["logliks"]) <= TOL_LOGLIK * np.abs(eo["logliks"]))
    assert np.array_equal(labg, eo["labels"])
    gpu.close()


def test_long_chunks_em(orc):
    """cfg2 shape at reduced size: chr-like contigs cut into 20 Mb chunks (up to 9999 windows per chunk)."""
    wl = synth.config2(total_bp=300_000_000, seed=22)
    K = api.best_num_collapsed_comps(int(wl.cov.max()), wl.region_coverages)
    cfg = _abi.make_config(n_regions=1, n_col_comps=K)
    params = api.model_init(cfg, wl.region_coverages, wl.window_len)
    gpu = api.HmmFlaggerGPU(cfg, wl)
    out = orc.estep(cfg, wl, synth.HIFI_ALPHA, params)
    stats, ll, labels, _ = _check_estep(gpu, out, wl, synth.HIFI_ALPHA, params)
    pg, llg, labg = gpu.run_em(synth.HIFI_ALPHA, params, 3, tol=1e-12)
    eo = orc.run_em(cfg, wl, synth.HIFI_ALPHA, params, 3, tol=1e-12)
    assert np.all(np.abs(llg - eo["logliks"]) <= TOL_LOGLIK * np.abs(eo["logliks"]))
    assert np.array_equal(labg, eo["labels"])
    gpu.close()


def test_determinism():
    wl = synth.small_mixed(n_regions=3, seed=5)
    cfg = _abi.make_config(n_regions=3, n_col_comps=4)
    params = api.model_init(cfg, wl.region_coverages, wl.window_len)
    gpu = api.HmmFlaggerGPU(cfg, wl)
    s1, l1, b1 = gpu.em_iteration(synth.HIFI_ALPHA, params)
    s2, l2, b2 = gpu.em_iteration(synth.HIFI_ALPHA, params)
    assert l1 == l2 and np.array_equal(b1, b2)
    assert np.array_equal(_abi.stats_as_flat(s1), _abi.stats_as_flat(s2))
    gpu.close()


def test_custom_exp_accuracy():
    """The kernel's exp_nonpos() against numpy (glibc) exp over the argument range the path produces."""
    cfg = _abi.make_config()
    gpu = api.HmmFlaggerGPU(cfg)
    rng = np.random.default_rng(0)
    q = -np.concatenate([np.linspace(0.0, 120.0, 200001), rng.exponential(5.0, 200000), [0.0, 1e-300, 1e-17, 700.0, 708.0]])
    got = gpu.debug_exp(q)
    want = np.exp(q)
    ulp = np.abs(got - want) / np.spacing(want)
    assert ulp.max() <= 2.0, ulp.max()  # Estrin evaluation: <= 2 ulp from glibc (libdevice exp itself is 1 ulp)
    # below the clamp the result is the clamp value (far under the 1e-40 floor applied to every pdf)
    assert gpu.debug_exp(np.array([-1e4]))[0] == gpu.debug_exp(np.array([-708.0]))[0] < 1e-300
    gpu.close()


import golden_util  # noqa: E402


@pytest.mark.parametrize("name", golden_util.NAMES)
def test_gpu_against_reference_golden(name):
    """CUDA path against vectors produced by the UNMODIFIED reference (tests/golden/make_golden.py)."""
    g, wl = golden_util.load(name)
    cfg, alpha = g["cfg"].copy(), g["alpha"]
    gpu = api.HmmFlaggerGPU(cfg, wl)
    stats, ll, labels = gpu.em_iteration(alpha, g["params0"])
    assert abs(ll - float(g["loglik"])) <= TOL_LOGLIK * abs(float(g["loglik"]))
    assert np.array_equal(labels, g["labels"])
    post = gpu.posteriors()
    big = g["posteriors"] > 1e-200
    assert np.all(_rel(post[big], g["posteriors"][big]) <= TOL_POST)
    so = _abi.stats_as_flat(g["stats"])
    assert np.all(np.abs(_abi.stats_as_flat(stats) - so) <= TOL_STATS * np.abs(so).max())
    cl = gpu.chunk_logliks()
    assert np.all(np.abs(cl - g["chunk_logliks"]) <= 1e-9 * np.abs(g["chunk_logliks"]) + 1e-9)
    assert abs(gpu.forward_only(alpha, g["params1"]) - float(g["fwd_only_loglik"])) <= TOL_LOGLIK * abs(float(g["fwd_only_loglik"]))
    pg, llg, labg = gpu.run_em(alpha, g["params0"], 5, tol=1e-12)
    assert np.all(np.abs(llg - g["em_logliks"]) <= TOL_LOGLIK * np.abs(g["em_logliks"]))
    assert np.array_equal(labg, g["em_labels"])
    assert np.allclose(_abi.params_as_flat(pg), _abi.params_as_flat(g["em_params"]), rtol=1e-8, atol=0)
    gpu.close()


def test_error_scale_underflow_is_reported(orc):
    """The reference exits with "scale is very low" when a forward scale drops below 1e-50 (hmm.c:412-415).  Reachable
    only with degenerate inputs: here Dup and Col are masked out by the MAPQ ratio, Err has zero density above its
    truncation point and every transition into Hap is zero, so no path survives."""
    wl = synth.small_mixed(n_regions=1, seed=4)
    wl.cov[:] = 40
    wl.cov_high_mapq[:] = 20  # ratio ~0.5: Dup invalid (> 0.25) and Col invalid (< 0.75)
    cfg = _abi.make_config(n_col_comps=3)
    params = api.model_init(cfg, wl.region_coverages, wl.window_len)
    params["trans"][0, :4, 2] = 0.0
    assert orc.estep(cfg, wl, np.zeros((4, 4)), params)["rc"] == _abi.ERR_SCALE_UNDERFLOW
    gpu = api.HmmFlaggerGPU(cfg, wl)
    with pytest.raises(api.HfgError) as e:
        gpu.em_iteration(np.zeros((4, 4)), params)
    assert e.value.code == _abi.ERR_SCALE_UNDERFLOW and "scale is very low" in str(e.value)
    gpu.close()


@pytest.mark.parametrize("n_regions,K", [(1, 10), (7, 10), (20, 4), (32, 6)])
def test_large_models_use_the_smaller_cta(orc, n_regions, K):
    """Many mixture components / regions: the per-thread statistics columns no longer fit shared memory with 512-thread
    CTAs, the library switches to the 256-thread instantiation of the same kernel.  K = 10 is the CLI's upper clamp
    (src/hmm_flagger.c:1012-1013); region indices go up to 63 (ptBlock.c:294-304)."""
    rng = np.random.default_rng(n_regions * 100 + K)
    wl = synth.small_mixed(n_regions=1, seed=60 + K)
    wl.region[:] = np.repeat(rng.integers(0, n_regions, size=wl.n_windows // 7 + 1), 7)[:wl.n_windows].astype(np.uint8)
    wl.region_coverages = rng.integers(25, 60, size=n_regions).astype(np.int32)
    cfg = _abi.make_config(n_regions=n_regions, n_col_comps=K)
    params = api.model_init(cfg, wl.region_coverages, wl.window_len)
    gpu = api.HmmFlaggerGPU(cfg, wl)
    out = orc.estep(cfg, wl, synth.HIFI_ALPHA, params)
    stats, ll, labels, _ = _check_estep(gpu, out, wl, synth.HIFI_ALPHA, params)
    p2, _ = api.mstep(cfg, params, stats)
    out2 = orc.estep(cfg, wl, synth.HIFI_ALPHA, p2)
    _check_estep(gpu, out2, wl, synth.HIFI_ALPHA, p2)
    gpu.close()


def test_device_em_loop_stops_at_convergence_like_the_reference_loop(orc):
    """hfg_run_em queues every iteration on the device (M-step in the kernel's tail, hfg_em_*): with a loose tolerance
    it must stop after the same number of E-steps as the host loop of the reference (src/hmm_flagger.c:337-431) and
    return the same parameters and labels; the hand-driven hfg_em_begin/enqueue/finish sequence must agree with it."""
    wl = synth.small_mixed(n_regions=3, seed=5)
    K = api.best_num_collapsed_comps(int(wl.cov.max()), wl.region_coverages)
    cfg = _abi.make_config(n_regions=3, n_col_comps=K)
    params = api.model_init(cfg, wl.region_coverages, wl.window_len)
    gpu = api.HmmFlaggerGPU(cfg, wl)
    n_seen = set()
    for tol in (0.9, 0.5, 1e-12):
        eo = orc.run_em(cfg, wl, synth.HIFI_ALPHA, params, 12, tol=tol)
        pg, llg, labg = gpu.run_em(synth.HIFI_ALPHA, params, 12, tol=tol)
        n_seen.add(len(llg))
        assert len(llg) == len(eo["logliks"])
        assert np.all(np.abs(llg - eo["logliks"]) <= TOL_LOGLIK * np.abs(eo["logliks"]))
        assert np.array_equal(labg, eo["labels"])
        assert np.allclose(_abi.params_as_flat(pg), _abi.params_as_flat(eo["params"]), rtol=1e-8, atol=0)
    assert len(n_seen) > 1 and max(n_seen) == 13  # the loose tolerances did stop early
    # by hand: 4 iterations + final pass, per-iteration device times available afterwards
    gpu.em_begin(synth.HIFI_ALPHA, params, tol=1e-12, max_esteps=5)
    for _ in range(4):
        gpu.em_enqueue()
    gpu.em_enqueue(final_pass=True)
    p2, ll2, conv, lab2 = gpu.em_finish()
    eo = orc.run_em(cfg, wl, synth.HIFI_ALPHA, params, 4, tol=1e-12)
    assert not conv and len(ll2) == 5
    assert np.all(np.abs(ll2 - eo["logliks"]) <= TOL_LOGLIK * np.abs(eo["logliks"]))
    assert np.array_equal(lab2, eo["labels"])
    assert all(gpu.em_enqueued_ms(i) > 0 for i in range(5))
    # the posteriors of the last E-step are those of the final parameters
    post = gpu.posteriors()
    assert np.array_equal(post.argmax(axis=1).astype(np.int8), lab2)
    gpu.close()


def test_pinned_label_buffer_and_repeated_calls(orc):
    """Labels written by the device straight into a page-locked caller buffer (hfg_host_alloc) equal those delivered
    through the staging copy, call after call (the kernel clears its own error flags and writes the statistics block
    into pinned memory itself), also when the two kinds of buffer alternate."""
    wl = synth.small_mixed(n_regions=3, seed=17)
    K = api.best_num_collapsed_comps(int(wl.cov.max()), wl.region_coverages)
    cfg = _abi.make_config(n_regions=3, n_col_comps=K)
    params = api.model_init(cfg, wl.region_coverages, wl.window_len)
    gpu = api.HmmFlaggerGPU(cfg, wl)
    out = orc.estep(cfg, wl, synth.HIFI_ALPHA, params)
    pinned = api.PinnedArray(wl.n_windows, np.int8)
    for rep in range(3):
        pinned.array[:] = -7
        s1, ll1, lab1 = gpu.em_iteration(synth.HIFI_ALPHA, params, labels=pinned.array)
        assert lab1 is pinned.array and np.array_equal(pinned.array, out["labels"])
        s2, ll2, lab2 = gpu.em_iteration(synth.HIFI_ALPHA, params)
        assert np.array_equal(lab2, out["labels"]) and ll1 == ll2
        assert np.array_equal(_abi.stats_as_flat(s1), _abi.stats_as_flat(s2))
        assert abs(ll1 - out["loglik"]) <= TOL_LOGLIK * abs(out["loglik"])
    gpu.close()
    pinned.free()


def test_blocking_fast_path_equals_graph_path(orc):
    """A single-region model takes the one-launch path of the blocking calls (parameters in the kernel arguments, completion
    word polled in pinned memory, no events: hfg_api.cu::run_blocking); with timing events requested the same call replays the
    captured graph (upload node, kernel, events, stream synchronisation).  Same kernel, same bits: statistics, log-likelihood
    and labels, with a page-locked label buffer and with a pageable one, over changing parameters; and the oracle's values."""
    wl = synth.small_mixed(n_regions=1, seed=29)
    K = api.best_num_collapsed_comps(int(wl.cov.max()), wl.region_coverages)
    cfg = _abi.make_config(n_regions=1, n_col_comps=K)
    params = api.model_init(cfg, wl.region_coverages, wl.window_len)
    fast, graph = api.HmmFlaggerGPU(cfg, wl), api.HmmFlaggerGPU(cfg, wl, timing=True)
    pinned = api.PinnedArray(wl.n_windows, np.int8)
    try:
        for it in range(4):
            out = orc.estep(cfg, wl, synth.HIFI_ALPHA, params)
            sf, llf, labf = fast.em_iteration(synth.HIFI_ALPHA, params, labels=pinned.array if it % 2 == 0 else None)
            labf = labf.copy()
            sg, llg, labg = graph.em_iteration(synth.HIFI_ALPHA, params)
            assert fast.last_estep_kernel_ms() < 0 and graph.last_estep_kernel_ms() > 0  # which path each one took
            assert llf == llg and np.array_equal(labf, labg) and np.array_equal(_abi.stats_as_flat(sf), _abi.stats_as_flat(sg))
            assert np.array_equal(labf, out["labels"]) and abs(llf - out["loglik"]) <= TOL_LOGLIK * abs(out["loglik"])
            assert abs(fast.forward_only(synth.HIFI_ALPHA, params) - llf) <= 1e-12 * abs(llf)
            params, _ = api.mstep(cfg, params, sf, tol=1e-12)
    finally:
        fast.close()
        graph.close()
        pinned.free()


def test_blocking_fast_path_on_recycled_result_blocks(orc):
    """The completion word lives behind the pinned result block, and those blocks are recycled from context to context: a
    new context must never take a predecessor's word for its own (contexts opened and closed in turn over different inputs,
    each checked against the oracle)."""
    for seed in (31, 32, 33, 34):
        wl = synth.small_mixed(n_regions=1, seed=seed)
        K = api.best_num_collapsed_comps(int(wl.cov.max()), wl.region_coverages)
        cfg = _abi.make_config(n_regions=1, n_col_comps=K)
        params = api.model_init(cfg, wl.region_coverages, wl.window_len)
        out = orc.estep(cfg, wl, synth.HIFI_ALPHA, params)
        gpu = api.HmmFlaggerGPU(cfg, wl)
        stats, ll, labels = gpu.em_iteration(synth.HIFI_ALPHA, params)
        gpu.close()
        assert abs(ll - out["loglik"]) <= TOL_LOGLIK * abs(out["loglik"]) and np.array_equal(labels, out["labels"])
        so = _abi.stats_as_flat(out["stats"])
        assert np.all(np.abs(_abi.stats_as_flat(stats) - so) <= TOL_STATS * np.abs(so).max())


def test_many_distinct_observation_keys(orc):
    """The kernel evaluates emissions once per distinct (x, previous x, region, mask, beta) key.  Coverage drawn uniformly
    from 0..250 with random MAPQ / clipping fractions makes almost every window its own key (the worst case for the key
    table and the per-key statistics lists, one window per tile): results must not depend on how many keys there are."""
    import dataclasses
    wl = synth.small_mixed(n_regions=3, seed=23)
    rng = np.random.default_rng(7)
    W = wl.n_windows
    cov = rng.integers(0, 251, W).astype(np.uint16)
    mapq = (cov * rng.choice([0.0, 0.1, 0.5, 0.9, 1.0], W)).astype(np.uint16)
    clip = (cov * rng.choice([0.0, 0.0, 0.0, 1.0], W)).astype(np.uint16)
    wl = dataclasses.replace(wl, cov=cov, cov_high_mapq=mapq, cov_high_clip=clip)
    K = api.best_num_collapsed_comps(int(wl.cov.max()), wl.region_coverages)
    cfg = _abi.make_config(n_regions=3, n_col_comps=K)
    ok, summary = api.layout_check(cfg, wl, 2048)
    assert ok and summary[4] > 0.8 * W  # nearly one key per window
    params = api.model_init(cfg, wl.region_coverages, wl.window_len)
    gpu = api.HmmFlaggerGPU(cfg, wl)
    out = orc.estep(cfg, wl, synth.HIFI_ALPHA, params)
    assert out["rc"] == 0
    _check_estep(gpu, out, wl, synth.HIFI_ALPHA, params)
    pg, llg, labg = gpu.run_em(synth.HIFI_ALPHA, params, 3, tol=1e-12)
    eo = orc.run_em(cfg, wl, synth.HIFI_ALPHA, params, 3, tol=1e-12)
    assert eo["rc"] == 0 and len(llg) == len(eo["logliks"])
    assert np.all(np.abs(llg - eo["logliks"]) <= TOL_LOGLIK * np.abs(eo["logliks"]))
    assert np.array_equal(labg, eo["labels"])
    gpu.close()


@pytest.mark.parametrize("factory,n_regions,adjust", [
    (lambda: synth.small_mixed(n_regions=3, seed=1), 3, True),
    (lambda: synth.small_mixed(n_regions=7, seed=2), 7, False),
    (lambda: synth.config1(), 1, True),
    (lambda: synth.config3(n_contigs=700, seed=3), 1, True),
    (lambda: synth.config2(total_bp=400_000_000, seed=4), 1, True),
    (lambda: synth.config4(total_bp=300_000_000, seed=5), 7, True),
])
def test_device_layout_equals_host_layout(factory, n_regions, adjust):
    """hfg_set_chunks builds the observation keys, their window lists and the statistics tiles on the device (radix sorts
    and scans, hfg_layout_dev.cuh); the host builder (pthread hash tables, hfg_layout.c) must produce the same bits in
    every table: key words, key descriptors, betas, lists, tiles, region offsets."""
    wl = factory()
    cfg = _abi.make_config(n_regions=n_regions, adjust_contig_ends=adjust, mean_read_length=wl.avg_alignment_len)
    gpu = api.HmmFlaggerGPU(cfg, wl)
    ok, msg = gpu.layout_matches_host(wl)
    assert ok, msg
    gpu.close()


def test_host_built_layout_gives_the_same_results(orc, monkeypatch):
    """HFG_HOST_LAYOUT=1 makes hfg_set_chunks use the host key builder (the checker of the device build): same results,
    bit for bit, as the default device-built layout."""
    wl = synth.small_mixed(n_regions=3, seed=29)
    K = api.best_num_collapsed_comps(int(wl.cov.max()), wl.region_coverages)
    cfg = _abi.make_config(n_regions=3, n_col_comps=K)
    params = api.model_init(cfg, wl.region_coverages, wl.window_len)
    dev = api.HmmFlaggerGPU(cfg, wl)
    s_dev, ll_dev, lab_dev = dev.em_iteration(synth.HIFI_ALPHA, params)
    dev.close()
    monkeypatch.setenv("HFG_HOST_LAYOUT", "1")
    host = api.HmmFlaggerGPU(cfg, wl)
    ok, msg = host.layout_matches_host(wl)
    assert ok, msg
    s_host, ll_host, lab_host = host.em_iteration(synth.HIFI_ALPHA, params)
    host.close()
    assert ll_dev == ll_host and np.array_equal(lab_dev, lab_host)
    assert np.array_equal(_abi.stats_as_flat(s_dev), _abi.stats_as_flat(s_host))
    out = orc.estep(cfg, wl, synth.HIFI_ALPHA, params)
    assert np.array_equal(lab_host, out["labels"]) and abs(ll_host - out["loglik"]) <= TOL_LOGLIK * abs(out["loglik"])
