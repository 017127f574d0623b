"""GPU: the negative-binomial kernel instantiation (hfg_estep_v3_kernel<THREADS, true> + the host fold in
hfg_api.cu::run_blocking_nb) against the oracle, the golden vectors, the reference binary (stand-alone CLI) and the drop-in
binding.

Bars: labels identical, log-likelihood within 1e-9 relative, statistics within 1e-9 of the region's largest (the reference
sums the histogram per chunk, the product per tile: rounding only), EM trajectories accordingly."""
import os

import numpy as np
import pytest

import golden_util
from flagger_b200 import _abi, api, synth

pytestmark = pytest.mark.gpu

NB = _abi.MODEL_NEGATIVE_BINOMIAL


@pytest.mark.parametrize("n_regions,seed", [(1, 81), (3, 82), (7, 83)])
def test_nb_estep_matches_oracle(orc, n_regions, seed):
    wl = synth.small_mixed(n_regions=n_regions, seed=seed)
    wl.cov[7:10] = 250  # folded into bin 249 of the histogram
    K = api.best_num_collapsed_comps(int(wl.cov.max()), wl.region_coverages)
    cfg = _abi.make_config(n_regions=n_regions, n_col_comps=K, model_type=NB, mean_read_length=wl.avg_alignment_len)
    p = api.model_init(cfg, wl.region_coverages, wl.window_len)
    want = orc.estep(cfg, wl, np.zeros((4, 4)), p)
    gpu = api.HmmFlaggerGPU(cfg, wl)
    try:
        stats, ll, labels = gpu.em_iteration(synth.HIFI_ALPHA, p)  # alpha is ignored by this model
        assert want["rc"] == 0 and abs(ll - want["loglik"]) <= 1e-9 * abs(want["loglik"])
        assert np.array_equal(labels, want["labels"])
        sg, so = _abi.stats_as_flat(stats), _abi.stats_as_flat(want["stats"])
        assert np.all(np.abs(sg - so) <= 1e-9 * np.abs(so).max())
        post = gpu.posteriors()
        big = want["posteriors"] > 1e-200
        assert np.all(np.abs(post[big] - want["posteriors"][big]) <= 1e-5 * want["posteriors"][big])
        assert abs(gpu.forward_only(synth.HIFI_ALPHA, p) - want["loglik"]) <= 1e-9 * abs(want["loglik"])
    finally:
        gpu.close()


@pytest.mark.parametrize("name", golden_util.NB_NAMES)
def test_nb_em_run_matches_reference_golden(name):
    g, wl = golden_util.load(name, ".nb.npz")
    cfg = g["cfg"]
    gpu = api.HmmFlaggerGPU(cfg, wl)
    try:
        params, logliks, labels = gpu.run_em(g["alpha"], g["params0"], 5, tol=1e-12)
        assert len(logliks) == len(g["em_logliks"])
        assert np.all(np.abs(logliks - g["em_logliks"]) <= 1e-9 * np.abs(g["em_logliks"]))
        assert np.array_equal(labels, g["em_labels"])
        a, b = params.view(np.float64).reshape(-1), g["em_params"].view(np.float64).reshape(-1)
        assert np.all(np.abs(a - b) <= 1e-8 * np.maximum(np.abs(b), 1e-12))
    finally:
        gpu.close()


def test_nb_cli_run_matches_reference_binary(tmp_path):
    """Whole run of the stand-alone binary with --modelType negative_binomial against the unmodified reference binary."""
    import test_standalone_cli as cli
    if not os.path.exists(cli.REF):
        pytest.skip("oracle/_ref/hmm_flagger_ref was not built")
    inp, alpha = cli._inputs(tmp_path, "cov")
    ref_out, out = str(tmp_path / "ref"), str(tmp_path / "cli")
    cli._run(cli.REF, inp, ref_out, extra=("-m", "negative_binomial"))
    cli._run(cli.CLI, inp, out, extra=("-m", "negative_binomial"))
    assert cli._read(os.path.join(ref_out, "final_flagger_prediction.bed")) == cli._read(os.path.join(out, "final_flagger_prediction.bed"))
    for name in ("loglikelihood.tsv", "emission_final.tsv", "transition_final.tsv"):
        a, b = cli._table(os.path.join(out, name)), cli._table(os.path.join(ref_out, name))
        assert a.shape == b.shape and np.allclose(a, b, rtol=1e-4, atol=1e-8), name
    for name in ("prediction_summary_initial.tsv", "prediction_summary_final.tsv"):
        assert cli._read(os.path.join(ref_out, name)) == cli._read(os.path.join(out, name)), name


def test_nb_dropin_binary_matches_reference_binary(tmp_path):
    """The reference's own CLI with its E-step replaced by libhfg (integration/hmm_estep_cuda.c), negative-binomial model."""
    import test_dropin_binary as dropin
    if not (os.path.exists(dropin.REF) and os.path.exists(dropin.GPU)):
        pytest.skip("oracle/_ref binaries were not built")
    wl = synth.small_mixed(n_regions=3, seed=56)
    inp, alpha = str(tmp_path / "in.bin"), str(tmp_path / "alpha.tsv")
    from flagger_b200 import binfmt
    binfmt.write_bin(wl, inp)
    binfmt.write_alpha_tsv(np.zeros((4, 4)), alpha)
    ref_out, gpu_out = str(tmp_path / "ref"), str(tmp_path / "gpu")
    dropin._run(dropin.REF, inp, ref_out, alpha, extra=("-m", "negative_binomial"))
    dropin._run(dropin.GPU, inp, gpu_out, alpha, extra=("-m", "negative_binomial"))
    assert open(os.path.join(ref_out, "final_flagger_prediction.bed")).read() == \
        open(os.path.join(gpu_out, "final_flagger_prediction.bed")).read()
    a, b = dropin._table(os.path.join(ref_out, "loglikelihood.tsv")), dropin._table(os.path.join(gpu_out, "loglikelihood.tsv"))
    assert a.shape == b.shape and np.allclose(a, b, rtol=1e-6, atol=2e-4)


@pytest.mark.parametrize("n_regions,seed,iters", [(1, 91, 6), (3, 92, 4), (7, 93, 3)])
def test_nb_device_resident_loop_equals_host_loop(monkeypatch, n_regions, seed, iters):
    """hfg_run_em for the negative-binomial model: the device-resident loop (pmf, histogram fold, estimator update and M-step on
    the device in fp64, hfg_nb_dev.cuh) against the same loop with the host between the iterations (libm table, long-double
    digamma: HFG_NB_HOST_LOOP=1).  Rounding only: labels identical, log-likelihoods and parameters to 1e-9 / 1e-7."""
    wl = synth.small_mixed(n_regions=n_regions, seed=seed)
    K = api.best_num_collapsed_comps(int(wl.cov.max()), wl.region_coverages)
    cfg = _abi.make_config(n_regions=n_regions, n_col_comps=K, model_type=NB, mean_read_length=wl.avg_alignment_len)
    p0 = api.model_init(cfg, wl.region_coverages, wl.window_len)
    gpu = api.HmmFlaggerGPU(cfg, wl)
    try:
        pd, lld, labd = gpu.run_em(np.zeros((4, 4)), p0, iters, tol=1e-12)
        monkeypatch.setenv("HFG_NB_HOST_LOOP", "1")
        ph, llh, labh = gpu.run_em(np.zeros((4, 4)), p0, iters, tol=1e-12)
        monkeypatch.delenv("HFG_NB_HOST_LOOP")
        assert len(lld) == len(llh) == iters + 1
        assert np.all(np.abs(lld - llh) <= 1e-9 * np.abs(llh)), (lld, llh)
        assert np.array_equal(labd, labh), int((labd != labh).sum())
        a, b = _abi.params_as_flat(pd), _abi.params_as_flat(ph)
        nz = np.abs(b) > 0
        assert np.all(np.abs(a[nz] - b[nz]) <= 1e-7 * np.abs(b[nz])), float(np.max(np.abs(a[nz] - b[nz]) / np.abs(b[nz])))
        # the loop can be driven step by step as for the other models
        gpu.em_begin(np.zeros((4, 4)), p0, tol=1e-12, max_esteps=3)
        gpu.em_enqueue()
        gpu.em_enqueue()
        gpu.em_enqueue(final_pass=True)
        p3, ll3, _, lab3 = gpu.em_finish()
        assert len(ll3) == 3 and np.array_equal(ll3[:2], lld[:2])
    finally:
        gpu.close()
