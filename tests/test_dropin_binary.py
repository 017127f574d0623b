"""GPU, end to end: the reference's own `hmm_flagger` CLI with its E-step replaced by libhfg (integration/
hmm_estep_cuda.c, built into oracle/_ref/hmm_flagger_gpu) against the unmodified reference binary, on the same input
file and flags.  Compares the files a user gets: final BED, log-likelihood table, parameter tables, summary tables."""
import os
import subprocess

import numpy as np
import pytest

from flagger_b200 import binfmt, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "hmm_flagger_ref")
GPU = os.path.join(ROOT, "oracle", "_ref", "hmm_flagger_gpu")


def _run(binary, inp, out, alpha, extra=()):
    os.makedirs(out, exist_ok=True)
    cmd = [binary, "-i", inp, "-o", out, "-A", alpha, "-W", "4000", "-C", "1000000", "-n", "6", "-t", "1e-12",
           "-l", "Err,Dup,Hap,Col", "-@", "4", *extra]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stderr


def _table(path):
    rows = []
    for line in open(path):
        if line.startswith("#"):
            continue
        for tok in line.rstrip("\n").split("\t"):
            for v in tok.split(","):
                try:
                    rows.append(float(v))
                except ValueError:
                    pass
    return np.array(rows)


@pytest.mark.parametrize("kind", ["bin", "cov.gz"])
def test_cli_outputs_match_reference(tmp_path, kind):
    if not (os.path.exists(REF) and os.path.exists(GPU)):
        pytest.skip("oracle/_ref binaries were not built (reference tree not mounted at build time)")
    wl = synth.small_mixed(n_regions=3, seed=55)
    inp = str(tmp_path / f"in.{kind}")
    (binfmt.write_bin if kind == "bin" else binfmt.write_cov)(wl, inp)
    alpha = str(tmp_path / "alpha.tsv")
    binfmt.write_alpha_tsv(synth.HIFI_ALPHA, alpha)
    ref_out, gpu_out = str(tmp_path / "ref"), str(tmp_path / "gpu")
    _run(REF, inp, ref_out, alpha, extra=("-w", "-P"))
    log = _run(GPU, inp, gpu_out, alpha, extra=("-w", "-P"))
    assert "resident on GPU" in log
    # labels: the BED a user gets must be byte-identical
    assert open(os.path.join(ref_out, "final_flagger_prediction.bed")).read() == \
        open(os.path.join(gpu_out, "final_flagger_prediction.bed")).read()
    for name in ("prediction_summary_initial.tsv", "prediction_summary_final.tsv"):
        assert open(os.path.join(ref_out, name)).read() == open(os.path.join(gpu_out, name)).read(), name
    # log-likelihoods are printed with 4 decimals
    a, b = _table(os.path.join(ref_out, "loglikelihood.tsv")), _table(os.path.join(gpu_out, "loglikelihood.tsv"))
    assert a.shape == b.shape and np.all(np.abs(a - b) <= 2e-4)
    # parameters are printed with %.5e
    for name in sorted(os.listdir(ref_out)):
        if name.startswith(("emission_", "transition_")):
            a, b = _table(os.path.join(ref_out, name)), _table(os.path.join(gpu_out, name))
            assert a.shape == b.shape and np.allclose(a, b, rtol=2e-5, atol=1e-12), name
    # posteriors are printed with %.2f
    a = _table(os.path.join(ref_out, "posterior_prediction_final.bed"))
    b = _table(os.path.join(gpu_out, "posterior_prediction_final.bed"))
    assert a.shape == b.shape and np.all(np.abs(a - b) <= 0.0101)
    pa = [ln.split("\t")[-1] for ln in open(os.path.join(ref_out, "posterior_prediction_final.bed"))]
    pb = [ln.split("\t")[-1] for ln in open(os.path.join(gpu_out, "posterior_prediction_final.bed"))]
    assert pa == pb


def test_cli_accelerate_matches_reference(tmp_path):
    """--accelerate (SQUAREM, hmm.c:820-1098) runs 3 E-steps and >= 1 forward-only pass per outer iteration through the
    seam (EM_runForwardForList -> hfg_forward_only); the host extrapolation stays the reference's code."""
    if not (os.path.exists(REF) and os.path.exists(GPU)):
        pytest.skip("oracle/_ref binaries were not built (reference tree not mounted at build time)")
    wl = synth.small_mixed(n_regions=1, seed=56)
    inp = str(tmp_path / "in.bin")
    binfmt.write_bin(wl, inp)
    alpha = str(tmp_path / "alpha.tsv")
    binfmt.write_alpha_tsv(synth.HIFI_ALPHA, alpha)
    ref_out, gpu_out = str(tmp_path / "ref"), str(tmp_path / "gpu")
    _run(REF, inp, ref_out, alpha, extra=("-s",))
    _run(GPU, inp, gpu_out, alpha, extra=("-s",))
    assert open(os.path.join(ref_out, "final_flagger_prediction.bed")).read() == \
        open(os.path.join(gpu_out, "final_flagger_prediction.bed")).read()
    a, b = _table(os.path.join(ref_out, "loglikelihood.tsv")), _table(os.path.join(gpu_out, "loglikelihood.tsv"))
    assert a.shape == b.shape and np.all(np.abs(a - b) <= 2e-4)
    for name in ("emission_final.tsv", "transition_final.tsv"):
        a, b = _table(os.path.join(ref_out, name)), _table(os.path.join(gpu_out, name))
        assert a.shape == b.shape and np.allclose(a, b, rtol=2e-5, atol=1e-12), name
