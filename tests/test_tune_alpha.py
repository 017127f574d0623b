"""CPU: the alpha-tuning driver (flagger_b200/tune_alpha.py) with the C restatement standing in for the GPU engine:
parametrisation and objective against a real run of the unmodified reference binary, and the search loop's contract."""
import os
import subprocess

import numpy as np
import pandas as pd
import pytest

from flagger_b200 import _abi, api, binfmt, synth, tune_alpha

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "hmm_flagger_ref")
needs_ref = pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/hmm_flagger_ref was not built")


class OracleEngine:
    """Test stand-in for tune_alpha.GpuEngine: same model set-up, the EM run done by the oracle (test infrastructure)."""

    def __init__(self, cov, orc, em_iterations, tol):
        wl = cov.workload
        K = api.best_num_collapsed_comps(int(wl.cov.max()), wl.region_coverages)
        self.cfg = _abi.make_config(n_regions=wl.n_regions, n_col_comps=K, mean_read_length=wl.avg_alignment_len)
        self.params0 = api.model_init(self.cfg, wl.region_coverages, wl.window_len)
        self.wl, self.orc, self.n, self.tol = wl, orc, em_iterations, tol

    def __call__(self, alpha):
        return self.orc.run_em(self.cfg, self.wl, np.ascontiguousarray(alpha, np.float64), self.params0, self.n, tol=self.tol)["labels"]


def test_parametrisation_round_trip():
    x = np.arange(1, 11) / 20.0
    a = tune_alpha.x_to_alpha(x)
    assert np.count_nonzero(a) == 10 and np.allclose(tune_alpha.alpha_to_x(a), x)
    # the positions of convertXToAlphaMatrix (programs/src/tune_alpha_hmm_flagger.py:54-66)
    assert (a[0, 0], a[0, 2], a[1, 1], a[1, 2], a[2, 0], a[2, 1], a[2, 2], a[2, 3], a[3, 2], a[3, 3]) == tuple(x)
    assert np.allclose(tune_alpha.alpha_to_x(synth.HIFI_ALPHA)[[0, 2, 6, 9]], [synth.HIFI_ALPHA[0, 0], synth.HIFI_ALPHA[1, 1],
                                                                                synth.HIFI_ALPHA[2, 2], synth.HIFI_ALPHA[3, 3]])


@needs_ref
def test_objective_equals_the_reference_drivers_score(tmp_path, orc):
    """One candidate alpha: Objective.score == the combined score the reference driver would compute from the files of a
    `hmm_flagger --alpha ...` run on the same input (tune_alpha_hmm_flagger.py:82-111,170-205)."""
    inp = str(tmp_path / "train.cov")
    binfmt.write_random_rle_cov(inp, [9000, 310_000, 1_250_000, 123_457], seed=21, n_regions=1, with_truth=True)
    x = np.array([0.4, 0.1, 0.3, 0.05, 0.0, 0.2, 0.5, 0.1, 0.15, 0.45])
    alpha_tsv = str(tmp_path / "alpha.tsv")
    np.savetxt(alpha_tsv, tune_alpha.x_to_alpha(x), delimiter="\t", fmt="%.3f")
    out = str(tmp_path / "ref")
    os.makedirs(out)
    cmd = [REF, "--alpha", alpha_tsv, "--input", inp, "--outputDir", out, "--modelType", "trunc_exp_gaussian", "-W", "4000",
           "-C", "1000000", "-n", "3", "-t", "1e-12", "-l", "Err,Dup,Hap,Col", "-@", "2"]
    assert subprocess.run(cmd, capture_output=True, text=True, timeout=600).returncode == 0
    t = pd.read_csv(os.path.join(out, "prediction_summary_final.benchmarking.tsv"), sep="\t").rename(columns={"#Metric_Type": "Metric_Type"})
    want = []
    for metric in ("overlap_based", "base_level"):
        row = t[(t.Metric_Type == metric) & (t.Category_Name == "whole_genome") & (t.Size_Bin_Name == "ALL_SIZES") &
                (t.Label == "HARMONIC_MEAN_NO_HAP")]
        want.append(float(row["F1-Score"].item()))
    a = pd.read_csv(os.path.join(out, "prediction_summary_final.benchmarking.auN_ratio.tsv"), sep="\t")
    want.append(100 * float(a[(a.Category_Name == "whole_genome") & (a.Size_Bin_Name == "ALL_SIZES") &
                              (a.Label == "HARMONIC_MEAN")]["auN_Ratio"].item()))
    cov = binfmt.NativeCov(inp, 1_000_000, 4000)
    obj = tune_alpha.Objective([(cov, OracleEngine(cov, orc, 3, 1e-12))])
    got = obj.score(x)
    assert abs(got - sum(want) / 3.0) < 1e-9, (got, want, obj.history[-1])
    cov.close()


def test_search_loop_contract(tmp_path, orc):
    inp = str(tmp_path / "train.cov.gz")
    binfmt.write_random_rle_cov(inp, [310_000, 650_000], seed=22, n_regions=1, with_truth=True)
    cov = binfmt.NativeCov(inp, 1_000_000, 4000)

    def run(seed):
        obj = tune_alpha.Objective([(cov, OracleEngine(cov, orc, 2, 1e-12))])
        best_x, best = tune_alpha.optimise(obj, 0.0, 0.8, n_start=3, n_iter=6, candidate_alpha=synth.HIFI_ALPHA, seed=seed)
        return obj, best_x, best

    obj, best_x, best = run(7)
    assert len(obj.history) == 9 and [h[0] for h in obj.history[:3]] == ["start"] * 3
    assert np.allclose(obj.history[0][1], tune_alpha.alpha_to_x(synth.HIFI_ALPHA))  # the candidate matrix is the first start point
    assert all(np.all(h[1] >= 0.0) and np.all(h[1] <= 0.8) for h in obj.history)
    assert best == max(h[2] for h in obj.history) and best >= max(h[2] for h in obj.history[:3])
    obj2, best_x2, best2 = run(7)
    assert best2 == best and np.array_equal(best_x2, best_x)  # seeded: reproducible
    # rounds of proposals (the batched engine's mode): same contract, start points scored as one batch
    obj3 = tune_alpha.Objective([(cov, OracleEngine(cov, orc, 2, 1e-12))])
    best_x3, best3 = tune_alpha.optimise(obj3, 0.0, 0.8, n_start=3, n_iter=6, candidate_alpha=synth.HIFI_ALPHA, seed=7, batch=4)
    assert len(obj3.history) == 9 and [h[0] for h in obj3.history] == ["start"] * 3 + ["iteration"] * 6
    assert [h[2] for h in obj3.history[:3]] == [h[2] for h in obj.history[:3]]  # same start points, same scores
    assert best3 == max(h[2] for h in obj3.history) and all(np.all(h[1] >= 0.0) and np.all(h[1] <= 0.8) for h in obj3.history)
    cov.close()
