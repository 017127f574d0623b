"""CPU: the summary-table writer on flat labels (flagger_b200/csrc/hfg_summary.c, include/hfg_io.h) against the
unmodified reference binary: prediction_summary_final.tsv of a real `hmm_flagger` run -- and, for inputs with truth
labels, prediction_summary_final.benchmarking.tsv / .benchmarking.auN_ratio.tsv -- reproduced from the run's own final
BED (expanded to one label per window) and the input's truth labels.  Byte-identical files: all three metric types
(overlap_based, base_level, truth_based_auN), all four comparison types, the reference's row order and formats."""
import os
import subprocess

import numpy as np
import pytest

from flagger_b200 import _abi, binfmt

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "hmm_flagger_ref")
needs_ref = pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/hmm_flagger_ref was not built")


def _labels_from_bed(bed, wl):
    """One label per window from the reference's final BED (blocks tile every contig; default --minimumLengths 0)."""
    idx = {n: i for i, n in enumerate(_abi.STATE_NAMES + ("Unk",))}
    blocks = {}
    for line in open(bed):
        if line.startswith("track"):
            continue
        t = line.split("\t")
        blocks.setdefault(t[0], []).append((int(t[1]), int(t[2]), idx[t[3]]))
    labels = np.full(wl.n_windows, -1, np.int8)
    for c, name in zip(wl.chunks, wl.contig_names):
        starts = np.array([b[0] for b in blocks[name]])
        for i in range(int(c["n_windows"])):
            s = int(c["s"]) + i * int(c["window_len"])
            b = blocks[name][int(np.searchsorted(starts, s, side="right")) - 1]
            assert b[0] <= s < b[1]
            labels[int(c["offset"]) + i] = b[2] if b[2] < 4 else -1
    return labels


@needs_ref
@pytest.mark.parametrize("with_truth,kind,seed,bins", [(False, "cov", 5, False), (True, "cov.gz", 6, False), (True, "cov", 7, True),
                                                         (True, "cov", 8, False), (False, "cov.gz", 9, True)])
def test_summary_tsv_matches_reference_run(tmp_path, with_truth, kind, seed, bins):
    inp = str(tmp_path / f"in.{kind}")
    binfmt.write_random_rle_cov(inp, [4000, 9000, 310_000, 1_250_000, 123_457], seed=seed, n_regions=3, with_truth=with_truth)
    out = str(tmp_path / "ref")
    os.makedirs(out)
    cmd = [REF, "-i", inp, "-o", out, "-W", "4000", "-C", "1000000", "-n", "2", "-t", "1e-12", "-l", "Err,Dup,Hap,Col", "-@", "2"]
    bin_file = None
    if bins:  # --binArrayFile: overlapping size bins, one of them open-ended
        bin_file = str(tmp_path / "bins.tsv")
        open(bin_file, "w").write("#start\tend\tname\n0\t8000\t0-8Kb\n8000\t40000\t8-40Kb\n20000\t1e9\t20Kb<\n0\t1e9\tALL\n")
        cmd += ["-a", bin_file]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    wl, _ = binfmt.read_cov_native(inp, 1_000_000, 4000)
    labels = _labels_from_bed(os.path.join(out, "final_flagger_prediction.bed"), wl)
    mine = str(tmp_path / "mine.tsv")
    binfmt.write_summary_native(inp, mine, prediction=labels, chunk_len=1_000_000, window_len=4000, bin_array_file=bin_file)
    names = ["prediction_summary_final.tsv"]
    if with_truth:
        names += ["prediction_summary_final.benchmarking.tsv", "prediction_summary_final.benchmarking.auN_ratio.tsv"]
    for name in names:
        want = open(os.path.join(out, name)).read()
        got = open(str(tmp_path / name.replace("prediction_summary_final", "mine"))).read()
        assert len(want.splitlines()) > (30 if name.endswith("final.tsv") else 10), name
        assert got == want, name
    main = open(mine).read()
    assert ("truth_based_auN" in main) == with_truth and ("TRUTH_VS_PREDICTION\t" in main) == with_truth


@needs_ref
def test_benchmark_scores_equal_the_numbers_the_tuning_driver_reads(tmp_path):
    """hfg_benchmark_scores (no files written) against the three numbers programs/src/tune_alpha_hmm_flagger.py:82-111
    parses out of the reference run's *.benchmarking*.tsv files."""
    import pandas as pd
    inp = str(tmp_path / "in.cov")
    binfmt.write_random_rle_cov(inp, [9000, 310_000, 1_250_000, 123_457], seed=13, n_regions=3, with_truth=True)
    out = str(tmp_path / "ref")
    os.makedirs(out)
    cmd = [REF, "-i", inp, "-o", out, "-W", "4000", "-C", "1000000", "-n", "2", "-t", "1e-12", "-l", "Err,Dup,Hap,Col", "-@", "2"]
    assert subprocess.run(cmd, capture_output=True, text=True, timeout=600).returncode == 0
    cov = binfmt.NativeCov(inp, 1_000_000, 4000)
    labels = _labels_from_bed(os.path.join(out, "final_flagger_prediction.bed"), cov.workload)
    got = cov.benchmark_scores(labels)
    t = pd.read_csv(os.path.join(out, "prediction_summary_final.benchmarking.tsv"), sep="\t").rename(columns={"#Metric_Type": "Metric_Type"})
    want = []
    for metric in ("overlap_based", "base_level"):
        row = t[(t.Metric_Type == metric) & (t.Category_Name == "whole_genome") & (t.Size_Bin_Name == "ALL_SIZES") &
                (t.Label == "HARMONIC_MEAN_NO_HAP")]
        want.append(float(row["F1-Score"].item()))
    a = pd.read_csv(os.path.join(out, "prediction_summary_final.benchmarking.auN_ratio.tsv"), sep="\t")
    row = a[(a.Category_Name == "whole_genome") & (a.Size_Bin_Name == "ALL_SIZES") & (a.Label == "HARMONIC_MEAN")]
    want.append(100 * float(row["auN_Ratio"].item()))
    assert all(abs(g - w) < 1e-9 for g, w in zip(got, want)), (got, want)
    cov.close()
