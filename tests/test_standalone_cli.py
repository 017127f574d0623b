"""The stand-alone `hmm_flagger_b200` binary (flagger_b200/csrc/hmm_flagger_b200.c: libhfg + the readers of hfg_io.h + its
own writers) against the unmodified reference binary on the same input file and flags.

CPU part: everything the binary writes BEFORE it needs the GPU (initial parameter tables, --dumpBin) is byte-identical,
and without a GPU it stops with an error instead of falling back.  GPU part: every output file of a full run."""
import os
import subprocess

import numpy as np
import pytest
import torch

from flagger_b200 import binfmt, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "hmm_flagger_ref")
CLI = os.path.join(ROOT, "flagger_b200", "hmm_flagger_b200")

needs_ref = pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/hmm_flagger_ref was not built (reference tree "
                                                               "not mounted at build time)")


def _run(binary, inp, out, extra=(), check=True):
    os.makedirs(out, exist_ok=True)
    cmd = [binary, "-i", inp, "-o", out, "-W", "4000", "-C", "1000000", "-n", "6", "-t", "1e-12", "-l", "Err,Dup,Hap,Col", *extra]
    if binary == REF:
        cmd += ["-@", "4"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    if check:
        assert r.returncode == 0, r.stderr[-2000:]
    return r


def _read(path):
    with open(path, "rb") as f:
        return f.read()


def _table(path):
    vals = []
    for line in open(path):
        if line.startswith(("#", "track")):
            continue
        for tok in line.rstrip("\n").split("\t"):
            for v in tok.split(","):
                try:
                    vals.append(float(v))
                except ValueError:
                    pass
    return np.array(vals)


def _inputs(tmp_path, kind, n_regions=3, seed=61):
    wl = synth.small_mixed(n_regions=n_regions, seed=seed)
    inp = str(tmp_path / f"in.{kind}")
    (binfmt.write_bin if kind == "bin" else binfmt.write_cov)(wl, inp)
    alpha = str(tmp_path / "alpha.tsv")
    binfmt.write_alpha_tsv(synth.HIFI_ALPHA, alpha)
    return inp, alpha


@needs_ref
@pytest.mark.parametrize("kind,extra", [("cov.gz", ()), ("bin", ()), ("cov", ("-m", "gaussian")), ("cov", ("-p", "3")),
                                        ("cov", ("-m", "negative_binomial"))])  # (stops at hfg_create unless opted in)
def test_pre_gpu_outputs_byte_identical(tmp_path, kind, extra):
    inp, alpha = _inputs(tmp_path, kind)
    ref_out, cli_out = str(tmp_path / "ref"), str(tmp_path / "cli")
    _run(REF, inp, ref_out, extra=("-A", alpha, "-B", "-n", "1", *extra))
    r = _run(CLI, inp, cli_out, extra=("-A", alpha, "-B", "-n", "1", *extra), check=False)
    if not torch.cuda.is_available():
        # no CPU fallback: the run must stop at the GPU context, loudly
        assert r.returncode != 0 and "Error" in r.stderr
        assert not os.path.exists(os.path.join(cli_out, "final_flagger_prediction.bed"))
    for name in ("transition_initial.tsv", "emission_initial.tsv", "chunks.c_1000000.w_4000.bin"):
        assert _read(os.path.join(ref_out, name)) == _read(os.path.join(cli_out, name)), name


def test_rejects_bad_arguments(tmp_path):
    inp, alpha = _inputs(tmp_path, "cov")
    for extra in (("-t", "0"), ("-x", "nanopore"), ("-M", "1,2"), ("-f", "1.5"), ("-D", "0.1")):  # -D: refused, not ignored
        r = _run(CLI, inp, str(tmp_path / "o"), extra=extra, check=False)
        assert r.returncode != 0 and "Error" in r.stderr, extra
    r = subprocess.run([CLI, "-i", str(tmp_path / "x.txt"), "-o", str(tmp_path)], capture_output=True, text=True)
    assert r.returncode != 0 and "cov/cov.gz" in r.stderr
    r = subprocess.run([CLI, "-i", inp, "-o", str(tmp_path / "missing_dir")], capture_output=True, text=True)
    assert r.returncode != 0 and "does not exist" in r.stderr


def _compare_runs(ref_out, cli_out, posterior=False):
    assert _read(os.path.join(ref_out, "final_flagger_prediction.bed")) == \
        _read(os.path.join(cli_out, "final_flagger_prediction.bed"))
    a, b = _table(os.path.join(ref_out, "loglikelihood.tsv")), _table(os.path.join(cli_out, "loglikelihood.tsv"))
    assert a.shape == b.shape and np.all(np.abs(a - b) <= 2e-4)
    names = [n for n in sorted(os.listdir(ref_out)) if n.startswith(("emission_", "transition_"))]
    assert names == [n for n in sorted(os.listdir(cli_out)) if n.startswith(("emission_", "transition_"))]
    for name in names:
        # same header / row labels, values printed with %.5e
        la, lb = open(os.path.join(ref_out, name)).read().split("\n"), open(os.path.join(cli_out, name)).read().split("\n")
        assert len(la) == len(lb) and la[0] == lb[0], name
        a, b = _table(os.path.join(ref_out, name)), _table(os.path.join(cli_out, name))
        assert a.shape == b.shape and np.allclose(a, b, rtol=2e-5, atol=1e-12), name
    # summary tables on the flat labels: byte-identical (these inputs carry no truth labels, so the reference writes no
    # truth_based_auN rows and the files hold exactly what hfg_write_summary_tsv covers)
    summaries = [n for n in sorted(os.listdir(ref_out)) if n.startswith("prediction_summary_")]
    assert "prediction_summary_initial.tsv" in summaries and "prediction_summary_final.tsv" in summaries
    for name in summaries:
        assert _read(os.path.join(ref_out, name)) == _read(os.path.join(cli_out, name)), name
    if posterior:
        name = "posterior_prediction_final.bed"
        ra, rb = open(os.path.join(ref_out, name)).readlines(), open(os.path.join(cli_out, name)).readlines()
        assert len(ra) == len(rb) and ra[0] == rb[0]
        assert [ln.split("\t")[:3] + [ln.split("\t")[-1]] for ln in ra] == [ln.split("\t")[:3] + [ln.split("\t")[-1]] for ln in rb]
        a, b = _table(os.path.join(ref_out, name)), _table(os.path.join(cli_out, name))
        assert a.shape == b.shape and np.all(np.abs(a - b) <= 0.0101)


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("kind,extra", [
    ("cov.gz", ("-w", "-P")),
    ("bin", ("-w",)),
    ("cov", ("-M", "20000,20000,20000")),          # short Err/Dup/Col blocks are relabelled Hap and merged
    ("cov", ("-m", "gaussian", "-p", "3")),
    ("cov", ("-e",)),                                # --disableAdjustContigEnds
    ("cov", ("-q", "0.1", "--minHighMapqRatio", "0.9", "-f", "0.5")),
    ("bin", ("-s", "-w")),                           # --accelerate (SQUAREM)
    ("bin", ("-k",)),                                # --writeBenchmarkingStatsPerIteration
    ("cov", ("-k", "-s")),
])
def test_full_run_matches_reference(tmp_path, kind, extra):
    inp, alpha = _inputs(tmp_path, kind)
    ref_out, cli_out = str(tmp_path / "ref"), str(tmp_path / "cli")
    _run(REF, inp, ref_out, extra=("-A", alpha, *extra))
    _run(CLI, inp, cli_out, extra=("-A", alpha, *extra))
    _compare_runs(ref_out, cli_out, posterior="-P" in extra)


@pytest.mark.gpu
@needs_ref
def test_preset_without_alpha_and_contig_subset(tmp_path):
    """A preset without --alphaTsv runs with alpha == 0 (the reference's preset tables are `int`, src/hmm_flagger.c:21-58);
    --contigsList keeps only the listed contigs."""
    inp, _ = _inputs(tmp_path, "cov", n_regions=1, seed=62)
    contigs = sorted({ln.split()[0][1:] for ln in open(inp) if ln.startswith(">")})
    assert len(contigs) >= 2
    lst = str(tmp_path / "contigs.txt")
    open(lst, "w").write("\n".join(contigs[1:]) + "\n")
    for extra in (("-x", "ont-r10", "-W", "4000"), ("-c", lst)):
        ref_out, cli_out = str(tmp_path / ("ref" + extra[0])), str(tmp_path / ("cli" + extra[0]))
        _run(REF, inp, ref_out, extra=extra)
        _run(CLI, inp, cli_out, extra=extra)
        _compare_runs(ref_out, cli_out)


@pytest.mark.gpu
@needs_ref
def test_truth_labelled_input_writes_the_benchmarking_files(tmp_path):
    """An input with truth labels and --binArrayFile: besides the prediction summaries the reference writes confusion
    tables, the truth_based_auN metric and the .benchmarking.tsv / .benchmarking.auN_ratio.tsv files; the stand-alone
    binary must write the same bytes (its labels come from the GPU, the tables from hfg_write_summary_tsv)."""
    inp = str(tmp_path / "truth.cov.gz")
    binfmt.write_random_rle_cov(inp, [4000, 9000, 310_000, 1_250_000, 123_457], seed=11, n_regions=3, with_truth=True)
    bins = str(tmp_path / "bins.tsv")
    open(bins, "w").write("#start\tend\tname\n0\t20000\t0-20Kb\n20000\t1e9\t20Kb<\n0\t1e9\tALL\n")
    ref_out, cli_out = str(tmp_path / "ref"), str(tmp_path / "cli")
    _run(REF, inp, ref_out, extra=("-a", bins, "-k"))
    _run(CLI, inp, cli_out, extra=("-a", bins, "-k"))
    _compare_runs(ref_out, cli_out)
    names = [n for n in sorted(os.listdir(ref_out)) if n.startswith("prediction_summary_")]
    assert any(n.endswith(".benchmarking.tsv") for n in names) and any(n.endswith(".auN_ratio.tsv") for n in names)
    assert "truth_based_auN" in open(os.path.join(cli_out, "prediction_summary_final.tsv")).read()
