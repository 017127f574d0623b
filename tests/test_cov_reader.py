"""CPU: the `.cov` / `.cov.gz` / `.bin` readers (flagger_b200/csrc/hfg_cov_reader.c, include/hfg_io.h) against golden
vectors produced by the unmodified reference's chunk builder, the reference's own fixtures when its tree is mounted, and
the pure-Python `.bin` round trip."""
import glob
import os
import shutil

import numpy as np
import pytest

import oracle_lib
from flagger_b200 import binfmt, synth

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = sorted(glob.glob(os.path.join(HERE, "golden", "*.golden.npz")))
REF_FIXTURES = "/root/reference/programs/tests/test_files/chunks_creator"


def _same(wl, hdr, r):
    assert np.array_equal(wl.chunks, r["chunks"])
    assert wl.contig_names == list(r["names"])
    assert np.array_equal(wl.cov, r["cov"]) and np.array_equal(wl.cov_high_mapq, r["mapq"])
    assert np.array_equal(wl.cov_high_clip, r["clip"])
    assert np.array_equal(hdr["annotation_flag"], r["flags"])
    assert np.array_equal(wl.region, (r["flags"] >> np.uint64(58)).astype(np.uint8))
    assert np.array_equal(wl.truth, r["truth"]) and np.array_equal(hdr["prediction_labels"], r["prediction"])
    assert np.array_equal(wl.region_coverages, r["region_coverages"])


@pytest.mark.parametrize("golden", GOLDEN, ids=[os.path.basename(g) for g in GOLDEN])
def test_cov_reader_matches_reference_golden(golden):
    g = np.load(golden)
    path = golden[: -len(".golden.npz")]
    wl, hdr = binfmt.read_cov_native(path, int(g["chunk_len"]), int(g["window_len"]))
    _same(wl, hdr, g)
    assert wl.avg_alignment_len == int(g["header"][6]) and hdr["n_labels"] == int(g["header"][2])
    assert hdr["truth"] == bool(g["header"][3]) and hdr["start_only"] == bool(g["header"][5])


def test_cov_reader_on_reference_fixtures(tmp_path):
    """The reference pins its window builder with tests/test_files/chunks_creator/* (tests/test_chunks_creator.c)."""
    if not os.path.isdir(REF_FIXTURES) or oracle_lib.reference() is None:
        pytest.skip("reference tree / oracle/_ref not available on this box")
    for name in ("test_1.cov", "test_1.cov.gz", "test_1_with_labels.cov"):
        for chunk_len, window_len in ((40, 20), (1000, 7), (30, 1)):
            work = str(tmp_path / f"{chunk_len}_{window_len}_{name}")
            shutil.copy(os.path.join(REF_FIXTURES, name), work)
            r = oracle_lib.reference_parse_cov(work, chunk_len, window_len)
            wl, hdr = binfmt.read_cov_native(os.path.join(REF_FIXTURES, name), chunk_len, window_len)
            _same(wl, hdr, r)
    # the truth table hard-coded in the reference's own test (tests/test_chunks_creator.c:12-19) for windowLen=20, chunkLen=40
    wl, _ = binfmt.read_cov_native(os.path.join(REF_FIXTURES, "test_1.cov"), 40, 20)
    assert wl.cov.tolist() == [5, 10, 10, 15, 16, 16, 7]
    assert wl.cov_high_mapq.tolist() == [2, 10, 10, 15, 16, 16, 7]
    assert wl.cov_high_clip.tolist() == [2, 10, 10, 13, 16, 16, 1]
    assert wl.region.tolist() == [0, 1, 1, 1, 1, 1, 0]
    assert [(int(c["s"]), int(c["e"])) for c in wl.chunks] == [(0, 39), (40, 109), (0, 9)]


def test_cov_reader_against_live_reference_random(tmp_path):
    if oracle_lib.reference() is None:
        pytest.skip("oracle/_ref not built")
    for seed in range(4):
        path = str(tmp_path / f"r{seed}.cov.gz")
        binfmt.write_random_rle_cov(path, [17_000 + 997 * seed, 5, 8_001], seed=10 + seed, float_values=seed == 3)
        for chunk_len, window_len in ((5000, 512), (100_000, 4000)):
            r = oracle_lib.reference_parse_cov(path, chunk_len, window_len)
            wl, hdr = binfmt.read_cov_native(path, chunk_len, window_len)
            _same(wl, hdr, r)


def test_window_level_cov_round_trip(tmp_path):
    """write_cov (one block per window) -> native reader reproduces the workload; same through the .bin writer/readers."""
    wl = synth.small_mixed(n_regions=3, seed=12)
    cov_path, bin_path = str(tmp_path / "w.cov.gz"), str(tmp_path / "w.bin")
    binfmt.write_cov(wl, cov_path, with_truth=True)
    binfmt.write_bin(wl, bin_path, with_truth=True)
    a, _ = binfmt.read_cov_native(cov_path, wl.chunk_len, wl.window_len)
    b, hb = binfmt.read_bin_native(bin_path)
    c, _ = binfmt.read_bin(bin_path)
    for other in (a, b, c):
        assert np.array_equal(other.chunks, wl.chunks) and np.array_equal(other.cov, wl.cov)
        assert np.array_equal(other.cov_high_mapq, wl.cov_high_mapq) and np.array_equal(other.region, wl.region)
        assert np.array_equal(other.truth, wl.truth) and np.array_equal(other.region_coverages, wl.region_coverages)
    assert hb["truth"] and b.window_len == wl.window_len and b.chunk_len == wl.chunk_len


def test_cov_reader_rejects_gaps(tmp_path):
    p = tmp_path / "gap.cov"
    p.write_text("#annotation:len:1\n#region:len:1\n#region:coverage:0:40\n>c 100\n1\t10\t4\t4\t0\t0\t0\n21\t100\t4\t4\t0\t0\t0\n")
    with pytest.raises(ValueError, match="tile the contig"):
        binfmt.read_cov_native(str(p), 1000, 10)


def test_one_block_across_several_chunk_boundaries(tmp_path):
    """A deliberate difference from the reference (documented in hfg_cov_reader.c and DESIGN.md section 7): one run-length
    block that crosses two or more chunk boundaries.  ChunksCreator_createCovIndex adds at most one chunk per track line
    (chunk.c:444-547; the assert that would catch it is compiled out), so the reference drops or empties the chunks in
    between: a single 12 345-base block with -C 3000 gives it 1 chunk / 30 windows.  This reader cuts the contig at every
    boundary, the short remainder merged into the last chunk as for any other contig, and every base lands in a window."""
    p = tmp_path / "long_block.cov"
    p.write_text("#annotation:len:1\n#region:len:1\n#region:coverage:0:40\n>c 12345\n1\t12345\t37\t30\t2\t0\t0\n")
    wl, _ = binfmt.read_cov_native(str(p), 3000, 100)
    assert wl.n_chunks == 4 and wl.n_windows == 124
    assert [int(v) for v in wl.chunks["n_windows"]] == [30, 30, 30, 34]
    assert np.all(wl.cov == 37) and np.all(wl.cov_high_mapq == 30) and np.all(wl.cov_high_clip == 2)
    # the same contig as one block per window gives the same windows
    q = tmp_path / "per_window.cov"
    lines = ["#annotation:len:1", "#region:len:1", "#region:coverage:0:40", ">c 12345"]
    lines += [f"{s}\t{min(s + 99, 12345)}\t37\t30\t2\t0\t0" for s in range(1, 12346, 100)]
    q.write_text("\n".join(lines) + "\n")
    w2, _ = binfmt.read_cov_native(str(q), 3000, 100)
    assert np.array_equal(w2.chunks, wl.chunks) and np.array_equal(w2.cov, wl.cov)
