#!/usr/bin/env python
"""Golden vectors for the `.cov` window builder: small synthetic run-length files (committed next to this script) parsed by
the UNMODIFIED reference (ChunksCreator_constructFromCov + ChunksCreator_parseChunks, through oracle/_ref).  Run in the
build container:  python tests/golden/make_golden_cov.py"""
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from flagger_b200 import binfmt  # noqa: E402
import oracle_lib  # noqa: E402

CASES = [
    # file name, contig lengths, chunk_len, window_len, kwargs
    ("cov_rle_a.cov.gz", [53_017, 4_000, 3_999, 12_001, 1], 20_000, 4_000, dict(seed=1, n_regions=3, with_truth=True)),
    ("cov_rle_b.cov", [30_011, 8_200], 7_000, 1_000, dict(seed=2, n_regions=2, with_truth=False)),
    ("cov_rle_float.cov.gz", [20_003], 20_000, 700, dict(seed=3, n_regions=1, with_truth=True, float_values=True)),
]


ODD = ("cov_odd_tokens.cov", [5000, 777, 12001], 3000, 64, {})


def write_odd_token_cov(path, lens, seed=5):
    """Number tokens the fast scanner of hfg_cov_reader.c does not take (leading zeros, decimals, exponents, values above
    the 250 clip), CR LF line ends on one contig, labels of -1, several annotations per block, no final newline."""
    rng = np.random.default_rng(seed)
    out = ["#annotation:len:4\n#annotation:name:0:no_annotation\n#annotation:name:1:whole_genome\n#annotation:name:2:a\n"
           "#annotation:name:3:b\n#region:len:3\n#region:coverage:0:40\n#region:coverage:1:50\n#region:coverage:2:30\n"
           "#label:len:4\n#label:name:0:Err\n#label:name:1:Dup\n#label:name:2:Hap\n#label:name:3:Col\n#truth:true\n"
           "#prediction:true\n#avg_alignment_len:15000\n#start-only:false\n"]
    for ci, L in enumerate(lens):
        eol = "\r\n" if ci == 1 else "\n"
        out.append(f">c{ci} {L}{eol}")
        pos = 1
        while pos <= L:
            ln = int(min(rng.integers(1, 400), L - pos + 1))
            cov = ["12", "007", "12.50", "1e1", "300", "0", "2.25", "40"][int(rng.integers(0, 8))]
            mq = ["3", "3.0", "0.75", "12", "0", "250", "251", "1"][int(rng.integers(0, 8))]
            cl = ["0", "1", "2.5", "0.0"][int(rng.integers(0, 4))]
            ann = ["1", "1,2", "1,2,3", "2", "1,3"][int(rng.integers(0, 5))]
            out.append(f"{pos}\t{pos + ln - 1}\t{cov}\t{mq}\t{cl}\t{ann}\t{int(rng.integers(0, 3))}\t"
                       f"{int(rng.integers(-1, 4))}\t{int(rng.integers(-1, 4))}{eol}")
            pos += ln
    with open(path, "w", newline="") as f:
        f.write("".join(out).rstrip("\n"))


def main():
    for name, lens, chunk_len, window_len, kw in CASES + [ODD]:
        path = os.path.join(HERE, name)
        if name == ODD[0]:
            write_odd_token_cov(path, lens)
        else:
            binfmt.write_random_rle_cov(path, lens, **kw)
        with tempfile.TemporaryDirectory() as tmp:  # the reference writes <input>.index next to the input
            work = os.path.join(tmp, name)
            shutil.copy(path, work)
            r = oracle_lib.reference_parse_cov(work, chunk_len, window_len)
        if r is None:
            raise SystemExit("oracle/_ref/libref_harness.so is missing")
        np.savez_compressed(os.path.join(HERE, name + ".golden.npz"), chunk_len=chunk_len, window_len=window_len,
                            chunks=r["chunks"], names=np.array(r["names"]), cov=r["cov"], mapq=r["mapq"], clip=r["clip"],
                            flags=r["flags"], truth=r["truth"], prediction=r["prediction"],
                            region_coverages=r["region_coverages"], header=r["header"])
        print(name, len(r["chunks"]), "chunks", len(r["cov"]), "windows")


if __name__ == "__main__":
    main()
