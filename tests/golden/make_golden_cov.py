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


def main():
    for name, lens, chunk_len, window_len, kw in CASES:
        path = os.path.join(HERE, name)
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
