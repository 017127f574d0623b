#!/usr/bin/env python
"""Host-side timing of the formats either side of the hot path, CPU only:

    python profiles/host_timing.py [--mbp 300] > profiles/host_timing_<round>.txt

  * `.cov` / `.cov.gz` -> windows: hfg_read_cov on a synthetic run-length file with bam2cov-like block lengths (mean 200 bp,
    ~5 blocks per kbp), against the unmodified reference's chunk builder on a 10x smaller file when oracle/_ref is built;
  * `.bin` -> windows;
  * prediction summary tables (hfg_write_summary_tsv) and the alpha-tuning scores (hfg_benchmark_scores) at 750 024 windows,
    one region / seven regions + eight annotations, with and without truth labels.
Nothing here touches the GPU; the E-step numbers are bench.py's.  Like bench.py's cpu_baseline leg, this measuring script
times the unmodified reference (oracle/_ref, through tests/oracle_lib.py) BESIDE the product; the product itself
(flagger_b200/) never loads anything from oracle/."""
import argparse
import os
import shutil
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from flagger_b200 import binfmt, synth  # noqa: E402


def write_rle_cov(path, contig_lens, mean_block=200, seed=0):
    rng = np.random.default_rng(seed)
    with open(path, "w") as f:
        f.write("#annotation:len:2\n#annotation:name:0:no_annotation\n#annotation:name:1:whole_genome\n#region:len:1\n"
                "#region:coverage:0:40\n#label:len:0\n#truth:false\n#prediction:false\n#avg_alignment_len:15000\n#start-only:false\n")
        for ci, L in enumerate(contig_lens):
            f.write(f">ctg{ci + 1} {L}\n")
            lens = rng.geometric(1.0 / mean_block, size=int(L / mean_block * 1.3) + 10)
            ends = np.cumsum(lens)
            k = int(np.searchsorted(ends, L))
            ends = ends[:k + 1].copy()
            ends[-1] = L
            starts = np.concatenate([[1], ends[:-1] + 1])
            cov = np.clip(40 + np.cumsum(rng.integers(-1, 2, size=len(ends))) % 17 - 8, 0, 300).astype(int)
            cols = [starts.astype(str), ends.astype(str), cov.astype(str), (cov * 0.9).astype(int).astype(str)]
            line = cols[0]
            for c in cols[1:]:
                line = np.char.add(np.char.add(line, "\t"), c)
            f.write("".join(np.char.add(line, "\t0\t1\t0\n").tolist()))


def best(fn, reps=3):
    out = []
    for _ in range(reps):
        t = time.perf_counter()
        r = fn()
        out.append(time.perf_counter() - t)
    return min(out), r


def first_and_best(fn, reps=3):
    """(first call, best of the following ones): the first call of a process pays for fresh pages -- in this microVM a
    first-touch page fault costs ~10 us, so tens of MB of index arrays show up as ~0.1 s -- later calls reuse the heap."""
    t = time.perf_counter()
    fn()
    first = time.perf_counter() - t
    return first, best(fn, reps)[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mbp", type=int, default=300, help="size of the synthetic coverage file in Mbp")
    ap.add_argument("--summary-part", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.summary_part:
        return summary_part()
    tmp = tempfile.mkdtemp(prefix="hfg_host_timing_")
    try:
        total = args.mbp * 1_000_000
        plain = os.path.join(tmp, "c.cov")
        write_rle_cov(plain, [total // 4] * 4)
        import zlib
        with open(plain, "rb") as fi:
            text = fi.read()
        # the reference writes coverage files with gzopen(path, "w6h"): level 6, Huffman-only (ptBlock.c:2271)
        for name, strategy in (("c.huffman.cov.gz", zlib.Z_HUFFMAN_ONLY), ("c.default.cov.gz", zlib.Z_DEFAULT_STRATEGY)):
            co = zlib.compressobj(6, zlib.DEFLATED, 31, 8, strategy)
            with open(os.path.join(tmp, name), "wb") as fo:
                fo.write(co.compress(text) + co.flush())
        n_lines = text.count(b"\n")
        del text
        print(f"coverage file: {args.mbp} Mbp, {n_lines} lines, {os.path.getsize(plain) / 1e6:.1f} MB text, "
              f"{os.path.getsize(os.path.join(tmp, 'c.huffman.cov.gz')) / 1e6:.1f} MB gzip as the reference writes it (Huffman-only), "
              f"{os.path.getsize(os.path.join(tmp, 'c.default.cov.gz')) / 1e6:.1f} MB gzip -6; {os.cpu_count()} host threads")
        for name in ("c.cov", "c.huffman.cov.gz", "c.default.cov.gz"):
            path = os.path.join(tmp, name)
            for zl in ((False, True) if name.endswith(".gz") else (False,)):
                if zl:
                    os.environ["HFG_ZLIB_INFLATE"] = "1"
                dt, (wl, _) = best(lambda: binfmt.read_cov_native(path, 20_000_000, 4000), reps=5)
                os.environ.pop("HFG_ZLIB_INFLATE", None)
                print(f"  hfg_read_cov {name:18s}{' (zlib inflate)' if zl else '               '} {dt:7.3f} s  ({dt / n_lines * 1e9:5.0f} ns/line, "
                      f"{wl.n_windows} windows) -> 3 Gbp: {dt * 3000 / args.mbp:5.1f} s")
        import oracle_lib
        if oracle_lib.reference() is not None:
            small = os.path.join(tmp, "s.cov")
            write_rle_cov(small, [total // 40] * 4)
            t = time.perf_counter()
            oracle_lib.reference_parse_cov(small, 20_000_000, 4000, threads=os.cpu_count())
            dt = time.perf_counter() - t
            print(f"  reference chunk builder, {args.mbp // 10} Mbp in 4 contigs (4 parse jobs): {dt:.2f} s -> 3 Gbp: {dt * 30000 / args.mbp:.0f} s "
                  f"of parse-job time (it parallelises over chunks)")
        # the table writers in a process of their own, so that "first call" means a cold heap
        import subprocess
        subprocess.run([sys.executable, os.path.abspath(__file__), "--summary-part"], check=True)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def summary_part():
    tmp = tempfile.mkdtemp(prefix="hfg_host_timing_")
    try:
        for name, wl in (("cfg2 (1 region, 2 annotations)", synth.config2(seed=1)), ("cfg4 (7 regions, 8 annotations)", synth.config4(seed=1))):
            path = os.path.join(tmp, "w.bin")
            binfmt.write_bin(wl, path, with_truth=True)
            dt, _ = best(lambda: binfmt.read_bin_native(path))
            print(f"{name}: {wl.n_windows} windows; hfg_read_bin {dt * 1e3:.0f} ms")
            pred = wl.truth.copy()
            pred[::97] = (pred[::97] + 1) % 4
            cov = binfmt.NativeCov(path, 20_000_000, 4000)
            read_ms = dt * 1e3
            for use_truth in (False, True):
                first, warm = first_and_best(lambda: binfmt.write_summary_native(path, os.path.join(tmp, "sum.tsv"), prediction=pred,
                                                                                use_truth=use_truth))
                print(f"  summary tables ({'prediction + truth, 3 files' if use_truth else 'prediction only'}): first call "
                      f"{first * 1e3 - read_ms:.0f} ms, then {warm * 1e3 - read_ms:.0f} ms (re-reading the .bin subtracted)")
            first, warm = first_and_best(lambda: cov.benchmark_scores(pred, annotation_label="whole_genome", size_label="ALL_SIZES",
                                                                      overlap_ratio_threshold=0.4, bin_array_file=None))
            sc = cov.benchmark_scores(pred, annotation_label="whole_genome", size_label="ALL_SIZES", overlap_ratio_threshold=0.4,
                                      bin_array_file=None)
            print(f"  hfg_benchmark_scores (alpha tuning, per candidate): first call {first * 1e3:.1f} ms, then {warm * 1e3:.1f} ms -> {sc}")
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
