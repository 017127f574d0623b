"""Load a golden fixture (tests/golden/*.npz, made by tests/golden/make_golden.py from the unmodified reference)."""
import glob
import os

import numpy as np

from flagger_b200 import synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))
               if not p.endswith((".golden.npz", ".squarem.npz", ".nb.npz")))  # .cov reader / SQUAREM / NB fixtures have their own tests
NB_NAMES = sorted(os.path.basename(p)[:-len(".nb.npz")] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.nb.npz"))
                  if not p.endswith("digamma.nb.npz"))  # negative-binomial model: oracle only (tests/test_oracle_nb.py)


def load_squarem(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".squarem.npz"))
    return {k: z[k] for k in z.files}


def load(name, suffix=".npz"):
    z = np.load(os.path.join(GOLDEN_DIR, name + suffix))
    g = {k: z[k] for k in z.files}
    wl = synth.Workload(name, int(g["window_len"]), 0, int(g["cfg"]["mean_read_length"][0]), g["region_coverages"],
                        ["ctg"] * len(g["chunks"]), g["chunks"], g["cov"], g["cov_high_mapq"], g["cov_high_clip"],
                        g["region"], np.full(len(g["cov"]), -1, np.int8))
    return g, wl
