#!/usr/bin/env python
"""Golden vectors for --accelerate (SQUAREM) from the UNMODIFIED reference (oracle/_ref/libref_harness.so):

    python tests/golden/make_golden_squarem.py

For each base fixture <name>.npz (inputs + initial parameters, see make_golden.py) writes <name>.squarem.npz with
  * the reference's SquareAccelerator candidates for (p0, p1, p2) = three successive EM parameter sets, after 0..5 step
    halvings (parameters, step length, feasibility), the same for two synthetic slowly-contracting triples, and
  * a 3-outer-iteration accelerated EM run (log-likelihood trajectory, accepted step lengths, final parameters, labels)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import golden_util  # noqa: E402
import oracle_lib  # noqa: E402

N_SHRINKS = 6
OUTER = 3


def main():
    ref = oracle_lib.reference(threads=2)
    if ref is None:
        raise SystemExit("oracle/_ref/libref_harness.so is missing: run `make -C oracle` with /root/reference mounted")
    for name in golden_util.NAMES:
        g, wl = golden_util.load(name)
        cfg, alpha, p0 = g["cfg"], g["alpha"], g["params0"]
        p1, _ = ref.mstep(cfg, p0, ref.estep(cfg, wl, alpha, p0)["stats"], tol=1e-3)
        p2, _ = ref.mstep(cfg, p1, ref.estep(cfg, wl, alpha, p1)["stats"], tol=1e-3)
        cands = [ref.squarem(cfg, p0, p1, p2, n) for n in range(N_SHRINKS)]
        # slowly contracting synthetic triples p2 = p1 + c (p1 - p0): step lengths of -1/(1-c), candidates that start
        # infeasible and become feasible while the step is halved towards -1
        syn = {}
        for c in (0.9, 0.5, 0.1):
            q2 = p1.copy()
            q2.view(np.float64)[:] = p1.view(np.float64) + c * (p1.view(np.float64) - p0.view(np.float64))
            cs = [ref.squarem(cfg, p0, p1, q2, n) for n in range(9)]
            tag = str(c).replace(".", "")
            syn[f"syn{tag}_p2"] = q2
            syn[f"syn{tag}_params"] = np.stack([x[0] for x in cs])
            syn[f"syn{tag}_rates"] = np.array([x[1] for x in cs])
            syn[f"syn{tag}_feasible"] = np.array([x[2] for x in cs])
            print("  c =", c, "rates", [round(x[1], 3) for x in cs], "feasible", [int(x[2]) for x in cs])
        acc = ref.run_em_accelerated(cfg, wl, alpha, p0, OUTER, tol=1e-12)
        assert acc["rc"] == 0
        out = os.path.join(HERE, name + ".squarem.npz")
        np.savez_compressed(out, p1=p1, p2=p2, cand_params=np.stack([c[0] for c in cands]),
                            cand_rates=np.array([c[1] for c in cands]), cand_feasible=np.array([c[2] for c in cands]),
                            acc_logliks=acc["logliks"], acc_rates=acc["alpha_rates"], acc_params=acc["params"],
                            acc_labels=acc["labels"], **syn)
        print(name, "rates", acc["alpha_rates"], "cand rates", [round(c[1], 4) for c in cands], os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
