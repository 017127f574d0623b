"""Kernel time of the E-step against the grid-sizing rule (HFG_MIN_WPT = minimum windows per thread before another CTA is
added) for the per-GPU shard sizes of a 1/2/4/8-GPU run of cfg2.  Run on the GPU box: python profiles/grid_sweep.py"""
import os, sys
sys.path.insert(0, ".")
import numpy as np
from flagger_b200 import _abi, api, synth, dist as hdist
wl_full = synth.config2()
K = api.best_num_collapsed_comps(int(wl_full.cov.max()), wl_full.region_coverages)
cfg = _abi.make_config(n_col_comps=K)
p = api.model_init(cfg, wl_full.region_coverages, wl_full.window_len)
for world in (1, 2, 4, 8, 16, 64):
    wl = hdist.shard_chunks(wl_full, 0, world)
    row = []
    for wpt in (1, 2, 4):
        os.environ["HFG_MIN_WPT"] = str(wpt)
        g = api.HmmFlaggerGPU(cfg, wl, timing=True)
        ts = []
        for i in range(12):
            g.em_iteration(synth.HIFI_ALPHA, p, want_labels=False)
            ts.append(g.last_estep_kernel_ms())
        grid = len(g.debug_phase_clocks())
        g.close()
        row.append(f"wpt{wpt}: grid {grid:3d} {np.median(ts[3:]):.4f} ms")
    print(f"shard 1/{world}: {wl.n_windows:7d} windows  " + "   ".join(row), flush=True)
