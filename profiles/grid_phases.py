"""Phase timeline (clock64 per CTA) of the E-step for small shards and different grids. Run on the GPU box."""
import os, sys
sys.path.insert(0, ".")
import numpy as np
from flagger_b200 import _abi, api, synth, dist as hdist
wl_full = synth.config2()
K = api.best_num_collapsed_comps(int(wl_full.cov.max()), wl_full.region_coverages)
cfg = _abi.make_config(n_col_comps=K)
p = api.model_init(cfg, wl_full.region_coverages, wl_full.window_len)
names = ["T+bar1", "A+B", "wait2", "C1(t0)", "C2+bar3", "S+Dred", "wait4"]  # see profiles/phase_timeline.py
for world, wpt in ((8, 1), (8, 2), (16, 1), (4, 1), (4, 4)):
    wl = hdist.shard_chunks(wl_full, 0, world)
    os.environ["HFG_MIN_WPT"] = str(wpt)
    g = api.HmmFlaggerGPU(cfg, wl, timing=True)
    for i in range(5):
        g.em_iteration(synth.HIFI_ALPHA, p, want_labels=False)
    c = g.debug_phase_clocks(); d = np.diff(c[:, :8], axis=1)
    print(f"shard 1/{world} wpt {wpt}: grid {len(c)} kernel {g.last_estep_kernel_ms():.4f} ms; total cycles (max over CTAs of end-start) {int((c[:,7]-c[:,0]).max())}")
    print("   " + "  ".join(f"{names[i]} {d[:, i].mean():.0f}/{d[:, i].max():.0f}" for i in range(7)), flush=True)
    g.close()
