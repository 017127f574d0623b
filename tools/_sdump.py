import sys, os
import numpy as np
sys.path.insert(0, ".")
from flagger_b200 import api, synth, _abi
wl = synth.config2()
K = api.best_num_collapsed_comps(int(wl.cov.max()), wl.region_coverages)
cfg = _abi.make_config(n_regions=1, n_col_comps=K)
p = api.model_init(cfg, wl.region_coverages, wl.window_len)
g = api.HmmFlaggerGPU(cfg, wl)
for i in range(4):
    g.em_iteration(synth.HIFI_ALPHA, p, want_labels=False)
c = g.debug_phase_clocks()[:-1]
np.set_printoptions(linewidth=250)
print("S total (slot 5 -> 6)      ", ((c[:, 6] - c[:, 5]) // 1000).tolist())
print("t0: last fold done (5->12) ", ((c[:, 12] - c[:, 5]) // 1000).tolist())
print("t0: loop done (12->13)     ", ((c[:, 13] - c[:, 12]) // 1000).tolist())
print("t0: 13 -> 6                ", ((c[:, 6] - c[:, 13]) // 1000).tolist())
