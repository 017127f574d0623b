import sys, os
import numpy as np
sys.path.insert(0, ".")
os.environ["HFG_DBG"] = os.environ.get("DBG", "8")
from flagger_b200 import api, synth, _abi
wl = synth.config2()
K = api.best_num_collapsed_comps(int(wl.cov.max()), wl.region_coverages)
cfg = _abi.make_config(n_regions=1, n_col_comps=K)
p = api.model_init(cfg, wl.region_coverages, wl.window_len)
g = api.HmmFlaggerGPU(cfg, wl)
for i in range(4):
    g.em_iteration(synth.HIFI_ALPHA, p, want_labels=False)
c = g.debug_phase_clocks()[:-1]
nw = int(os.environ.get("HFG_THREADS", "640")) // 32
print("CTA 0, per warp (cycles after barrier 3): loop entry, records loaded, fold done, loop left")
for w in range(nw):
    print(w, c[100 + w, :4].tolist())
