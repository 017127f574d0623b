import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from flagger_b200 import _abi, api, synth
wl = synth.config2(); K = api.best_num_collapsed_comps(int(wl.cov.max()), wl.region_coverages)
cfg = _abi.make_config(n_col_comps=K); p = api.model_init(cfg, wl.region_coverages, wl.window_len)
torch.cuda.set_device(0)
g0 = api.HmmFlaggerGPU(cfg, wl)
for i in range(3): g0.em_iteration(synth.HIFI_ALPHA, p, want_labels=False)
t0=time.perf_counter(); g = api.HmmFlaggerGPU(cfg); g.set_chunks(wl); t1=time.perf_counter()
stats = np.zeros(1, dtype=_abi.region_stats_dtype); labels = np.empty(wl.n_windows, np.int8)
ts=[]
for i in range(10):
    a=time.perf_counter(); g.em_iteration(synth.HIFI_ALPHA, p, stats=stats, labels=labels); ts.append(time.perf_counter()-a)
print("setup", t1-t0, "iters", [round(x*1e3,3) for x in ts])
