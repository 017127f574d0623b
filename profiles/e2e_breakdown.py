"""Where the end-to-end time of the blocking C-ABI goes on the cfg2 workload: context creation, hfg_set_chunks, and the
per-iteration call.  Run on the GPU box."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
from flagger_b200 import _abi, api, synth
wl = synth.config2()
K = api.best_num_collapsed_comps(int(wl.cov.max()), wl.region_coverages)
cfg = _abi.make_config(n_col_comps=K)
p = api.model_init(cfg, wl.region_coverages, wl.window_len)
g0 = api.HmmFlaggerGPU(cfg, wl); g0.em_iteration(synth.HIFI_ALPHA, p)   # warm the CUDA context / module load
for rep in range(3):
    t0 = time.perf_counter(); g = api.HmmFlaggerGPU(cfg); t1 = time.perf_counter()
    g.set_chunks(wl); t2 = time.perf_counter()
    stats = np.zeros(1, dtype=_abi.region_stats_dtype); labels = np.empty(wl.n_windows, np.int8)
    ts = []
    pp = p
    for i in range(20):
        a = time.perf_counter(); s, ll, _ = g.em_iteration(synth.HIFI_ALPHA, pp, stats=stats, labels=labels); b = time.perf_counter()
        pp, _ = api.mstep(cfg, pp, s, tol=1e-12); c = time.perf_counter()
        ts.append((b - a, c - b, g.last_estep_kernel_ms()))
    ts = np.array(ts[3:])
    a = time.perf_counter(); s, ll, _ = g.em_iteration(synth.HIFI_ALPHA, pp, want_labels=False); b = time.perf_counter()
    print(f"create {1e3*(t1-t0):.2f} ms  set_chunks {1e3*(t2-t1):.2f} ms  em_iteration(with labels) {1e3*ts[:,0].mean():.3f} ms "
          f"(kernel {ts[:,2].mean():.3f} ms)  mstep(py) {1e3*ts[:,1].mean():.3f} ms  em_iteration(no labels) {1e3*(b-a):.3f} ms")
    g.close()
