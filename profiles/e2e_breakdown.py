"""Per-call cost of the blocking drop-in entry point hfg_em_iteration (host parameters in, host statistics [+ labels]
out) on cfg2: wall time per call, device span of the call (upload, kernel, read-back), kernel time."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
from flagger_b200 import api, synth, _abi
wl = synth.config2(); K = api.best_num_collapsed_comps(int(wl.cov.max()), wl.region_coverages)
cfg = _abi.make_config(n_col_comps=K); p = api.model_init(cfg, wl.region_coverages, wl.window_len)
g = api.HmmFlaggerGPU(cfg, wl, timing=True)
stats = np.zeros(1, dtype=_abi.region_stats_dtype); labels = np.empty(wl.n_windows, np.int8)
for want in (False, True):
    for i in range(5): g.em_iteration(synth.HIFI_ALPHA, p, want_labels=want, stats=stats, labels=labels if want else None)
    wall, span, kern = [], [], []
    for i in range(40):
        t = time.perf_counter(); g.em_iteration(synth.HIFI_ALPHA, p, want_labels=want, stats=stats, labels=labels if want else None)
        wall.append(time.perf_counter() - t); span.append(g.last_call_device_ms()); kern.append(g.last_estep_kernel_ms())
    print(f"labels={want}: wall/call {1e3*np.median(wall):.3f} ms, device span {np.median(span):.3f} ms, kernel {np.median(kern):.3f} ms")
t = time.perf_counter()
for i in range(200): api.mstep(cfg, p, stats)
print(f"api.mstep (python + host C): {1e6*(time.perf_counter()-t)/200:.1f} us per call")
