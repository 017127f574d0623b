"""Per-CTA phase timeline of the E-step kernel on the cfg2 workload (hfg_debug_phase_clocks). Run on the GPU box."""
import numpy as np, sys
sys.path.insert(0,".")
from flagger_b200 import api, synth, _abi
wl=synth.config2(); K=api.best_num_collapsed_comps(int(wl.cov.max()), wl.region_coverages)
cfg=_abi.make_config(n_col_comps=K); p=api.model_init(cfg, wl.region_coverages, wl.window_len)
g=api.HmmFlaggerGPU(cfg, wl)
for i in range(3): g.em_iteration(synth.HIFI_ALPHA, p, want_labels=False)
c=g.debug_phase_clocks(); d=np.diff(c[:,:8],axis=1)
print("grid",len(c),"kernel_ms",g.last_estep_kernel_ms())
names=["A","B","wait1","C1(t0)","C2","Dreduce","wait2"]
for i in range(7): print(f"{names[i]:8s} mean {d[:,i].mean():9.0f} p10 {np.percentile(d[:,i],10):9.0f} p50 {np.percentile(d[:,i],50):9.0f} p90 {np.percentile(d[:,i],90):9.0f} max {d[:,i].max():9.0f}")
sm=c[:,8]; 
import collections
cnt=collections.Counter(sm.tolist()); print("blocks per SM:", collections.Counter(cnt.values()))
# A time vs position
print("A by block idx (every 37):", d[::37,0].tolist())
for ph in (0, 4):
    order = np.argsort(-d[:, ph])[:8]
    print(names[ph], "slowest blocks:", [(int(b), int(d[b, ph]), int(sm[b])) for b in order])
# start skew and absolute end of A relative to the earliest start (per-SM clocks are not synchronised; indicative only)
lay_seg = None
