"""Per-CTA phase timeline of the E-step kernel (hfg_debug_phase_clocks). Run on the GPU box:
    python profiles/phase_timeline.py [cfg2|cfg3|cfg4]
Slots: 0 start, 1 key table done everywhere (barrier 1 released), 2 arrival at barrier 2 (segment products + block scans),
3 release, 4 thread 0's end of C1, 5 barrier 3 released (C1 + C2 done everywhere), 6 arrival at barrier 4 (statistics),
7 release."""
import numpy as np, sys
sys.path.insert(0,".")
from flagger_b200 import api, synth, _abi
which = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
wl={"cfg2": synth.config2, "cfg3": synth.config3, "cfg4": synth.config4}[which]()
K=api.best_num_collapsed_comps(int(wl.cov.max()), wl.region_coverages)
cfg=_abi.make_config(n_regions=len(wl.region_coverages), n_col_comps=K); p=api.model_init(cfg, wl.region_coverages, wl.window_len)
g=api.HmmFlaggerGPU(cfg, wl, timing=True)
ms=[]
for i in range(6):
    g.em_iteration(synth.HIFI_ALPHA, p, want_labels=False); ms.append(g.last_estep_kernel_ms())
c=g.debug_phase_clocks(); d=np.diff(c[:,:8],axis=1)
print(which, "grid",len(c),"kernel_ms",[round(m,4) for m in ms])
names=["T+bar1","A+B","wait2","C1(t0)","C2+bar3","S+Dred","wait4"]
for i in range(7): print(f"{names[i]:8s} mean {d[:,i].mean():9.0f} p10 {np.percentile(d[:,i],10):9.0f} p50 {np.percentile(d[:,i],50):9.0f} p90 {np.percentile(d[:,i],90):9.0f} max {d[:,i].max():9.0f}")
print("total cycles (max over CTAs of end-start)", int((c[:,7]-c[:,0]).max()))
print("block 0 tail: barrier-4 release -> grid totals", int(c[0,11]-c[0,7]), "-> statistics block written", int(c[0,9]-c[0,11]), "cycles; kernel", round(ms[-1]*1.965e3), "kcycles; block 0 start->barrier-4 release", int(c[0,7]-c[0,0]))
# device-resident iteration (E-step + M-step in the tail)
g.em_begin(synth.HIFI_ALPHA, p, tol=1e-12, max_esteps=4)
for i in range(4): g.em_enqueue()
g.em_finish(want_labels=False)
c=g.debug_phase_clocks()
print("device EM iteration ms", [round(g.em_enqueued_ms(i),4) for i in range(4)], "; block 0: tail", int(c[0,9]-c[0,7]), "M-step", int(c[0,10]-c[0,9]), "cycles")
