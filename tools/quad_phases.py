"""Phase timeline of the quad E-step kernel (hfg_debug_phase_clocks) on a GPU box: python tools/quad_phases.py [cfg2|cfg3|cfg4|medium]
Per-CTA slots: 0 start, 1 key table done everywhere (barrier 1 released), 2 arrival at barrier 2 (segment products + scans),
3 release, 4 thread 0's end of C1, 5 barrier 3 released (C1 + C2 done everywhere), 6 statistics done (partials written).
Tail row (the last CTA to arrive): 0 start, 1 totals in shared memory, 2 statistics block written, 3/4 exchange, 5 M-step done, 6 end."""
import sys
import numpy as np
sys.path.insert(0, ".")
from flagger_b200 import api, synth, _abi
which = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
if which.startswith("cfg2:"):
    bp = int(float(which.split(":")[1]))
    which = "cfg2"
    synth_cfg2 = lambda: synth.config2(total_bp=bp)
else:
    synth_cfg2 = synth.config2
wl = {"cfg2": synth_cfg2, "cfg3": synth.config3, "cfg4": synth.config4,
      "medium": lambda: synth.config2(total_bp=300_000_000, seed=22), "small": lambda: synth.small_mixed(n_regions=1, seed=12)}[which]()
K = api.best_num_collapsed_comps(int(wl.cov.max()), wl.region_coverages)
cfg = _abi.make_config(n_regions=len(wl.region_coverages), n_col_comps=K)
p = api.model_init(cfg, wl.region_coverages, wl.window_len)
g = api.HmmFlaggerGPU(cfg, wl, timing=True)
ms = []
for i in range(6):
    g.em_iteration(synth.HIFI_ALPHA, p, want_labels=False)
    ms.append(g.last_estep_kernel_ms())
c = g.debug_phase_clocks()
tail, c = c[-1], c[:-1]
d = np.diff(c[:, :7], axis=1)
print(which, "grid", len(c), "kernel_ms", [round(m, 4) for m in ms])
names = ["T+bar1", "A+B", "wait2", "walk+C1(t0)", "C2+bar3", "S"]
for i in range(6):
    print(f"{names[i]:12s} mean {d[:, i].mean():9.0f} p10 {np.percentile(d[:, i], 10):9.0f} p50 {np.percentile(d[:, i], 50):9.0f}"
          f" p90 {np.percentile(d[:, i], 90):9.0f} max {d[:, i].max():9.0f}")
if c[:, 7].max() > 0:  # third-generation kernel: thread 0's own clocks inside the phases
    ex = [("hot table fill", 1, 7), ("A loop (t0)", 7, 8), ("warp scans + stash (t0)", 8, 9), ("level 2 (t0)", 9, 2),
          ("walk + messages (t0)", 3, 10), ("C1 loop (t0)", 10, 4), ("C2 loop (t0)", 4, 11), ("labels out + barrier 3", 11, 5),
          ("S: records + fold (t0)", 5, 12), ("S: finalize (t0)", 12, 13), ("S: wait for the CTA", 13, 14), ("S: partials", 14, 6)]
    for name, a, b in ex:
        dd = c[:, b] - c[:, a]
        print(f"  {name:26s} mean {dd.mean():9.0f} p10 {np.percentile(dd, 10):9.0f} p90 {np.percentile(dd, 90):9.0f}")
print("tail (CTA %d): wait for the last arrival -> totals %d -> statistics block %d -> end %d cycles"
      % (tail[7], tail[1] - tail[0], tail[2] - tail[1], tail[6] - tail[2]))
g.em_begin(synth.HIFI_ALPHA, p, tol=1e-12, max_esteps=4)
for i in range(4):
    g.em_enqueue()
g.em_finish(want_labels=False)
c = g.debug_phase_clocks()
tail = c[-1]
print("device EM iteration ms", [round(g.em_enqueued_ms(i), 4) for i in range(4)], "; tail: totals", int(tail[1] - tail[0]),
      "statistics block", int(tail[2] - tail[1]), "M-step", int(tail[5] - tail[2]), "end", int(tail[6] - tail[5]), "cycles;",
      "M-step: copy-in", int(tail[8] - tail[2]), "updates (thread 0)", int(tail[10] - tail[8]),
      "rest", int(tail[5] - tail[10]), "; rate fit rounds: tree", int(tail[11]) & 255, "predicted", (int(tail[11]) >> 8) & 255, "serial steps", int(tail[11]) >> 16)
