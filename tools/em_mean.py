"""Mean / median / minimum of the device-resident EM iteration time over 52 iterations (warm L2 and flushed): the number an A/B of
two builds needs (single runs are quantised to ~2 us by the event timer and differ by +-3 % between boxes).
    python tools/em_mean.py [cfg2|cfg3|cfg4];  same-box A/B of two builds: tools/ab_run.sh (libraries in tools/ab/)"""
import sys
import numpy as np
sys.path.insert(0, ".")
from flagger_b200 import api, synth, _abi
which = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
wl = {"cfg2": synth.config2, "cfg3": synth.config3, "cfg4": synth.config4}[which]()
K = api.best_num_collapsed_comps(int(wl.cov.max()), wl.region_coverages)
cfg = _abi.make_config(n_regions=len(wl.region_coverages), n_col_comps=K)
p = api.model_init(cfg, wl.region_coverages, wl.window_len)
g = api.HmmFlaggerGPU(cfg, wl)
out = []
for flush in (0, 1):
    g.em_begin(synth.HIFI_ALPHA, p, tol=1e-12, max_esteps=64)
    for i in range(60):
        if flush:
            g.l2_flush(256 << 20)
        g.em_enqueue()
    g.em_finish(want_labels=False)
    ms = np.array([g.em_enqueued_ms(i) for i in range(8, 60)])
    out.append(f"{'flushed' if flush else 'warm L2'}: mean {ms.mean():.5f} median {np.median(ms):.5f} min {ms.min():.5f}")
print(which, "; ".join(out))
