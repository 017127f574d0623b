"""Where a device-resident EM iteration spends its time: plain E-step call vs queued E+M iterations vs queued final
passes (no M-step), with and without the L2 flush in between.  Run on the GPU box."""
import sys
sys.path.insert(0, ".")
import numpy as np
from flagger_b200 import api, synth, _abi
wl = synth.config2(); K = api.best_num_collapsed_comps(int(wl.cov.max()), wl.region_coverages)
cfg = _abi.make_config(n_col_comps=K); p = api.model_init(cfg, wl.region_coverages, wl.window_len)
g = api.HmmFlaggerGPU(cfg, wl, timing=True)
ms = []
for i in range(6):
    g.em_iteration(synth.HIFI_ALPHA, p, want_labels=False); ms.append(g.last_estep_kernel_ms())
print("plain hfg_em_iteration kernel ms", np.round(ms, 4))
for flush in (False, True):
    for final in (False, True):
        g.em_begin(synth.HIFI_ALPHA, p, tol=1e-12, max_esteps=12)
        for i in range(12):
            if flush: g.l2_flush()
            g.em_enqueue(final_pass=final)
        g.em_finish(want_labels=False)
        print(f"queued x12 flush={flush} final_pass(no M-step)={final}:", np.round([g.em_enqueued_ms(i) for i in range(12)], 4))
c = g.debug_phase_clocks()
print("block 0 (cycles): barrier-4 release -> totals written", int(c[0, 9] - c[0, 7]), "(final pass, no M-step)")
g.em_begin(synth.HIFI_ALPHA, p, tol=1e-12, max_esteps=2); g.em_enqueue(); g.em_enqueue(); g.em_finish(want_labels=False)
c = g.debug_phase_clocks()
print("block 0 (cycles): barrier-4 release -> totals written", int(c[0, 9] - c[0, 7]), "; M-step", int(c[0, 10] - c[0, 9]),
      "; kernel start -> barrier-4 release", int(c[0, 7] - c[0, 0]))
