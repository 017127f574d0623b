import sys
import numpy as np
sys.path.insert(0, ".")
from flagger_b200 import api, synth, _abi
wl = synth.config2()
K = api.best_num_collapsed_comps(int(wl.cov.max()), wl.region_coverages)
cfg = _abi.make_config(n_regions=1, n_col_comps=K, model_type=_abi.MODEL_NEGATIVE_BINOMIAL, mean_read_length=wl.avg_alignment_len)
p = api.model_init(cfg, wl.region_coverages, wl.window_len)
g = api.HmmFlaggerGPU(cfg, wl)
g.em_begin(np.zeros((4, 4)), p, tol=1e-12, max_esteps=8)
for i in range(6):
    g.em_enqueue()
pp, ll, conv, _ = g.em_finish(want_labels=False)
t = g.debug_phase_clocks()[-1]
print("NB cfg2: device EM iteration ms", [round(g.em_enqueued_ms(i), 4) for i in range(6)], "tail: totals", int(t[1] - t[0]), "stats block", int(t[2] - t[1]),
      "estimators + M-step", int(t[5] - t[2]), "of which: constants + histogram in", int(t[12] - t[2]), "pmf", int(t[13] - t[12]), "estimator sums", int(t[14] - t[13]),
      "weight denominators", int(t[15] - t[14]), "exchange + M-step", int(t[5] - t[15]))
