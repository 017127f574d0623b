"""What compute-sanitizer is run on (profiles/sanitizer_r2.txt): every launch type of the library on small workloads."""
import sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from flagger_b200 import _abi, api, synth
for R, seed in ((1, 12), (3, 14)):
    wl = synth.small_mixed(n_regions=R, seed=seed)
    K = api.best_num_collapsed_comps(int(wl.cov.max()), wl.region_coverages)
    cfg = _abi.make_config(n_regions=R, n_col_comps=K)
    p = api.model_init(cfg, wl.region_coverages, wl.window_len)
    g = api.HmmFlaggerGPU(cfg, wl)
    st, ll, lab = g.em_iteration(synth.HIFI_ALPHA, p)
    pe, lle, labe = g.run_em(synth.HIFI_ALPHA, p, 3, tol=1e-12)
    fo = g.forward_only(synth.HIFI_ALPHA, p)
    post = g.posteriors()
    print(R, ll, fo, len(lle), int(labe.sum()), float(post.sum()))
    g.close()
b = api.HmmFlaggerBatch(cfg, wl, n_lanes=3)
ps, lls, labs = b.run_em(np.array([synth.HIFI_ALPHA] * 4), p, 2)
print("batch", [len(x) for x in lls])
b.close()
