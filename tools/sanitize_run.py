"""What compute-sanitizer is run on (profiles/sanitizer_r2.txt): every launch type of the library on small workloads (both
blocking paths, the device-resident loop, forward-only, posteriors, the negative-binomial model, batched runs)."""
import sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from flagger_b200 import _abi, api, synth
for R, seed in ((1, 12), (3, 14)):
    wl = synth.small_mixed(n_regions=R, seed=seed)
    K = api.best_num_collapsed_comps(int(wl.cov.max()), wl.region_coverages)
    cfg = _abi.make_config(n_regions=R, n_col_comps=K)
    p = api.model_init(cfg, wl.region_coverages, wl.window_len)
    g = api.HmmFlaggerGPU(cfg, wl)  # one region: the one-launch fast path of the blocking calls; three: the graph path
    st, ll, lab = g.em_iteration(synth.HIFI_ALPHA, p)
    gt = api.HmmFlaggerGPU(cfg, wl, timing=True)  # the graph path with its events, whatever the model
    stt, llt, labt = gt.em_iteration(synth.HIFI_ALPHA, p)
    assert llt == ll and np.array_equal(lab, labt)
    gt.close()
    pe, lle, labe = g.run_em(synth.HIFI_ALPHA, p, 3, tol=1e-12)
    fo = g.forward_only(synth.HIFI_ALPHA, p)
    post = g.posteriors()
    print(R, ll, fo, len(lle), int(labe.sum()), float(post.sum()))
    g.close()
# negative binomial: blocking E-step (table from the host, histogram folded by the grid) and the device-resident loop
for R, seed in ((1, 21), (3, 22)):
    wl_nb = synth.small_mixed(n_regions=R, seed=seed)
    K = api.best_num_collapsed_comps(int(wl_nb.cov.max()), wl_nb.region_coverages)
    cfg_nb = _abi.make_config(n_regions=R, n_col_comps=K, model_type=_abi.MODEL_NEGATIVE_BINOMIAL, mean_read_length=wl_nb.avg_alignment_len)
    p_nb = api.model_init(cfg_nb, wl_nb.region_coverages, wl_nb.window_len)
    g = api.HmmFlaggerGPU(cfg_nb, wl_nb)
    st, ll, lab = g.em_iteration(np.zeros((4, 4)), p_nb)
    pe, lle, labe = g.run_em(np.zeros((4, 4)), p_nb, 3, tol=1e-12)
    print("nb", R, ll, len(lle), int(labe.sum()))
    g.close()
b = api.HmmFlaggerBatch(cfg, wl, n_lanes=3)
ps, lls, labs = b.run_em(np.array([synth.HIFI_ALPHA] * 4), p, 2)
print("batch", [len(x) for x in lls])
b.close()
