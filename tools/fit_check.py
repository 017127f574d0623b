"""The predicted rounds of the device rate fit (hfg_fit_rate_cta) against the tree-only search (HFG_DBG=16) on a GPU box:
parameters of a 12-iteration device-resident EM run must have the same bits.   python tools/fit_check.py [small|cfg2|cfg4 ...]"""
import os
import sys

import numpy as np

sys.path.insert(0, ".")
from flagger_b200 import _abi, api, synth  # noqa: E402


def make(kind):
    if kind == "small":
        return synth.small_mixed(n_regions=3, seed=14)
    if kind == "medium":
        return synth.config2(total_bp=300_000_000, seed=22)
    return {"cfg2": synth.config2, "cfg3": synth.config3, "cfg4": synth.config4}[kind]()


def run(wl, cfg, params, dbg):
    if dbg:
        os.environ["HFG_DBG"] = str(dbg)
    else:
        os.environ.pop("HFG_DBG", None)
    g = api.HmmFlaggerGPU(cfg, wl)
    g.em_begin(synth.HIFI_ALPHA, params, tol=1e-12, max_esteps=12)
    rounds, cyc = [], []
    for _ in range(12):
        g.em_enqueue()
        import torch
        torch.cuda.synchronize()
        t = g.debug_phase_clocks()[-1]
        rounds.append((int(t[11]) & 255, (int(t[11]) >> 8) & 255, int(t[11]) >> 16))
        cyc.append(dict(fit=int(t[10] - t[8]), start=int(t[8] - t[0]), first=int(t[12] - t[8]), walk=int(t[14] - t[12]), objective=int(t[13] - t[14]), rest=int(t[15] - t[13]), end=int(t[10] - t[15]), last_scout=int(t[9] - t[0])))
    p, ll, _, _ = g.em_finish(want_labels=False)
    g.close()
    return _abi.params_as_flat(p).copy(), np.array(ll), rounds, cyc


ok = True
for kind in sys.argv[1:] or ["small", "cfg2"]:
    wl = make(kind)
    R = len(wl.region_coverages)
    K = api.best_num_collapsed_comps(int(wl.cov.max()), wl.region_coverages)
    cfg = _abi.make_config(n_regions=R, n_col_comps=K)
    params = api.model_init(cfg, wl.region_coverages, wl.window_len)
    pa, la, ra, ca = run(wl, cfg, params, 0)
    pb, lb, rb, cb = run(wl, cfg, params, 16)
    same = np.array_equal(pa.view(np.uint64), pb.view(np.uint64)) and np.array_equal(la.view(np.uint64), lb.view(np.uint64))
    ok &= same
    print(f"{kind}: R={R}; bits identical: {same}; rounds (tree, predicted, serial steps) per iteration, group 0: {ra} vs tree only {rb}")
    print(f"   cycles, last iteration: predicted {ca[-1]}\n      tree only {cb[-1]}")
print("FIT_CHECK", "OK" if ok else "FAILED")
sys.exit(0 if ok else 1)
