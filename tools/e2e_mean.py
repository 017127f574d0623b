"""The blocking end-to-end job of bench.py's `e2e` leg, alone: hfg_create + hfg_set_chunks + 51 x [hfg_em_iteration with host
buffers + host M-step] driven from C (hfg_debug_blocking_steps), L2 flushed before every step.  Prints set-up and mean step
time for the one-launch fast path (default) and for the graph path with timing events (timing=True): an A/B inside one build.
    python tools/e2e_mean.py [cfg2|cfg3|cfg4] [jobs]"""
import sys
import time
import numpy as np
sys.path.insert(0, ".")
from flagger_b200 import api, synth, _abi
which = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
jobs = int(sys.argv[2]) if len(sys.argv) > 2 else 6
wl = {"cfg2": synth.config2, "cfg3": synth.config3, "cfg4": synth.config4}[which]()
K = api.best_num_collapsed_comps(int(wl.cov.max()), wl.region_coverages)
R = len(wl.region_coverages)
cfg = _abi.make_config(n_regions=R, n_col_comps=K)
p0 = api.model_init(cfg, wl.region_coverages, wl.window_len)
lab = api.PinnedArray(wl.n_windows, np.int8)
ref = None
for timing in (False, True, False, True):
    rows = []
    for rep in range(jobs):
        stats = np.zeros(R, dtype=_abi.region_stats_dtype)
        t0 = time.perf_counter()
        g = api.HmmFlaggerGPU(cfg, timing=timing)
        g.set_chunks(wl)
        t_up = time.perf_counter() - t0
        params, stats, ll, labels, secs = g.blocking_steps(synth.HIFI_ALPHA, p0, 51, stats=stats, labels=lab.array, flush_bytes=256 << 20, tol=1e-12)
        g.close()
        rows.append((1e3 * t_up, 1e3 * float(np.mean(secs)), 1e3 * float(np.median(secs)), 1e3 * float(np.max(secs)), 1e3 * (t_up + float(np.sum(secs)))))
        if ref is None:
            ref = (ll.copy(), labels.copy(), _abi.params_as_flat(params).copy())
        else:
            assert np.array_equal(ll, ref[0]) and np.array_equal(labels, ref[1]) and np.array_equal(_abi.params_as_flat(params), ref[2]), "paths disagree"
    r = np.array(rows[1:])
    print(f"{which} timing={timing}: set-up {np.median(r[:, 0]):.3f} ms, step mean {np.median(r[:, 1]):.4f} median {np.median(r[:, 2]):.4f} max {np.median(r[:, 3]):.3f} ms, "
          f"job {np.median(r[:, 4]):.3f} ms (median of {len(r)} jobs)")
print("both paths: identical log-likelihoods, labels and parameters")
