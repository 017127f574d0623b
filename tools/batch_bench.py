"""Batched EM runs against one-after-the-other runs on a GPU box: candidate runs per second for B candidates x `iters`
iterations (+ final inference) over the same windows.   python tools/batch_bench.py [total_bp ...]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from flagger_b200 import _abi, api, synth  # noqa: E402

B, ITERS = 8, 50
for bp in [float(a) for a in sys.argv[1:]] or [3e8, 3e9]:
    wl = synth.config2(total_bp=int(bp), seed=22)
    K = api.best_num_collapsed_comps(int(wl.cov.max()), wl.region_coverages)
    cfg = _abi.make_config(n_regions=1, n_col_comps=K)
    p0 = api.model_init(cfg, wl.region_coverages, wl.window_len)
    rng = np.random.default_rng(1)
    alphas = np.array([np.clip(synth.HIFI_ALPHA * rng.uniform(0.5, 1.4, (4, 4)), 0, 0.95) for _ in range(B)])
    one = api.HmmFlaggerGPU(cfg, wl)
    one.run_em(alphas[0], p0, ITERS, tol=1e-12)
    t0 = time.perf_counter()
    seq = [one.run_em(a, p0, ITERS, tol=1e-12) for a in alphas]
    t_seq = time.perf_counter() - t0
    one.close()
    line = f"{wl.n_windows} windows, {B} candidates x {ITERS} iterations: one after the other {1e3 * t_seq:.2f} ms ({B / t_seq:.1f} runs/s)"
    for lanes in (2, 4, 8):
        batch = api.HmmFlaggerBatch(cfg, wl, n_lanes=lanes)
        batch.run_em(alphas, p0, ITERS, tol=1e-12)
        t0 = time.perf_counter()
        _, ll, lab = batch.run_em(alphas, p0, ITERS, tol=1e-12)
        t_b = time.perf_counter() - t0
        batch.close()
        same = all(np.array_equal(lab[r], seq[r][2]) for r in range(B))
        line += f"; {lanes} lanes {1e3 * t_b:.2f} ms ({B / t_b:.1f} runs/s, x{t_seq / t_b:.2f}, labels identical: {same})"
    print(line, flush=True)
