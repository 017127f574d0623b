"""Full job parity on the metric's workload (cfg2, ~750k windows): 50 EM iterations + final decode on the GPU against the
unmodified reference (oracle/_ref, all host threads).  Prints the label mismatch count and the worst relative
log-likelihood deviation.  Run on the GPU box."""
import os, sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import oracle_lib
from flagger_b200 import _abi, api, synth

for name, factory in (("cfg2", synth.config2), ("cfg3", synth.config3), ("cfg4", synth.config4)):
    wl = factory()
    K = api.best_num_collapsed_comps(int(wl.cov.max()), wl.region_coverages)
    cfg = _abi.make_config(n_regions=wl.n_regions, n_col_comps=K, mean_read_length=wl.avg_alignment_len)
    p0 = api.model_init(cfg, wl.region_coverages, wl.window_len)
    gpu = api.HmmFlaggerGPU(cfg, wl)
    t = time.time(); pg, llg, labg = gpu.run_em(synth.HIFI_ALPHA, p0, 50, tol=1e-12); tg = time.time() - t
    ref = oracle_lib.reference(threads=os.cpu_count())
    t = time.time(); want = ref.run_em(cfg, wl, synth.HIFI_ALPHA, p0, 50, tol=1e-12); tr = time.time() - t
    rel = np.abs(llg - want["logliks"]) / np.abs(want["logliks"])
    prel = np.abs(_abi.params_as_flat(pg) - _abi.params_as_flat(want["params"])) / np.maximum(np.abs(_abi.params_as_flat(want["params"])), 1e-300)
    print(f"{name}: windows {wl.n_windows} chunks {wl.n_chunks} regions {wl.n_regions} K {K} | E-steps {len(llg)} | "
          f"label mismatches {(labg != want['labels']).sum()} / {wl.n_windows} | max rel loglik dev {rel.max():.3e} | "
          f"max rel param dev {prel[_abi.params_as_flat(want['params']) != 0].max():.3e} | gpu {tg:.3f} s, reference ({os.cpu_count()} threads) {tr:.1f} s "
          f"(E-step only {want['estep_seconds']:.1f} s)", flush=True)
    gpu.close()
