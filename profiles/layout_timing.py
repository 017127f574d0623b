"""Set-up timing (HFG_TIMING=1): where hfg_set_chunks spends its time for the three 3 Gbp workloads."""
import os, sys
os.environ["HFG_TIMING"] = "1"
sys.path.insert(0, ".")
from flagger_b200 import api, synth, _abi
print("host threads", len(os.sched_getaffinity(0)), file=sys.stderr)
for name in ("config2", "config3", "config4"):
    wl = getattr(synth, name)(); cfg = _abi.make_config(n_regions=len(wl.region_coverages), n_col_comps=4)
    print("==", name, file=sys.stderr)
    for i in range(3):
        g = api.HmmFlaggerGPU(cfg, wl); g.close()
