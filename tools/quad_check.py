"""A/B of the two E-step kernels on a GPU box: the quad kernel (default) against the first-generation one (HFG_KERNEL=v1)
and against the CPU oracle, with the differences spelled out (which statistic, which window).
    python tools/quad_check.py [small|medium|cfg2|cfg3|cfg4 ...]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from flagger_b200 import _abi, api, synth  # noqa: E402


CANDIDATE = os.environ.get("HFG_CANDIDATE", "")  # "" = the default kernel, "quad" = the four-lanes-per-segment kernel


def make(kind):
    if kind == "small":
        return synth.small_mixed(n_regions=3, seed=14)
    if kind == "small1":
        return synth.small_mixed(n_regions=1, seed=12)
    if kind == "medium":
        return synth.config2(total_bp=300_000_000, seed=22)
    return {"cfg2": synth.config2, "cfg3": synth.config3, "cfg4": synth.config4}[kind]()


def ctx(kernel, cfg, wl):
    if kernel == "v1":
        os.environ["HFG_KERNEL"] = "v1"
    elif CANDIDATE:
        os.environ["HFG_KERNEL"] = CANDIDATE
    else:
        os.environ.pop("HFG_KERNEL", None)
    return api.HmmFlaggerGPU(cfg, wl, timing=True)


def diff_stats(a, b, tag):
    fa, fb = _abi.stats_as_flat(a), _abi.stats_as_flat(b)
    scale = np.abs(fb).max()
    d = np.abs(fa - fb) / scale
    i = int(d.argmax())
    print(f"  {tag}: max |d stats| / scale = {d.max():.3e} at flat index {i} ({fa.ravel()[i]!r} vs {fb.ravel()[i]!r})")
    return d.max()


def main():
    kinds = sys.argv[1:] or ["small", "medium"]
    ok = True
    for kind in kinds:
        wl = make(kind)
        R = len(wl.region_coverages)
        K = api.best_num_collapsed_comps(int(wl.cov.max()), wl.region_coverages)
        cfg = _abi.make_config(n_regions=R, n_col_comps=K)
        params = api.model_init(cfg, wl.region_coverages, wl.window_len)
        alpha = synth.HIFI_ALPHA
        print(f"== {kind}: {wl.n_windows} windows, {wl.n_chunks} chunks, R={R}, K={K}", flush=True)
        res = {}
        for kernel in ("v1", "quad"):
            g = ctx(kernel, cfg, wl)
            st, ll, lab = g.em_iteration(alpha, params)
            ms = []
            for _ in range(5):
                g.em_iteration(alpha, params, want_labels=False)
                ms.append(g.last_estep_kernel_ms())
            p2, _ = api.mstep(cfg, params, st)
            st2, ll2, lab2 = g.em_iteration(alpha, p2)
            cl = g.chunk_logliks()
            fo = g.forward_only(alpha, p2)
            # device-resident loop
            pe, lle, labe = g.run_em(alpha, params, 4, tol=1e-12)
            g.em_begin(alpha, params, tol=1e-12, max_esteps=6)
            for _ in range(6):
                g.em_enqueue()
            g.em_finish(want_labels=False)
            em_ms = [g.em_enqueued_ms(i) for i in range(6)]
            res[kernel] = dict(st=st.copy(), ll=ll, lab=lab.copy(), st2=st2.copy(), ll2=ll2, lab2=lab2.copy(), cl=cl, fo=fo,
                               pe=pe.copy(), lle=lle.copy(), labe=labe.copy())
            print(f"  {kernel}: kernel ms {[round(m, 4) for m in ms]}  device EM iteration ms {[round(m, 4) for m in em_ms]}",
                  flush=True)
            if kind in ("small", "small1", "medium"):
                g.em_iteration(alpha, params)
                res["post_" + kernel] = g.posteriors()
            g.close()
        a, b = res["quad"], res["v1"]
        print(f"  quad vs v1: loglik rel {abs(a['ll'] - b['ll']) / abs(b['ll']):.3e}, iteration 2 {abs(a['ll2'] - b['ll2']) / abs(b['ll2']):.3e},"
              f" forward-only {abs(a['fo'] - b['fo']) / abs(b['fo']):.3e}; labels differ {int((a['lab'] != b['lab']).sum())} /"
              f" {int((a['lab2'] != b['lab2']).sum())}; chunk logliks max rel {np.max(np.abs(a['cl'] - b['cl']) / np.abs(b['cl'])):.3e}")
        d1 = diff_stats(a["st"], b["st"], "quad vs v1, iteration 1")
        d2 = diff_stats(a["st2"], b["st2"], "quad vs v1, iteration 2")
        pa, pb = _abi.params_as_flat(a["pe"]), _abi.params_as_flat(b["pe"])
        nz = np.abs(pb) > 0
        dp = np.max(np.abs(pa[nz] - pb[nz]) / np.abs(pb[nz]))
        print(f"  device loop (4 iterations + final): logliks rel {np.max(np.abs(a['lle'] - b['lle']) / np.abs(b['lle'])):.3e},"
              f" params rel {dp:.3e}, labels differ {int((a['labe'] != b['labe']).sum())}")
        bad = (a["lab"] != b["lab"]).sum() + (a["lab2"] != b["lab2"]).sum() + (a["labe"] != b["labe"]).sum()
        if bad or max(d1, d2) > 1e-9 or dp > 1e-8 or abs(a["ll"] - b["ll"]) > 1e-9 * abs(b["ll"]):
            ok = False
            w = np.nonzero(a["lab"] != b["lab"])[0]
            print("  MISMATCH; first differing windows:", w[:20])
        if kind in ("small", "small1", "medium"):
            import oracle_lib
            orc = oracle_lib.oracle()
            out = orc.estep(cfg, wl, alpha, params)
            print(f"  quad vs oracle: loglik rel {abs(a['ll'] - out['loglik']) / abs(out['loglik']):.3e}, labels differ"
                  f" {int((a['lab'] != out['labels']).sum())}")
            d = diff_stats(a["st"], out["stats"], "quad vs oracle")
            big = out["posteriors"] > 1e-200
            for kernel in ("v1", "quad"):
                rel = np.where(big, np.abs(res["post_" + kernel] - out["posteriors"]) / np.maximum(out["posteriors"], 1e-300), 0.0)
                pr = rel.max()
                wi = np.unravel_index(int(rel.argmax()), rel.shape)
                print(f"  posteriors {kernel}: max rel {pr:.3e} at window {wi[0]} state {wi[1]}: {res['post_' + kernel][wi]!r} vs"
                      f" {out['posteriors'][wi]!r}; row {res['post_' + kernel][wi[0]]} vs {out['posteriors'][wi[0]]}; windows with rel > 1e-5:"
                      f" {int((rel.max(axis=1) > 1e-5).sum())}, first {np.nonzero(rel.max(axis=1) > 1e-5)[0][:12]}")
            if d > 1e-9 or pr > 1e-5 or (a["lab"] != out["labels"]).any():
                ok = False
    print("QUAD_CHECK", "OK" if ok else "FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    t0 = time.time()
    rc = main()
    print(f"({time.time() - t0:.1f} s)")
    sys.exit(rc)
