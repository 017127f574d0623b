"""GPU parity at BASELINE.json's full sizes (configs[1..3]): ~750k windows each.  The checker is the unmodified
reference (oracle/_ref, multi-threaded) when it was built, else the single-threaded restatement with fewer iterations.
Size-independent properties are checked too: posteriors sum to 1, transition counts sum to the number of counted pairs,
the log-likelihood of the sharded run equals the sum over shards."""
import os

import numpy as np
import pytest

import oracle_lib
from flagger_b200 import _abi, api, synth
from flagger_b200 import dist as hdist

pytestmark = pytest.mark.gpu


def _checker():
    ref = oracle_lib.reference(threads=os.cpu_count() or 4)
    return (ref, 10) if ref is not None else (oracle_lib.oracle(), 2)


def _setup(wl):
    K = api.best_num_collapsed_comps(int(wl.cov.max()), wl.region_coverages)
    cfg = _abi.make_config(n_regions=wl.n_regions, n_col_comps=K, mean_read_length=wl.avg_alignment_len)
    return cfg, api.model_init(cfg, wl.region_coverages, wl.window_len)


@pytest.mark.parametrize("factory", [synth.config2, synth.config3, synth.config4], ids=["cfg2", "cfg3", "cfg4"])
def test_full_size_em_matches_reference(factory):
    wl = factory()
    cfg, params = _setup(wl)
    chk, iters = _checker()
    gpu = api.HmmFlaggerGPU(cfg, wl)
    pg, llg, labg = gpu.run_em(synth.HIFI_ALPHA, params, iters, tol=1e-12)
    want = chk.run_em(cfg, wl, synth.HIFI_ALPHA, params, iters, tol=1e-12)
    assert len(llg) == len(want["logliks"]) == iters + 1
    assert np.all(np.abs(llg - want["logliks"]) <= 1e-9 * np.abs(want["logliks"]))
    mism = int((labg != want["labels"]).sum())
    assert mism == 0, f"{mism} of {wl.n_windows} final labels differ"
    assert np.allclose(_abi.params_as_flat(pg), _abi.params_as_flat(want["params"]), rtol=1e-7, atol=0)
    gpu.close()


def test_full_size_properties():
    wl = synth.config2()
    cfg, params = _setup(wl)
    gpu = api.HmmFlaggerGPU(cfg, wl)
    stats, ll, labels = gpu.em_iteration(synth.HIFI_ALPHA, params)
    post = gpu.posteriors()
    assert np.all(np.abs(post.sum(axis=1) - 1.0) < 1e-12) and post.min() >= 0.0
    assert np.array_equal(labels, post.argmax(axis=1).astype(np.int8))  # first maximum == numpy's argmax
    # pairs (i -> i+1), i = 1 .. L-2, each contribute total weight 1 to the 4x4 counts of their region
    L = wl.chunks["n_windows"].astype(np.int64)
    n_pairs = int(np.maximum(L - 2, 0).sum())
    assert abs(stats["trans_count"].sum() - n_pairs) <= 1e-6 * n_pairs
    # the per-chunk log-likelihoods add up, and sharding the chunks over 4 "ranks" changes nothing but the association
    assert abs(gpu.chunk_logliks().sum() - ll) <= 1e-10 * abs(ll)
    flat = np.zeros_like(_abi.stats_as_flat(stats))
    ll_sum = 0.0
    for r in range(4):
        g = api.HmmFlaggerGPU(cfg, hdist.shard_chunks(wl, r, 4))
        s, l, _ = g.em_iteration(synth.HIFI_ALPHA, params)
        flat += _abi.stats_as_flat(s)
        ll_sum += l
        g.close()
    assert abs(ll_sum - ll) <= 1e-11 * abs(ll)
    assert np.allclose(flat, _abi.stats_as_flat(stats), rtol=1e-10, atol=1e-9)
    gpu.close()
