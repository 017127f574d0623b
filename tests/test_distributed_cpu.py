"""CPU, world_size 2 over gloo: the host logic of the multi-GPU path.  Chunks are sharded across ranks, each rank runs
the E-step on its shard (here the CPU oracle stands in for the kernel: this test is about the sharding and the single
all-reduce, not about the arithmetic), the statistics are summed with ONE all-reduce, and every rank runs the same host
M-step.  The result must equal the single-process run."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import oracle_lib
    from flagger_b200 import _abi, api, synth
    from flagger_b200 import dist as hdist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc = oracle_lib.oracle()
    wl = synth.small_mixed(n_regions=3, seed=77)
    cfg = _abi.make_config(n_regions=3, n_col_comps=4)
    params = api.model_init(cfg, wl.region_coverages, wl.window_len)
    shard = hdist.shard_chunks(wl, rank, world)
    logliks = []
    for _ in range(3):
        e = orc.estep(cfg, shard, synth.HIFI_ALPHA, params)
        stats, ll = hdist.allreduce_stats_host(e["stats"], e["loglik"], dist)
        params, _ = api.mstep(cfg, params, stats, tol=1e-12)
        logliks.append(ll)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), params=params, logliks=np.array(logliks),
             n_windows=shard.n_windows, n_chunks=shard.n_chunks)
    dist.destroy_process_group()


def test_two_rank_em_equals_single_process(tmp_path, orc):
    from flagger_b200 import _abi, api, synth
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    wl = synth.small_mixed(n_regions=3, seed=77)
    assert int(r0["n_windows"]) + int(r1["n_windows"]) == wl.n_windows
    assert int(r0["n_chunks"]) + int(r1["n_chunks"]) == wl.n_chunks and int(r0["n_chunks"]) > 0 < int(r1["n_chunks"])
    # every rank ends with the same parameters (identical all-reduced statistics -> identical host M-step)
    assert np.array_equal(r0["params"].view(np.float64), r1["params"].view(np.float64))
    assert np.array_equal(r0["logliks"], r1["logliks"])
    # ... and they match the single-process run up to the association of the sums
    cfg = _abi.make_config(n_regions=3, n_col_comps=4)
    params = api.model_init(cfg, wl.region_coverages, wl.window_len)
    ll = []
    for _ in range(3):
        e = orc.estep(cfg, wl, synth.HIFI_ALPHA, params)
        params, _ = api.mstep(cfg, params, e["stats"], tol=1e-12)
        ll.append(e["loglik"])
    assert np.allclose(r0["logliks"], ll, rtol=1e-12, atol=0)
    assert np.allclose(r0["params"].view(np.float64), params.view(np.float64), rtol=1e-9, atol=0)


def test_shard_bounds_balanced():
    from flagger_b200 import dist as hdist, synth
    wl = synth.config2(total_bp=3_000_000_000)
    for world in (1, 2, 4, 8):
        b = hdist.shard_bounds(wl.chunks["n_windows"], world)
        assert b[0] == 0 and b[-1] == wl.n_chunks and all(x < y for x, y in zip(b, b[1:]))
        sizes = [int(wl.chunks["n_windows"][b[r]:b[r + 1]].sum()) for r in range(world)]
        assert sum(sizes) == wl.n_windows and max(sizes) <= 1.15 * wl.n_windows / world


def test_shard_bounds_never_empty():
    """Ragged chunk sizes (one chunk holding most windows): every rank still owns >= 1 chunk; fewer chunks than ranks is
    refused with a message instead of producing an empty shard."""
    import pytest
    from flagger_b200 import dist as hdist, synth
    wl = synth.small_mixed(n_regions=3, seed=91)
    for world in range(1, wl.n_chunks + 1):
        b = hdist.shard_bounds(wl.chunks["n_windows"], world)
        assert b[0] == 0 and b[-1] == wl.n_chunks and all(x < y for x, y in zip(b, b[1:])), (world, b)
        assert sum(hdist.shard_chunks(wl, r, world).n_windows for r in range(world)) == wl.n_windows
    with pytest.raises(ValueError):
        hdist.shard_bounds(wl.chunks["n_windows"], wl.n_chunks + 1)
    assert hdist.shard_bounds([5, 1, 1, 1], 4) == [0, 1, 2, 3, 4]
    b = hdist.shard_bounds([1, 1, 1, 50], 3)
    assert b[0] == 0 and b[-1] == 4 and all(x < y for x, y in zip(b, b[1:]))
