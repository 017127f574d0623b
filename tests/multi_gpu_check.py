"""Multi-GPU parity check, one process per GPU (launched by tests/test_gpu_multi.py or by hand):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
        tests/multi_gpu_check.py [--workload small|cfg2|cfg3|cfg4] [--iters K]

Every rank owns a contiguous range of chunks (flagger_b200.dist.shard_chunks).  Checked on every rank, K EM iterations:
  * fused path (the E-step kernel sums statistics + log-likelihood over the ranks through NVLink peer memory): the result
    is IDENTICAL on all ranks (bitwise), equals the sum of the per-rank single-GPU results to 1e-12 relative, and equals a
    single-GPU run over all chunks to 1e-10 relative;
  * labels of the shard == labels of the same windows in the single-GPU run over all chunks (bit-exact);
  * the NCCL variant (hfg_em_iteration_device + dist.all_reduce) gives the same sums to 1e-12 relative;
  * forward-only passes (SQUAREM) sum the log-likelihood the same way;
  * the device-resident EM loop (hfg_run_em: M-step in the kernel tail on the all-reduced statistics) ends with the same
    parameters on every rank (bitwise), and with the log-likelihoods / labels of the single-GPU loop.
Prints "MULTI_GPU_CHECK OK ranks=N" on rank 0."""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flagger_b200 import _abi, api, synth  # noqa: E402
from flagger_b200 import dist as hdist  # noqa: E402


def flat(stats):
    return _abi.stats_as_flat(stats)


def close(a, b, rtol):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    scale = np.maximum(np.abs(a), np.abs(b))
    return bool(np.all(np.abs(a - b) <= rtol * np.maximum(scale, 1e-300) + 1e-290))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="small")
    ap.add_argument("--iters", type=int, default=4)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    if args.workload == "small":
        wl_full = synth.small_mixed(n_regions=3, seed=91)
    else:
        wl_full = getattr(synth, {"cfg2": "config2", "cfg3": "config3", "cfg4": "config4"}[args.workload])()
    R = wl_full.n_regions
    K = api.best_num_collapsed_comps(int(wl_full.cov.max()), wl_full.region_coverages)
    cfg = _abi.make_config(n_regions=R, n_col_comps=K, mean_read_length=wl_full.avg_alignment_len)
    cfg["device"] = local
    alpha = synth.HIFI_ALPHA
    params = api.model_init(cfg, wl_full.region_coverages, wl_full.window_len)
    params0 = params.copy()
    b = hdist.shard_bounds(wl_full.chunks["n_windows"], world)
    wl = hdist.shard_chunks(wl_full, rank, world)
    lo = int(wl_full.chunks["offset"][b[rank]]) if wl.n_chunks else 0

    solo_all = api.HmmFlaggerGPU(cfg, wl_full)    # single-GPU run over all chunks (the truth for this check)
    solo_own = api.HmmFlaggerGPU(cfg, wl)         # this rank's shard, no exchange
    fused = api.HmmFlaggerGPU(cfg, wl)
    fused.peer_connect(dist)
    n_d = fused.stats_device_bytes() // 8
    SD = _abi.region_stats_dtype.itemsize // 8
    ok = True

    def check(cond, what):
        nonlocal ok
        if not cond:
            ok = False
            print(f"[rank {rank}] FAILED: {what}", flush=True)

    for it in range(args.iters):
        s_all, ll_all, lab_all = solo_all.em_iteration(alpha, params)
        s_all = s_all.copy()
        s_own, ll_own, _ = solo_own.em_iteration(alpha, params, want_labels=False)
        own = torch.from_numpy(np.concatenate([flat(s_own), [ll_own]])).to(dev)
        dist.all_reduce(own, op=dist.ReduceOp.SUM)
        own = own.cpu().numpy()
        s_f, ll_f, lab_f = fused.em_iteration(alpha, params)
        mine = np.concatenate([flat(s_f), [ll_f]])
        # identical on every rank
        g = [torch.zeros(len(mine), dtype=torch.float64, device=dev) for _ in range(world)]
        dist.all_gather(g, torch.from_numpy(mine).to(dev))
        check(all(torch.equal(g[0], x) for x in g), f"iter {it}: fused result differs between ranks")
        check(close(mine, own, 1e-12), f"iter {it}: fused sum != sum of per-rank results")
        check(close(mine, np.concatenate([flat(s_all), [ll_all]]), 1e-10), f"iter {it}: fused sum != single-GPU run")
        check(np.array_equal(lab_f, lab_all[lo:lo + wl.n_windows]), f"iter {it}: shard labels differ from the single-GPU run")
        # NCCL variant
        stats_dev = torch.zeros(n_d, dtype=torch.float64, device=dev)
        stream = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(stream):
            solo_own.em_iteration_device(alpha, params, stats_dev.data_ptr(), stream.cuda_stream)
            dist.all_reduce(stats_dev, op=dist.ReduceOp.SUM)
        stream.synchronize()
        nc = stats_dev.cpu().numpy()
        check(nc[n_d - 1] == 0, f"iter {it}: error flags {nc[n_d - 1]}")
        check(close(np.concatenate([nc[: R * SD], [nc[n_d - 2]]]), mine, 1e-12), f"iter {it}: NCCL sums != fused sums")
        # forward-only (SQUAREM step trials)
        llf = fused.forward_only(alpha, params)
        check(close([llf], [ll_f], 1e-13), f"iter {it}: fused forward-only log-likelihood {llf} != {ll_f}")
        params, _ = api.mstep(cfg, params, s_f, tol=1e-12)
    # device-resident EM loop over the ranks
    p_all, ll_loop_all, lab_loop_all = solo_all.run_em(alpha, params0, args.iters, tol=1e-12)
    p_f, ll_loop_f, lab_loop_f = fused.run_em(alpha, params0, args.iters, tol=1e-12)
    pf = torch.from_numpy(_abi.params_as_flat(p_f).copy()).to(dev)
    g = [torch.zeros_like(pf) for _ in range(world)]
    dist.all_gather(g, pf)
    check(all(torch.equal(g[0], x) for x in g), "device EM loop: parameters differ between ranks")
    check(len(ll_loop_f) == args.iters + 1 and close(ll_loop_f, ll_loop_all, 1e-10), "device EM loop: log-likelihoods differ")
    check(close(_abi.params_as_flat(p_f), _abi.params_as_flat(p_all), 1e-8), "device EM loop: parameters != single-GPU loop")
    check(np.array_equal(lab_loop_f, lab_loop_all[lo:lo + wl.n_windows]), "device EM loop: shard labels differ")
    flag = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(flag)
    for g_ in (fused, solo_own, solo_all):
        g_.close()
    dist.barrier()
    if rank == 0:
        print(("MULTI_GPU_CHECK OK" if flag.item() == 0 else "MULTI_GPU_CHECK FAILED") + f" ranks={world} "
              f"workload={wl_full.name} windows={wl_full.n_windows} iters={args.iters}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 0 else 1)


if __name__ == "__main__":
    main()
