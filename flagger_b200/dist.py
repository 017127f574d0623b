"""Multi-GPU host logic: chunks are independent chains inside an E-step, so they shard across ranks with no data-path
collective; the only exchange is ONE sum all-reduce of the EM statistics (+ log-likelihood) per iteration, after which
every rank runs the identical host M-step (the reference merges per-chunk estimators serially,
hmm.c:759-763).  torch.distributed is plumbing here: NCCL over NVLink on GPUs, gloo in the CPU tests."""
import numpy as np

from . import _abi


def shard_bounds(n_windows_per_chunk, world):
    """Contiguous chunk ranges in list order, balanced by window count: rank r owns chunks [b[r], b[r+1]).  Every rank
    gets at least one chunk (a chunk is the unit: it cannot be split); fewer chunks than ranks is an error."""
    n = np.asarray(n_windows_per_chunk, dtype=np.int64)
    if len(n) < world:
        raise ValueError(f"{len(n)} chunks cannot be sharded over {world} ranks: use at most {len(n)} GPUs "
                         "(or a smaller --chunkLen)")
    cum = np.cumsum(n)
    total = int(cum[-1])
    bounds = [0]
    for r in range(1, world):
        k = int(np.searchsorted(cum, total * r / world, side="left")) + 1
        bounds.append(min(max(k, bounds[-1] + 1), len(n)))
    bounds.append(len(n))
    for r in range(world - 1, 0, -1):  # leave one chunk for each of the ranks behind
        bounds[r] = min(bounds[r], bounds[r + 1] - 1)
    return bounds


def shard_chunks(wl, rank, world):
    if world == 1:
        return wl
    b = shard_bounds(wl.chunks["n_windows"], world)
    return wl.subset(range(b[rank], b[rank + 1]), name=f"{wl.name}[rank{rank}/{world}]")


def allreduce_stats_host(stats, loglik, dist, device=None):
    """Sum the flat statistics vector and the log-likelihood over ranks (host numpy in/out)."""
    import torch
    flat = np.concatenate([_abi.stats_as_flat(stats), [loglik]])
    t = torch.from_numpy(flat)
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    out = t.cpu().numpy()
    return out[:-1].copy().view(_abi.region_stats_dtype), float(out[-1])
