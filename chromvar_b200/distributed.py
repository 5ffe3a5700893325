"""Cell sharding across GPUs: one process per GPU (torchrun), torch.distributed for plumbing.

The path shards by cells (columns of the counts matrix): every output column depends only on
its own counts column plus three replicated, cell-independent objects (annotation sets,
background matrix, expectation).  Exactly two cross-shard reductions exist:

  1. per-peak fragment totals (int64[P+1], slot P = grand total) -> all-reduce(sum), feeding
     computeExpectations and the sampler's log10(fragments per peak);
  2. per-annotation variability partials (n_finite, mean, M2, n_nonfinite) -> all-gather and a
     Chan merge (a plain sum of means would be wrong; the merge is order-fixed by rank).

With NCCL both run in place on the engine's device buffers (no host staging); with gloo (CPU
tests) the same functions take host arrays.  The background matrix needs no broadcast: the
sampler is a pure function of (global totals, bias, seed), and init() broadcasts ONE seed from
rank 0 (share_seed) from which every rank derives the same per-call sampler seeds.
"""
import os

import numpy as np


def world():
    """(rank, world_size, local_rank) from the torchrun environment."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def init(backend=None):
    """init_process_group from the torchrun environment (no-op for world_size 1)."""
    import torch
    import torch.distributed as dist
    rank, ws, local = world()
    if ws == 1 or dist.is_initialized():
        return rank, ws, local
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(local)
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    else:
        dist.init_process_group(backend=backend)
    share_seed()
    return rank, ws, local


def share_seed(seed=None):
    """One sampler seed stream for all ranks: rank 0's seed (given, or fresh entropy) is broadcast and
    becomes every rank's api.set_seed().  Each later getBackgroundPeaks() call advances that stream
    identically on every rank, so all cell shards are bias-corrected against the SAME background
    matrix without broadcasting the matrix itself."""
    import torch
    import torch.distributed as dist
    from . import api
    if seed is None:
        seed = int(np.random.SeedSequence().entropy) & (2 ** 62 - 1)
    t = torch.tensor([int(seed)], dtype=torch.int64)
    if dist.is_initialized() and dist.get_world_size() > 1:
        if dist.get_backend() == "nccl":
            t = t.cuda()
        dist.broadcast(t, src=0)
    seed = int(t.item())
    api.set_seed(seed)
    return seed


def shard_columns(colptr, world_size):
    """Contiguous column blocks balanced by nonzeros rather than by cell count.
    Returns world_size + 1 boundaries (boundaries[r] .. boundaries[r+1] is rank r's block)."""
    colptr = np.asarray(colptr, np.int64)
    N = colptr.shape[0] - 1
    nnz = colptr[-1]
    targets = (nnz * np.arange(1, world_size, dtype=np.float64) / world_size)
    cuts = np.searchsorted(colptr, targets, side="left")
    b = np.concatenate([[0], np.clip(cuts, 0, N), [N]]).astype(np.int64)
    return np.maximum.accumulate(b)


def take_columns(C, start, end):
    """Column block [start, end) of a scipy CSC matrix as (colptr, rowidx, x) without copying x."""
    colptr = C.indptr[start:end + 1].astype(np.int64)
    lo, hi = int(colptr[0]), int(colptr[-1])
    return colptr - lo, C.indices[lo:hi], C.data[lo:hi]


def allreduce_peak_totals(engine=None, host_totals=None):
    """Reduction 1.  engine: all-reduce the device buffer in place over NCCL and commit it.
    host_totals (gloo / tests): all-reduce a host int64 array and return it."""
    import torch
    import torch.distributed as dist
    if host_totals is not None:
        t = torch.from_numpy(np.ascontiguousarray(host_totals, np.int64).copy())
        if dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.numpy()
    if dist.is_initialized() and dist.get_world_size() > 1:
        buf = engine.peak_totals_device()
        t = torch.as_tensor(buf, device=torch.device("cuda", engine.device))
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        torch.cuda.synchronize(engine.device)
        engine.peak_totals_commit()
    return None


def gather_variability(engine=None, host_partials=None, na_rm=True):
    """Reduction 2.  Returns the global row SD of z (length K) on every rank."""
    import torch
    import torch.distributed as dist
    from . import _lib
    multi = dist.is_initialized() and dist.get_world_size() > 1
    if host_partials is not None:
        mine = torch.from_numpy(np.ascontiguousarray(host_partials, np.float64).copy())
    else:
        buf = engine.variability_partials_device()
        mine = torch.as_tensor(buf, device=torch.device("cuda", engine.device)).clone()
    if multi:
        parts = [torch.empty_like(mine) for _ in range(dist.get_world_size())]
        dist.all_gather(parts, mine)
    else:
        parts = [mine]
    stacked = torch.stack([p.reshape(-1, 4) for p in parts]).cpu().numpy()
    return _lib.merge_variability(stacked, na_rm=na_rm)


def welford_partials(z):
    """Host restatement of the per-shard partial table (used by the gloo tests)."""
    z = np.asarray(z, np.float64)
    out = np.zeros((z.shape[0], 4))
    for k in range(z.shape[0]):
        f = z[k][np.isfinite(z[k])]
        out[k, 0] = f.size
        out[k, 1] = f.mean() if f.size else 0.0
        out[k, 2] = ((f - out[k, 1]) ** 2).sum() if f.size else 0.0
        out[k, 3] = z.shape[1] - f.size
    return out
