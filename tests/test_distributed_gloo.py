"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: nnz-balanced column sharding and
the two cross-shard reductions (per-peak totals all-reduce, variability partial merge)."""
import os
import socket
import sys

import numpy as np
import scipy.sparse as sp
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make():
    rng = np.random.default_rng(12)
    C = sp.random(300, 41, density=0.15, random_state=np.random.RandomState(3), format="csc")
    C.data = np.ceil(C.data * 4)
    z = rng.normal(size=(7, 41))
    z[2, 5] = np.nan
    z[4, 30:] = np.nan
    return C, z


def _worker(rank, ws, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(ws), LOCAL_RANK=str(rank))
    import torch.distributed as dist
    from chromvar_b200 import distributed as cvd
    cvd.init(backend="gloo")
    C, z = _make()
    b = cvd.shard_columns(C.indptr, ws)
    lo, hi = int(b[rank]), int(b[rank + 1])
    colptr, rowidx, x = cvd.take_columns(C, lo, hi)
    local = sp.csc_matrix((x, rowidx, colptr), shape=(C.shape[0], hi - lo))
    tot = np.zeros(C.shape[0] + 1, np.int64)
    tot[:-1] = np.asarray(local.sum(axis=1)).ravel().astype(np.int64)
    tot[-1] = int(local.sum())
    glob = cvd.allreduce_peak_totals(host_totals=tot)
    sd = cvd.gather_variability(host_partials=cvd.welford_partials(z[:, lo:hi]), na_rm=True)
    sd_all = cvd.gather_variability(host_partials=cvd.welford_partials(z[:, lo:hi]), na_rm=False)
    # init() broadcast one seed: every rank now draws the same per-call sampler seeds
    from chromvar_b200 import api
    seeds = [api._next_sampler_seed() for _ in range(3)]
    q.put((rank, lo, hi, glob, sd, sd_all, seeds))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_reductions():
    ws, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, ws, port, q)) for r in range(ws)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(ws))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    C, z = _make()
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == C.shape[1]
    ref_tot = np.append(np.asarray(C.sum(axis=1)).ravel(), C.sum()).astype(np.int64)
    ref_sd = np.array([np.std(r[np.isfinite(r)], ddof=1) for r in z])
    assert res[0][6] == res[1][6] and len(set(res[0][6])) == 3      # same stream on both ranks, a fresh seed per call
    for _, _, _, glob, sd, sd_all, _ in res:
        assert np.array_equal(glob, ref_tot)
        assert np.allclose(sd, ref_sd, rtol=1e-12)
        assert np.isnan(sd_all[[2, 4]]).all() and np.allclose(np.delete(sd_all, [2, 4]), np.delete(ref_sd, [2, 4]), rtol=1e-12)


def test_shard_columns_balances_nonzeros():
    from chromvar_b200 import distributed as cvd
    rng = np.random.default_rng(0)
    lens = np.concatenate([rng.integers(0, 5, 500), rng.integers(50, 200, 100), np.zeros(30, int)])
    colptr = np.concatenate([[0], np.cumsum(lens)])
    for ws in (1, 2, 4, 8):
        b = cvd.shard_columns(colptr, ws)
        assert b[0] == 0 and b[-1] == len(lens) and np.all(np.diff(b) >= 0)
        per = np.diff(colptr[b])
        assert per.sum() == colptr[-1]
        assert per.max() <= colptr[-1] / ws + lens.max()
