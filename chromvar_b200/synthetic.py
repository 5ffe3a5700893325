"""Synthetic inputs of the shapes BASELINE.json names.

There is no network for real scATAC data, so bench.py and the large-size tests use these
generators.  They produce, deterministically from a seed:
  * a sparse peaks x cells fragment-count matrix (CSC) with log-normal peak accessibility,
    log-normal cell depth and a few planted cell types that up-weight the peaks of some motifs,
  * a GC-bias vector,
  * K peak-set annotations (CSR of 1-based peak indices) with log-uniform match fractions and
    GC dependence.
"""
import numpy as np
import scipy.sparse as sp

CONFIGS = {
    # name: (P, N, K, density, seed)
    "cfg2_scatac_motifs": dict(P=100_000, N=10_000, K=870, density=0.03, seed=20171, kind="motif"),
    "cfg3_kmers": dict(P=200_000, N=50_000, K=2080, density=0.03, seed=20174, kind="kmer"),
    "cfg4_bulk": dict(P=500_000, N=200, K=870, density=1.0, seed=20175, kind="motif"),
    "cfg5_atlas_shard": dict(P=500_000, N=125_000, K=870, density=0.01, seed=20176, kind="motif"),
    "small": dict(P=20_000, N=2_000, K=64, density=0.03, seed=7, kind="motif"),
    "tiny": dict(P=2_000, N=300, K=12, density=0.05, seed=3, kind="motif"),
}


def make_bias(P, rng):
    return np.round(np.clip(rng.normal(0.52, 0.09, P), 0.25, 0.85), 3)


def make_annotations(P, K, bias, rng, kind="motif"):
    """(annoptr int64[K+1], annoidx int32 1-based, sorted within a set)."""
    if kind == "kmer":
        f = np.where(np.arange(K) < K - 64, 0.215, 0.114) * rng.uniform(0.8, 1.2, K)
    else:
        f = np.exp(rng.uniform(np.log(0.005), np.log(0.30), K))
    s = rng.uniform(-1.0, 1.0, K)
    ptr = [0]
    idx = []
    for k in range(K):
        prob = np.clip(f[k] * (1.0 + 2.0 * (bias - 0.52) * s[k]), 0.0, 1.0)
        members = np.nonzero(rng.random(P) < prob)[0]
        if members.size == 0:
            members = np.array([rng.integers(P)])
        idx.append(members.astype(np.int32) + 1)
        ptr.append(ptr[-1] + members.size)
    return np.asarray(ptr, np.int64), np.concatenate(idx)


def make_counts(P, N, density, rng, annoptr=None, annoidx=None, n_types=5, planted=20,
                dtype=np.float64):
    """CSC fragment counts: column c ~ Multinomial(D_c, a * m_type(c)); every peak gets >= 1."""
    a = rng.lognormal(0.0, 1.0, P)
    a /= a.sum()
    # duplicates (two fragments in one peak of one cell) cost a few percent of the nonzeros
    mean_depth = density * P * 1.08
    depth = np.maximum(1, rng.lognormal(np.log(mean_depth), 0.5, N).astype(np.int64))
    depth = np.minimum(depth, 20 * int(mean_depth) + 10)
    types = rng.integers(0, n_types, N)
    K = 0 if annoptr is None else len(annoptr) - 1
    cols_i, cols_x, indptr = [None] * N, [None] * N, np.zeros(N + 1, np.int64)
    for t in range(n_types):
        w = a.copy()
        if K:
            for k in rng.choice(K, size=min(planted, K), replace=False):
                members = annoidx[annoptr[k]:annoptr[k + 1]] - 1
                w[members] *= rng.uniform(1.5, 3.0)
        cdf = np.cumsum(w)
        cdf /= cdf[-1]
        cells = np.nonzero(types == t)[0]
        if cells.size == 0:
            continue
        tot = int(depth[cells].sum())
        draws = np.searchsorted(cdf, rng.random(tot), side="right").astype(np.int64)
        np.minimum(draws, P - 1, out=draws)
        owner = np.repeat(np.arange(cells.size, dtype=np.int64), depth[cells])
        key = owner * P + draws
        uniq, cnt = np.unique(key, return_counts=True)
        # single-cell counts stay <= 127 (bundled data: max 8); the log-normal tails of
        # depth x accessibility would otherwise put a handful of entries in the hundreds
        np.minimum(cnt, 127, out=cnt)
        o = uniq // P
        r = (uniq - o * P).astype(np.int32)
        bounds = np.searchsorted(o, np.arange(cells.size + 1))
        for j, c in enumerate(cells):
            cols_i[c] = r[bounds[j]:bounds[j + 1]]
            cols_x[c] = cnt[bounds[j]:bounds[j + 1]]
    for c in range(N):
        if cols_i[c] is None:
            cols_i[c] = np.zeros(0, np.int32)
            cols_x[c] = np.zeros(0, np.int64)
        indptr[c + 1] = indptr[c] + cols_i[c].size
    indices = np.concatenate(cols_i).astype(np.int32)
    data = np.concatenate(cols_x).astype(dtype)
    C = sp.csc_matrix((data, indices, indptr), shape=(P, N))
    # the reference stop()s on an all-zero peak (R/compute_deviations.R:362): give each one a fragment
    empty = np.nonzero(np.bincount(indices, minlength=P) == 0)[0]
    if empty.size:
        extra = sp.csc_matrix((np.ones(empty.size, dtype), (empty, rng.integers(0, N, empty.size))), shape=(P, N))
        C = (C + extra).tocsc()
    C.sort_indices()
    return C


def make_counts_device(P, N, density, rng, annoptr=None, annoidx=None, device="cuda", n_types=5,
                       planted=20, block_cells=16384, pin=True):
    """The generator of make_counts with the bulk of the work (categorical draws by inverse CDF,
    duplicate merging by sort) on a torch device: the atlas shard (config 5: 7e8 nonzeros) takes
    seconds instead of minutes.  Same model, different random stream (torch's generator seeded from
    `rng`), so the matrix is NOT the one make_counts gives for the same seed.  Cells are generated
    in contiguous blocks; with the global cell index as the high part of the sort key the sorted
    unique keys of a block ARE its CSC columns.  Host arrays are pinned when the device is CUDA."""
    import torch
    dev = torch.device(device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(int(rng.integers(1 << 62)))
    a = rng.lognormal(0.0, 1.0, P)
    a /= a.sum()
    mean_depth = density * P * 1.08
    depth = np.maximum(1, rng.lognormal(np.log(mean_depth), 0.5, N).astype(np.int64))
    depth = np.minimum(depth, 20 * int(mean_depth) + 10)
    types = rng.integers(0, n_types, N)
    K = 0 if annoptr is None else len(annoptr) - 1
    cdfs = []
    for t in range(n_types):
        w = a.copy()
        if K:
            for k in rng.choice(K, size=min(planted, K), replace=False):
                members = annoidx[annoptr[k]:annoptr[k + 1]] - 1
                w[members] *= rng.uniform(1.5, 3.0)
        cdf = np.cumsum(w)
        cdfs.append(torch.from_numpy(cdf / cdf[-1]).to(dev))
    depth_d, types_d = torch.from_numpy(depth).to(dev), torch.from_numpy(types).to(dev)
    keys, cnts = [], []
    peak_seen = torch.zeros(P, dtype=torch.bool, device=dev)
    for c0 in range(0, N, block_cells):
        c1 = min(N, c0 + block_cells)
        part = []
        for t in range(n_types):
            cells = torch.nonzero(types_d[c0:c1] == t).ravel() + c0
            if cells.numel() == 0:
                continue
            dep = depth_d[cells]
            tot = int(dep.sum().item())
            u = torch.rand(tot, generator=gen, device=dev, dtype=torch.float64)
            draws = torch.searchsorted(cdfs[t], u, right=True).clamp_(max=P - 1)
            part.append(torch.repeat_interleave(cells, dep) * P + draws)
        key = torch.cat(part) if part else torch.zeros(0, dtype=torch.int64, device=dev)
        uniq, cnt = torch.unique(key, return_counts=True)          # sorted: cell-major, peaks ascending
        peak_seen[uniq % P] = True
        keys.append(uniq)
        cnts.append(cnt.clamp_(max=127).to(torch.int32))
    key = torch.cat(keys)
    cnt = torch.cat(cnts)
    del keys, cnts
    # the reference stop()s on an all-zero peak (R/compute_deviations.R:362): give each one a fragment
    empty = torch.nonzero(~peak_seen).ravel()
    if empty.numel():
        cells = torch.from_numpy(rng.integers(0, N, int(empty.numel()))).to(dev)
        key = torch.cat([key, cells * P + empty])
        cnt = torch.cat([cnt, torch.ones(empty.numel(), dtype=torch.int32, device=dev)])
        key, order = torch.sort(key)
        cnt = cnt[order]
        del order
    nnz = int(key.numel())
    owner = key // P
    per_cell = torch.bincount(owner, minlength=N)
    rows = (key - owner * P).to(torch.int32)
    del key, owner
    it = np.int32 if nnz < 2 ** 31 - 1 else np.int64

    def host(t, dtype):
        t = t.to(dtype)
        if dev.type == "cuda" and pin:
            out = torch.empty(t.shape, dtype=dtype, pin_memory=True)
            out.copy_(t)
            return out.numpy()
        return t.cpu().numpy()

    indptr = np.zeros(N + 1, it)
    indptr[1:] = np.cumsum(per_cell.cpu().numpy())
    indices = host(rows, torch.int32)
    del rows
    data = host(cnt, torch.float64)
    del cnt
    C = sp.csc_matrix((data, indices.astype(it, copy=False), indptr), shape=(P, N), copy=False)
    C.has_sorted_indices = True
    return C


def make_dense_bulk(P, N, rng, dtype=np.float64):
    """Bulk ATAC / DNase-like dense counts (config 4): NegBinomial(mean = a_p * D_c, size = 5)."""
    a = rng.lognormal(0.0, 1.0, P)
    a /= a.sum()
    depth = 2e7 * rng.uniform(0.5, 1.5, N)
    out = np.empty((P, N), dtype=dtype, order="F")
    for c in range(N):
        mean = a * depth[c]
        lam = rng.gamma(5.0, mean / 5.0)
        out[:, c] = rng.poisson(lam)
    zero = np.nonzero(out.sum(axis=1) == 0)[0]
    out[zero, rng.integers(0, N, zero.size)] = 1
    return out


def make_config(name, seed_offset=0, device=None, **override):
    """Returns dict(counts, bias, annoptr, annoidx, cfg) for a named configuration.  `device`
    (e.g. "cuda:0"): sparse counts come from make_counts_device (large configurations)."""
    cfg = dict(CONFIGS[name])
    cfg.update(override)
    P, N, K = cfg["P"], cfg["N"], cfg["K"]
    rng_b = np.random.default_rng(cfg["seed"] + 1)
    bias = make_bias(P, rng_b)
    annoptr, annoidx = make_annotations(P, K, bias, np.random.default_rng(cfg["seed"] + 2), cfg["kind"])
    rng_c = np.random.default_rng(cfg["seed"] + 1000 * seed_offset)
    if cfg["density"] >= 1.0:
        counts = make_dense_bulk(P, N, rng_c)
    elif device is not None:
        counts = make_counts_device(P, N, cfg["density"], rng_c, annoptr, annoidx, device=device)
        cfg["generator"] = "torch device generator (make_counts_device)"
    else:
        counts = make_counts(P, N, cfg["density"], rng_c, annoptr, annoidx)
    return dict(counts=counts, bias=bias, annoptr=annoptr, annoidx=annoidx, cfg=cfg)


def background_matrix(counts, bias, niterations=50, w=0.1, bs=50, seed=20173):
    """A GC-bias x intensity matched background matrix (P x niterations, 1-based, column-major int32)
    as INPUT for both bench arms: numpy statement of the generative form of getBackgroundPeaks
    (whiten (log10 fragments, bias) by the Cholesky factor of their covariance, bs x bs grid, nearest
    grid point, destination bin ~ dnorm(distance, 0, w) over occupied bins, uniform peak inside it).
    The timed GPU sampler (cv_background_peaks) and its parity tests are separate; this exists so the
    CPU arm and the GPU arm of bench.py are fed the identical matrix without a GPU."""
    rng = np.random.default_rng(seed)
    fpp = np.asarray(counts.sum(axis=1)).ravel().astype(np.float64)
    P = fpp.size
    X = np.column_stack([np.log10(fpp), np.asarray(bias, np.float64)])
    L = np.linalg.cholesky(np.cov(X.T))
    T = np.linalg.solve(L, X.T).T
    lo, hi = T.min(axis=0), T.max(axis=0)
    g1, g2 = np.linspace(lo[0], hi[0], bs), np.linspace(lo[1], hi[1], bs)
    b1 = np.clip(np.rint((T[:, 0] - lo[0]) / (g1[1] - g1[0])), 0, bs - 1).astype(np.int64)
    b2 = np.clip(np.rint((T[:, 1] - lo[1]) / (g2[1] - g2[0])), 0, bs - 1).astype(np.int64)
    memb = b1 * bs + b2
    order = np.argsort(memb, kind="stable")
    dens = np.bincount(memb, minlength=bs * bs)
    start = np.concatenate([[0], np.cumsum(dens)])
    occ = np.nonzero(dens)[0]
    xy = np.column_stack([g1[occ // bs], g2[occ % bs]])
    out = np.empty((P, niterations), np.int32, order="F")
    for b in occ:
        src = order[start[b]:start[b + 1]]
        d = np.sqrt(((xy - np.array([g1[b // bs], g2[b % bs]])) ** 2).sum(axis=1))
        pr = np.exp(-0.5 * (d / w) ** 2)
        cdf = np.cumsum(pr / pr.sum())
        cdf[-1] = 1.0
        dest = occ[np.searchsorted(cdf, rng.random((src.size, niterations)), side="right").clip(max=occ.size - 1)]
        within = (rng.random((src.size, niterations)) * dens[dest]).astype(np.int64)
        out[src, :] = order[start[dest] + np.minimum(within, dens[dest] - 1)] + 1
    return out


def random_background(P, B, rng):
    """A stand-in background matrix (1-based, P x B) for tests that only need *some* valid bg."""
    return np.asfortranarray(rng.integers(1, P + 1, size=(P, B)).astype(np.int32))
