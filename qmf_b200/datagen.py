"""Seeded synthetic datasets for the BASELINE.json configs.  The reference ships no
dataset generator (qmf/gen_uniform.cpp only writes initial factors), so the shapes are defined
here: `uniform` = nnz distinct (user, item) cells drawn uniformly, weights U{1..5}.

`uniform_csr_torch` builds both orientations directly on a device (torch is plumbing here: RNG,
sort, prefix sums) so that the 100 M-nnz configs are not bottlenecked on host text parsing.
"""
import numpy as np

CONFIGS = {
    # name: (nusers, nitems, nnz, nfactors)
    "c1": (10_000, 5_000, 500_000, 30),
    "c3": (138_000, 27_000, 20_000_000, 64),
    "c4": (480_000, 17_800, 100_000_000, 128),
}


def uniform_csr_torch(nusers, nitems, nnz, seed, device):
    """Returns (csr_user, csr_item), each (row_ptr int64, col int32, val f64) on `device`, in the
    reference's order (rows ascending, entries ascending within a row; WALSEngine.cpp:156-163).
    Identical for identical (shape, seed, device type)."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    ncell = nusers * nitems
    draw = int(nnz * 1.02) + 1024
    cells = torch.randint(0, ncell, (draw,), generator=g, device=device, dtype=torch.int64)
    cells = torch.unique(cells)  # sorted ascending == (user, item) lexicographic
    if cells.numel() < nnz:
        raise RuntimeError("not enough distinct cells drawn; increase the oversampling factor")
    if cells.numel() > nnz:
        keep = torch.randperm(cells.numel(), generator=g, device=device)[:nnz]
        cells = cells[torch.sort(keep).values]
    u = cells // nitems
    i = cells - u * nitems
    del cells
    val = torch.randint(1, 6, (nnz,), generator=g, device=device).to(torch.float64)
    urp = torch.zeros(nusers + 1, dtype=torch.int64, device=device)
    urp[1:] = torch.cumsum(torch.bincount(u, minlength=nusers), 0)
    csr_user = (urp, i.to(torch.int32), val)
    # item orientation: stable sort by item keeps users ascending within an item
    perm = torch.argsort(i, stable=True)
    irp = torch.zeros(nitems + 1, dtype=torch.int64, device=device)
    irp[1:] = torch.cumsum(torch.bincount(i, minlength=nitems), 0)
    csr_item = (irp, u[perm].to(torch.int32), val[perm])
    return csr_user, csr_item


def uniform_row_sample(nrows_total, ncols, p, nsample, seed):
    """A statistically exact sample of `nsample` rows of the uniform dataset without building it:
    each row has Binomial(ncols, p) distinct columns chosen uniformly.  Returns local CSR
    (row_ptr int64, col int32, val f64)."""
    rng = np.random.default_rng(seed)
    lens = rng.binomial(ncols, p, size=nsample).astype(np.int64)
    lens = np.maximum(lens, 1)
    row_ptr = np.zeros(nsample + 1, dtype=np.int64)
    np.cumsum(lens, out=row_ptr[1:])
    col = np.empty(int(row_ptr[-1]), dtype=np.int32)
    for r in range(nsample):
        # distinct columns: sample with replacement, unique, top up (p is small so collisions are rare)
        n = int(lens[r])
        c = np.unique(rng.integers(0, ncols, size=n))
        while c.size < n:
            c = np.unique(np.concatenate([c, rng.integers(0, ncols, size=n - c.size)]))
        col[row_ptr[r]:row_ptr[r + 1]] = c
    val = rng.integers(1, 6, size=col.size).astype(np.float64)
    return row_ptr, col, val


def init_item_factors(nitems, k, seed, bound=0.01):
    """seeded stand-in for qmf/gen_uniform.cpp: U(-bound, bound), 9 decimals, (item idx, factor) order"""
    rng = np.random.default_rng(seed)
    return np.round(rng.uniform(-bound, bound, size=(nitems, k)), 9)
