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
    # power-law (powerlaw_shard_torch): user degree by a rank-size law with exponent 1/1.1 clipped to
    # [1, 1e5], item popularity Zipf(1.0); nnz is the number of DRAWS - heavy users hit the head items repeatedly,
    # so the distinct cells are fewer (722 M of 1 B at full size)
    "c5": (10_000_000, 1_000_000, 1_000_000_000, 128),
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


def powerlaw_degrees(nusers, nnz, seed, exponent=1.0 / 1.1, dmax=100_000):
    """Number of draws per user (int64 [nusers], sum ~= nnz): rank-size law d(j) = clip(C j^-exponent, 1, dmax)
    with C solved for the total, assigned to user indices by a seeded permutation.  Host numpy, identical on
    every rank ("user degree ~Zipf(1.1) clipped to [1, 1e5]")."""
    w = np.arange(1, nusers + 1, dtype=np.float64) ** -exponent
    lo, hi = 0.0, float(nnz)
    for _ in range(60):
        c = 0.5 * (lo + hi)
        if np.clip(c * w, 1.0, dmax).sum() < nnz:
            lo = c
        else:
            hi = c
    deg = np.clip(np.rint(hi * w), 1, dmax).astype(np.int64)
    return deg[np.random.default_rng(seed).permutation(nusers)]


def _powerlaw_chunks(deg, target):
    """cut the users into contiguous chunks of about `target` draws: list of (u0, u1)"""
    cum = np.concatenate([[0], np.cumsum(deg)])
    cuts, u = [0], 0
    while u < len(deg):
        u = int(np.searchsorted(cum, cum[u] + target, side="left"))
        u = min(max(u, cuts[-1] + 1), len(deg))
        cuts.append(u)
    return list(zip(cuts[:-1], cuts[1:]))


def _powerlaw_chunk_keys(deg_dev, u0, u1, nitems, seed, chunk_index, device, mul, add):
    """distinct cells (u * nitems + item), ascending, of the users [u0, u1): deg[u] draws of an item with
    P(popularity rank i) ~ 1/i (inverse CDF floor((ni + 1)^U) - 1), popularity rank -> item idx by the
    affine bijection (mul * i + add) mod nitems so that blockbusters are scattered over the id range"""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed * 1_000_003 + chunk_index)
    d = deg_dev[u0:u1]
    n = int(d.sum().item())
    u = torch.repeat_interleave(torch.arange(u0, u1, device=device, dtype=torch.int64), d, output_size=n)
    x = torch.rand(n, generator=g, device=device, dtype=torch.float64)
    r = torch.exp(x * float(np.log(nitems + 1.0))).to(torch.int64).clamp_(1, nitems) - 1
    del x
    item = (r * mul + add) % nitems
    del r
    keys = torch.unique(u * nitems + item)
    return keys


def _cell_values(keys):
    """weight of a cell in {1..5}, a pure function of the cell so both orientations agree without storing it"""
    import torch
    h = (keys * 2654435761 + 0x9E3779B9) & 0x7FFFFFFF
    return ((h >> 7) % 5 + 1).to(dtype=torch.float64)


def powerlaw_shard_torch(nusers, nitems, nnz, seed, device, rank=0, world=1, chunk_draws=100_000_000):
    """C5-shaped problem WITHOUT materialising it on any one GPU: every rank streams the same seeded chunks of
    users twice - pass 1 counts signals per user and per item (-> nnz-balanced contiguous row ranges, identical
    on all ranks), pass 2 keeps the user rows and the item rows of THIS rank only.
    Returns dict(ranges=(user_ranges, item_ranges), csr_user=(row_ptr, col, val), csr_item=(...), nnz=total
    distinct cells, test_items=int32 [nusers]: one held-out test item per user from the same popularity law)."""
    import torch
    from .wals_dist import balanced_row_ranges
    deg = powerlaw_degrees(nusers, nnz, seed)
    chunks = _powerlaw_chunks(deg, chunk_draws)
    deg_dev = torch.from_numpy(deg).to(device)
    mul = 1
    for cand in (7_368_787, 982_451_653, 15_485_863, 104_729, 7919, 1):
        if np.gcd(cand % nitems, nitems) == 1 and cand % nitems != 0:
            mul = cand % nitems
            break
    add = 12_345 % nitems
    ucnt = torch.zeros(nusers, dtype=torch.int64, device=device)
    icnt = torch.zeros(nitems, dtype=torch.int64, device=device)
    for ci, (u0, u1) in enumerate(chunks):
        keys = _powerlaw_chunk_keys(deg_dev, u0, u1, nitems, seed, ci, device, mul, add)
        u = keys // nitems
        ucnt += torch.bincount(u, minlength=nusers)
        icnt += torch.bincount(keys - u * nitems, minlength=nitems)
        del keys, u
    urp = torch.zeros(nusers + 1, dtype=torch.int64, device=device)
    urp[1:] = torch.cumsum(ucnt, 0)
    irp = torch.zeros(nitems + 1, dtype=torch.int64, device=device)
    irp[1:] = torch.cumsum(icnt, 0)
    total = int(urp[-1].item())
    uranges, iranges = balanced_row_ranges(urp, world), balanced_row_ranges(irp, world)
    (ub, ue), (ib, ie) = uranges[rank], iranges[rank]
    ucol, uval, iitem, iuser, ival = [], [], [], [], []
    for ci, (u0, u1) in enumerate(chunks):
        keys = _powerlaw_chunk_keys(deg_dev, u0, u1, nitems, seed, ci, device, mul, add)
        u = keys // nitems
        it = keys - u * nitems
        if u1 > ub and u0 < ue:
            m = (u >= ub) & (u < ue)
            ucol.append(it[m].to(torch.int32))
            uval.append(_cell_values(keys[m]))
        m = (it >= ib) & (it < ie)
        iitem.append(it[m])
        iuser.append(u[m].to(torch.int32))
        ival.append(_cell_values(keys[m]))
        del keys, u, it, m
    lurp = (urp[ub:ue + 1] - urp[ub]).contiguous()
    csr_user = (lurp, torch.cat(ucol) if ucol else torch.zeros(0, dtype=torch.int32, device=device),
                torch.cat(uval) if uval else torch.zeros(0, dtype=torch.float64, device=device))
    iitem, iuser, ival = torch.cat(iitem), torch.cat(iuser), torch.cat(ival)
    perm = torch.argsort(iitem, stable=True)  # users stay ascending within an item
    lirp = (irp[ib:ie + 1] - irp[ib]).contiguous()
    csr_item = (lirp, iuser[perm].contiguous(), ival[perm].contiguous())
    del iitem, perm
    g = torch.Generator(device=device)
    g.manual_seed(seed * 1_000_003 + 999_983)
    x = torch.rand(nusers, generator=g, device=device, dtype=torch.float64)
    r = torch.exp(x * float(np.log(nitems + 1.0))).to(torch.int64).clamp_(1, nitems) - 1
    test_items = ((r * mul + add) % nitems).to(torch.int32)
    return dict(ranges=(uranges, iranges), csr_user=csr_user, csr_item=csr_item, nnz=total, test_items=test_items,
                max_item_len=int(icnt.max().item()), max_user_len=int(ucnt.max().item()))


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
