"""Shared helpers for the parity tests: seeded synthetic datasets shaped like SURVEY.md §8(d)."""
import numpy as np


def uniform_dataset(nu, ni, nnz, seed, weights=True, id_scale=(1, 1), dup=0):
    """nnz distinct (user,item) cells drawn uniformly without replacement, 1-based ids (optionally
    scaled so that raw ids are not dense), weight ~ U{1..5}; `dup` extra duplicated lines."""
    rng = np.random.default_rng(seed)
    cells = rng.choice(nu * ni, size=nnz, replace=False)
    u = (cells // ni + 1).astype(np.int64) * id_scale[0]
    i = (cells % ni + 1).astype(np.int64) * id_scale[1]
    v = rng.integers(1, 6, size=nnz).astype(np.float64) if weights else np.ones(nnz)
    if dup:
        pick = rng.integers(0, nnz, size=dup)
        u = np.concatenate([u, u[pick]])
        i = np.concatenate([i, i[pick]])
        v = np.concatenate([v, rng.integers(1, 6, size=dup).astype(np.float64)])
    perm = rng.permutation(len(u))
    return u[perm], i[perm], v[perm]


def init_factors(n, k, seed, bound=0.01):
    """seeded U(-bound, bound) rounded to 9 decimals (what a --distribution_file carries)"""
    rng = np.random.default_rng(seed)
    return np.round(rng.uniform(-bound, bound, size=(n, k)), 9)


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
