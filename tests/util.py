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
    """global metric max|a - b| / max|b| - for matrices that are not sets of independent rows (Gram)"""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def rel_err_rows(a, b):
    """The factor tolerance of north_star / SURVEY.md §7: max over rows r of
    ||a_r - b_r||_inf / ||b_r||_inf.  A row that is exactly zero in the reference (a row without
    signals and a zero right-hand side) must be exactly zero here as well."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.ndim == 1:
        a, b = a[None, :], b[None, :]
    assert a.shape == b.shape
    num = np.abs(a - b).max(axis=1)
    den = np.abs(b).max(axis=1)
    zero = den == 0.0
    if np.any(zero & (num != 0.0)):
        return float("inf")
    return float((num[~zero] / den[~zero]).max()) if np.any(~zero) else 0.0


def planted_bpr_dataset(nu, ni, per_user, n_test, seed, rank=8, noise=1.0):
    """Implicit-feedback pairs from a planted rank-`rank` preference (SURVEY.md §8d "planted"): user u
    likes the `per_user` items with the largest  a_u . b_i + noise * Gumbel;  `n_test` of them (random) go
    to the test split.  Returns (train_u, train_i, test_u, test_i) as 0-based int64 arrays in a shuffled
    line order.  On uniform data both implementations sit at AUC 0.5, which cannot fail a broken kernel."""
    rng = np.random.default_rng(seed)
    A, B = rng.normal(size=(nu, rank)), rng.normal(size=(ni, rank))
    tu, ti, eu, ei = [], [], [], []
    for u0 in range(0, nu, 1000):
        S = A[u0:u0 + 1000] @ B.T + noise * rng.gumbel(size=(min(1000, nu - u0), ni))
        top = np.argpartition(-S, per_user, axis=1)[:, :per_user]
        for r in range(top.shape[0]):
            items = rng.permutation(top[r])
            eu.append(np.full(n_test, u0 + r)); ei.append(items[:n_test])
            tu.append(np.full(per_user - n_test, u0 + r)); ti.append(items[n_test:])
    tu, ti, eu, ei = (np.concatenate(x).astype(np.int64) for x in (tu, ti, eu, ei))
    p, q = rng.permutation(len(tu)), rng.permutation(len(eu))
    return tu[p], ti[p], eu[q], ei[q]
