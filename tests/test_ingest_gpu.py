"""GPU dataset ingest (qmfb_signals_*: IdIndex + WALSEngine::groupSignals on the device) against the
CPU oracle: dense index assignment and both CSR orientations are BIT-EXACT (north_star: "dataset
indexing and CSR construction are bit-exact")."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _oracle_csr(oracle_lib, rid, cid, val, col_ids):
    n = len(rid)
    perm, rids, rptr = np.empty(n, np.int64), np.empty(n, np.int64), np.empty(n + 1, np.int64)
    nrows = oracle_lib.qmfo_group_signals(np.ascontiguousarray(rid), np.ascontiguousarray(cid), n, perm, rids, rptr)
    col = np.searchsorted(col_ids, cid[perm]).astype(np.int32)
    return rids[:nrows].copy(), rptr[:nrows + 1].copy(), col, val[perm].copy()


@pytest.mark.parametrize("nu,ni,nnz,seed,signed", [(300, 200, 6000, 1, False), (5000, 40, 100000, 2, True), (1, 1, 1, 3, False),
                                                   (70000, 3000, 400000, 4, True), (17, 100000, 50000, 5, False)])
def test_signals_match_oracle(oracle_lib, nu, ni, nnz, seed, signed):
    from qmf_b200 import Signals
    rng = np.random.default_rng(seed)
    # distinct (user, item) cells: the order of exact duplicates is unspecified in the reference (std::sort)
    cells = rng.choice(nu * ni, size=min(nnz, nu * ni), replace=False)
    u, i = cells // ni, cells % ni
    uid = u.astype(np.int64) * 7919 + (-(2 ** 40) if signed else 3)   # raw ids: sparse, possibly negative, > 2^32
    iid = i.astype(np.int64) * 104729 + (2 ** 35 if signed else 0)
    val = rng.integers(1, 6, size=len(cells)).astype(np.float64) + rng.random(len(cells))
    s = Signals(uid, iid, val)
    uids, iids = np.unique(uid), np.unique(iid)
    assert (s.nusers, s.nitems, s.nnz) == (len(uids), len(iids), len(cells))
    assert np.array_equal(s.ids(0), uids) and np.array_equal(s.ids(1), iids)
    for side, (rid, cid, col_ids) in enumerate([(uid, iid, iids), (iid, uid, uids)]):
        rids, rptr, col, v = _oracle_csr(oracle_lib, rid, cid, val, col_ids)
        g_rptr, g_col, g_val, g_order = s.csr(side)
        assert np.array_equal(g_rptr, rptr)
        assert np.array_equal(g_col, col)
        assert np.array_equal(g_val.view(np.int64), v.view(np.int64))
        lens = np.diff(rptr)
        assert np.array_equal(g_order, np.argsort(-lens, kind="stable").astype(np.int32))


def test_signals_duplicates_keep_file_order(oracle_lib):
    from qmf_b200 import Signals
    uid = np.array([5, 5, 2, 5, 2, 2], np.int64)
    iid = np.array([9, 9, 1, 9, 1, 0], np.int64)
    val = np.array([1.0, 2.0, 3.0, 4.0, 5.0, 6.0])
    s = Signals(uid, iid, val)
    rptr, col, v, order = s.csr(0)
    assert rptr.tolist() == [0, 3, 6] and col.tolist() == [0, 1, 1, 2, 2, 2]
    assert v.tolist() == [6.0, 3.0, 5.0, 1.0, 2.0, 4.0]
    rptr, col, v, order = s.csr(1)
    assert rptr.tolist() == [0, 1, 3, 6] and col.tolist() == [0, 0, 0, 1, 1, 1] and v.tolist() == [6.0, 3.0, 5.0, 1.0, 2.0, 4.0]


def test_engine_from_signals_matches_engine_from_host_csr(oracle_lib):
    from qmf_b200 import Signals, WalsEngineHandle, csr_from_coo
    from util import init_factors, uniform_dataset
    u, i, v = uniform_dataset(300, 200, 6000, 77, id_scale=(7, 3))
    uids, urp, uci, uv = csr_from_coo(u, i, v)
    iids, irp, ici, iv = csr_from_coo(i, u, v)
    k = 30
    a = WalsEngineHandle(len(uids), len(iids), k)
    a.set_csr(0, urp, uci, uv)
    a.set_csr(1, irp, ici, iv)
    s = Signals(u, i, v)
    b = WalsEngineHandle(s.nusers, s.nitems, k)
    b.set_signals(s)
    Y0 = init_factors(len(iids), k, seed=5)
    for h in (a, b):
        h.set_factors(1, Y0)
    for _ in range(2):
        la = (a.half_step(0, 40.0, 0.05), a.half_step(1, 40.0, 0.05))
        lb = (b.half_step(0, 40.0, 0.05), b.half_step(1, 40.0, 0.05))
        assert la == lb
    assert np.array_equal(a.get_factors(0), b.get_factors(0)) and np.array_equal(a.get_factors(1), b.get_factors(1))


@pytest.mark.parametrize("name", ["wals_k30", "wals_k64", "wals_k128"])
def test_signals_match_reference_fixture(name):
    """tests/golden/*.npz hold what the REFERENCE's own groupSignals / IdIndex produced for these datasets."""
    from qmf_b200 import Signals
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))
    u, i, v = g["u"], g["i"], g["v"]
    s = Signals(u, i, v)
    for side in (0, 1):
        assert np.array_equal(s.ids(side), g["rid%d" % side])
        rptr, col, val, _ = s.csr(side)
        assert np.array_equal(rptr, g["rp%d" % side]) and np.array_equal(col, g["ci%d" % side])
        # duplicates (same user AND item) may come out in either order from the reference's std::sort
        rows = np.repeat(np.arange(len(rptr) - 1), np.diff(rptr))
        assert sorted(zip(rows, col, val)) == sorted(zip(rows, g["ci%d" % side], g["va%d" % side]))
