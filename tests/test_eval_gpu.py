"""The DMMA scoring kernel with exact re-check (eval_kernels.cuh) against the CPU oracle's exact-order
scores: the integer bucket counts must be EQUAL - on continuous random factors (no ties: the fast path
decides everything), with massive ties (the re-score path decides everything), with users that have
thousands of positives (global-memory buckets, device-wide segmented sort) and through the engines'
resident-factor entry points, single and sharded."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _counts_from_scores(scores, rows):
    """bucket counts of the negatives against the ascending positives, as Metrics.cpp's order implies"""
    out, ps = [], []
    for t, pos in enumerate(rows):
        s = scores[t]
        sp = np.sort(s[pos])
        neg = np.ones(len(s), bool)
        neg[pos] = False
        c = np.bincount(np.searchsorted(sp, s[neg], side="left"), minlength=len(sp) + 1).astype(np.int32)
        out.append(c)
        ps.append(sp)
    return np.concatenate(out), np.concatenate(ps)


def _oracle_scores(oracle_lib, U, V, b, tu):
    import oracle
    scores = np.zeros((len(tu), V.shape[0]))
    oracle_lib.qmfo_compute_test_scores(np.ascontiguousarray(U), np.ascontiguousarray(V), oracle.ptr(b), V.shape[0], U.shape[1],
                                        tu.astype(np.int64), len(tu), scores)
    return scores


@pytest.mark.parametrize("nu,ni,k,biases,maxpos", [(700, 9000, 128, False, 30), (300, 20011, 30, True, 8), (130, 4000, 200, True, 50),
                                                   (65, 700, 64, False, 5), (1, 100, 7, True, 3), (900, 333, 96, False, 12)])
def test_counts_equal_the_exact_order_scores(oracle_lib, nu, ni, k, biases, maxpos):
    from qmf_b200.evalrank import eval_rank, labels_to_csr
    rng = np.random.default_rng(nu + ni + k)
    U, V = rng.normal(size=(nu, k)), rng.normal(size=(ni, k))
    b = rng.normal(size=ni) if biases else None
    tu = rng.permutation(nu).astype(np.int32)
    rows = [np.sort(rng.choice(ni, size=int(rng.integers(1, maxpos + 1)), replace=False)).astype(np.int32) for _ in tu]
    lp, li = labels_to_csr(rows)
    cnt, pos_scores = eval_rank(U, V, b, tu, lp, li)
    want_cnt, want_ps = _counts_from_scores(_oracle_scores(oracle_lib, U, V, b, tu), rows)
    assert np.array_equal(pos_scores, want_ps)        # exact-order scores of the positives, bit for bit
    assert np.array_equal(cnt, want_cnt)


def test_all_scores_tied_goes_through_the_exact_path(oracle_lib):
    """a zero user row and duplicated item rows: every item is within the error bound of a positive"""
    from qmf_b200.evalrank import eval_rank, labels_to_csr
    rng = np.random.default_rng(4)
    nu, ni, k = 70, 1500, 30
    U, V = np.round(rng.normal(size=(nu, k)), 1), np.round(rng.normal(size=(ni, k)), 1)
    U[::3] = 0.0
    V[ni // 3:2 * ni // 3] = V[:ni // 3]
    tu = np.arange(nu, dtype=np.int32)
    rows = [np.sort(rng.choice(ni, size=6, replace=False)).astype(np.int32) for _ in tu]
    lp, li = labels_to_csr(rows)
    cnt, _ = eval_rank(U, V, None, tu, lp, li)
    want_cnt, _ = _counts_from_scores(_oracle_scores(oracle_lib, U, V, None, tu), rows)
    assert np.array_equal(cnt, want_cnt)


def test_users_with_thousands_of_positives(oracle_lib):
    """more positives than the shared-memory bucket table (2 048 per 64 users) and than one CTA's sort
    (4 096): buckets in global memory, positives sorted by the device-wide segmented sort.  The reference
    has no limit on positives per test user."""
    from qmf_b200.evalrank import eval_rank, labels_to_csr
    rng = np.random.default_rng(11)
    nu, ni, k = 5, 12000, 40
    U, V = rng.normal(size=(nu, k)), rng.normal(size=(ni, k))
    tu = np.arange(nu, dtype=np.int32)
    rows = [np.sort(rng.choice(ni, size=n, replace=False)).astype(np.int32) for n in (3000, 5000, 10, 1, 11999)]
    lp, li = labels_to_csr(rows)
    cnt, pos_scores = eval_rank(U, V, None, tu, lp, li)
    want_cnt, want_ps = _counts_from_scores(_oracle_scores(oracle_lib, U, V, None, tu), rows)
    assert np.array_equal(pos_scores, want_ps)
    assert np.array_equal(cnt, want_cnt)


def test_engines_evaluate_their_resident_factors():
    """qmfb_wals_eval_rank / qmfb_wals_sharded_eval_rank / qmfb_bpr_eval_rank == the host-buffer call"""
    from qmf_b200 import WalsEngineHandle
    from qmf_b200.bpr import BprEngineHandle
    from qmf_b200.evalrank import eval_rank, labels_to_csr
    from qmf_b200.wals import ShardedWalsHandle
    rng = np.random.default_rng(2)
    nu, ni = 500, 3000
    tu = rng.permutation(nu)[:333].astype(np.int32)
    rows = [np.sort(rng.choice(ni, size=int(rng.integers(0, 9)), replace=False)).astype(np.int32) for _ in tu]
    rows[0] = np.zeros(0, np.int32)               # a test user without positive labels
    lp, li = labels_to_csr(rows)
    for k in (30, 128):
        U, V = rng.normal(size=(nu, k)), rng.normal(size=(ni, k))
        want = eval_rank(U, V, None, tu, lp, li)
        h = WalsEngineHandle(nu, ni, k)
        h.set_factors(0, U)
        h.set_factors(1, V)
        got = h.eval_rank(tu, lp, li)
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
        h.close()
        sw = ShardedWalsHandle(nu, ni, k, [0, 0, 0])
        sw.set_factors(0, U)
        sw.set_factors(1, V)
        got = sw.eval_rank(tu, lp, li)
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
        sw.close()
        b = rng.normal(size=ni)
        want = eval_rank(U, V, b, tu, lp, li)
        hb = BprEngineHandle(nu, ni, k, use_biases=True)
        hb.set_factors(0, U)
        hb.set_factors(1, V)
        hb.set_biases(b)
        got = hb.eval_rank(tu, lp, li)
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
        hb.close()
