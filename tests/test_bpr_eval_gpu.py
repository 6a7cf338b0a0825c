"""GPU parity of BPR SGD (single steps exact-order replay, eval loss, Hogwild epochs statistically)
and of the ranking evaluation (bit-exact integer rank statistics and metric values) against the
CPU oracle, through the C ABI."""
import numpy as np
import pytest

from util import rel_err, rel_err_rows

pytestmark = pytest.mark.gpu


def _planted(nu, ni, npairs, seed, rank=6):
    """implicit-feedback pairs from a planted low-rank preference so that AUC >> 0.5 is learnable"""
    rng = np.random.default_rng(seed)
    A, B = rng.normal(size=(nu, rank)), rng.normal(size=(ni, rank))
    S = A @ B.T + 0.3 * rng.normal(size=(nu, ni))
    flat = np.argsort(-S, axis=None)[: npairs * 2]
    pick = rng.choice(flat, size=npairs, replace=False)
    u, i = (pick // ni).astype(np.int32), (pick % ni).astype(np.int32)
    return u, i, S


@pytest.mark.parametrize("k,biases", [(30, True), (30, False), (64, True), (100, False), (7, True), (128, True),
                                      (160, False), (256, True), (96, True), (200, False)])
def test_update_triplets_match_oracle(oracle_lib, k, biases):
    from qmf_b200.bpr import BprEngineHandle
    import oracle
    rng = np.random.default_rng(k)
    nu, ni, n = 40, 30, 400
    P = rng.uniform(-0.3, 0.3, (nu, k))
    Q = rng.uniform(-0.3, 0.3, (ni, k))
    b = rng.uniform(-0.3, 0.3, ni) if biases else None
    u = rng.integers(0, nu, n).astype(np.int32)
    i = rng.integers(0, ni, n).astype(np.int32)
    j = ((i + rng.integers(1, ni, n)) % ni).astype(np.int32)
    h = BprEngineHandle(nu, ni, k, use_biases=biases)
    h.set_factors(0, P)
    h.set_factors(1, Q)
    if biases:
        h.set_biases(b)
    lr, lu, li, lb = 0.05, 0.025, 0.0025, 1.0
    h.update_triplets(u, i, j, lr, lu, li, lb)
    Po, Qo, bo = P.copy(), Q.copy(), (b.copy() if biases else None)
    for t in range(n):
        oracle_lib.qmfo_bpr_update(Po, Qo, oracle.ptr(bo), k, int(u[t]), int(i[t]), int(j[t]), lr, lu, li, lb)
    # 400 dependent steps; only the dot-product summation order differs (warp tree vs sequential)
    assert rel_err_rows(h.get_factors(0), Po) < 1e-12
    assert rel_err_rows(h.get_factors(1), Qo) < 1e-12
    if biases:
        assert rel_err(h.get_biases(), bo) < 1e-12


@pytest.mark.parametrize("k,biases,nthreads", [(30, True, 16), (64, False, 7), (128, True, 1)])
def test_eval_loss_matches_oracle(oracle_lib, k, biases, nthreads):
    from qmf_b200.bpr import BprEngineHandle
    import oracle
    rng = np.random.default_rng(100 + k)
    nu, ni, n = 300, 200, 5003
    P, Q = rng.uniform(-1, 1, (nu, k)), rng.uniform(-1, 1, (ni, k))
    b = rng.uniform(-1, 1, ni) if biases else None
    u, i, j = (rng.integers(0, m, n) for m in (nu, ni, ni))
    h = BprEngineHandle(nu, ni, k, use_biases=biases)
    h.set_factors(0, P)
    h.set_factors(1, Q)
    if biases:
        h.set_biases(b)
    got = h.eval_loss(u, i, j, nthreads)
    want = oracle_lib.qmfo_bpr_eval_loss(P, Q, oracle.ptr(b), k, u.astype(np.int64), i.astype(np.int64), j.astype(np.int64), n,
                                         nthreads)
    assert abs(got - want) <= 1e-12 * abs(want)


def test_hogwild_epochs_learn_planted_preferences(oracle_lib):
    """Hogwild is nondeterministic in both implementations: parity is statistical.  After 10 epochs
    on planted data the GPU run must reach the same train-loss level as the sequential CPU oracle
    (within 3 % relative) and a clearly better-than-chance AUC close to the oracle's."""
    from qmf_b200.bpr import BprEngineHandle
    from qmf_b200.evalrank import eval_rank, labels_to_csr, user_metrics
    import oracle
    nu, ni, npairs, k = 600, 400, 24000, 30
    u, i, S = _planted(nu, ni, npairs, seed=5)
    rng = np.random.default_rng(1)
    P0, Q0, b0 = (rng.uniform(-0.01, 0.01, s) for s in ((nu, k), (ni, k), ni))
    lr0, lu, li, lb, decay, num_neg = 0.05, 0.025, 0.0025, 1.0, 0.9, 3
    # fixed evaluation triplets (sampled with the oracle's mt19937 restatement, BPREngine.cpp:85-87)
    order = np.lexsort((i, u))
    us, is_ = u[order].astype(np.int64), i[order].astype(np.int64)
    pos_ptr = np.zeros(nu + 1, np.int64)
    np.cumsum(np.bincount(us, minlength=nu), out=pos_ptr[1:])
    neg = np.zeros(npairs * 3, np.int64)
    oracle_lib.qmfo_bpr_sample_negatives(u.astype(np.int64), npairs, 3, ni, pos_ptr, is_, 42, neg)
    eu, ei, ej = np.repeat(u, 3), np.repeat(i, 3), neg.astype(np.int32)
    assert not any(ej[t] in set(is_[pos_ptr[eu[t]]:pos_ptr[eu[t] + 1]]) for t in range(0, len(ej), 97))

    h = BprEngineHandle(nu, ni, k, use_biases=True)
    h.set_data(u, i)
    h.set_factors(0, P0); h.set_factors(1, Q0); h.set_biases(b0)
    Po, Qo, bo = P0.copy(), Q0.copy(), b0.copy()
    lr = lr0
    g = np.random.default_rng(9)
    for epoch in range(10):
        n_upd = h.epoch(lr, lu, li, lb, num_neg, seed=1234, epoch=epoch, shuffle=True)
        assert n_upd == npairs * num_neg
        for p in g.permutation(npairs):           # sequential CPU oracle, its own negatives
            for _ in range(num_neg):
                while True:
                    jn = int(g.integers(0, ni))
                    lo, hi = pos_ptr[u[p]], pos_ptr[u[p] + 1]
                    if jn not in is_[lo:hi]:
                        break
                oracle_lib.qmfo_bpr_update(Po, Qo, oracle.ptr(bo), k, int(u[p]), int(i[p]), jn, lr, lu, li, lb)
        lr *= decay
    loss_gpu = h.eval_loss(eu, ei, ej, 16)
    loss_cpu = oracle_lib.qmfo_bpr_eval_loss(Po, Qo, oracle.ptr(bo), k, eu.astype(np.int64), ei.astype(np.int64),
                                             ej.astype(np.int64), len(eu), 16)
    assert loss_gpu < 0.6 and abs(loss_gpu - loss_cpu) < 0.03 * loss_cpu, (loss_gpu, loss_cpu)

    # ranking quality against the planted preference (top 5 % of S per user as "test positives")
    thr = np.quantile(S, 0.95, axis=1)
    rows = [np.flatnonzero(S[t] >= thr[t]).astype(np.int32) for t in range(nu)]
    lp, lit = labels_to_csr(rows)
    tu = np.arange(nu, dtype=np.int32)

    def mean_auc(P, Q, b):
        cnt, _ = eval_rank(P, Q, b, tu, lp, lit)
        return np.mean([user_metrics(cnt[lp[t] + t: lp[t + 1] + t + 1], ni, ["auc"])["auc"] for t in range(nu)])

    auc_gpu = mean_auc(h.get_factors(0), h.get_factors(1), h.get_biases())
    auc_cpu = mean_auc(Po, Qo, bo)
    assert auc_gpu > 0.7 and abs(auc_gpu - auc_cpu) < 0.03, (auc_gpu, auc_cpu)


@pytest.mark.parametrize("nu,ni,k,biases", [(50, 333, 30, True), (20, 1000, 64, False), (8, 77, 128, True), (5, 40, 3, False),
                                            (6, 90, 200, True), (1, 17, 16, False), (9, 2100, 17, True)])
def test_rank_statistics_bit_exact(oracle_lib, nu, ni, k, biases):
    """Scores bit-identical to Engine::computeTestScores, therefore identical ranking: AUC (exact
    replay), AP, P@k and R@k must EQUAL the oracle's values, ties included."""
    from qmf_b200.evalrank import eval_rank, labels_to_csr, user_metrics
    import oracle
    rng = np.random.default_rng(ni)
    U = np.round(rng.uniform(-1, 1, (nu, k)), 2)      # coarse values -> many exactly tied scores
    V = np.round(rng.uniform(-1, 1, (ni, k)), 1)
    V[ni // 2:] = V[: ni - ni // 2]                  # duplicated items: guaranteed ties
    b = np.round(rng.uniform(-1, 1, ni), 1) if biases else None
    if biases:
        b[ni // 2:] = b[: ni - ni // 2]
    tu = rng.permutation(nu)[: max(1, nu - 2)].astype(np.int32)
    rows = [np.sort(rng.choice(ni, size=int(rng.integers(1, min(40, ni - 1))), replace=False)).astype(np.int32) for _ in tu]
    lp, li = labels_to_csr(rows)
    cnt, pos_scores = eval_rank(U, V, b, tu, lp, li)
    scores = np.zeros((len(tu), ni))
    oracle_lib.qmfo_compute_test_scores(U, V, oracle.ptr(b), ni, k, tu.astype(np.int64), len(tu), scores)
    names = ["auc", "ap", "p@1", "p@10", "r@10", "r@5"]
    for t in range(len(tu)):
        labels = np.zeros(ni)
        labels[rows[t]] = 1.0
        c = cnt[lp[t] + t: lp[t + 1] + t + 1]
        assert c.sum() == ni - len(rows[t])
        assert np.array_equal(np.sort(scores[t][rows[t]]), pos_scores[lp[t]:lp[t + 1]])   # bit-identical scores
        got = user_metrics(c, ni, names)
        for name in names:
            kind, kk = oracle.metric_kind(name)
            want = oracle_lib.qmfo_metric_one(kind, kk, labels, scores[t], ni)
            assert got[name] == want, (t, name, got[name], want)


def test_non_finite_gradient_is_reported():
    """CHECK(std::isfinite(e)) in BPREngine::update (qmf/bpr/BPREngine.cpp:184-185) aborts the
    reference; the C ABI returns QMFB_ERR_NOT_FINITE."""
    from qmf_b200 import capi
    from qmf_b200.bpr import BprEngineHandle
    h = BprEngineHandle(4, 6, 8)
    P = np.full((4, 8), 1e200)
    Q = np.zeros((6, 8))
    Q[1] = 1e200                    # p_u . (q_1 - q_j) overflows to +inf -> exp(inf) = inf -> e = 0 is finite;
    Q[2] = np.nan                   # a NaN row makes e NaN
    h.set_factors(0, P)
    h.set_factors(1, Q)
    with pytest.raises(capi.QmfbError) as e:
        h.update_triplets([0], [2], [3], 0.05, 0.025, 0.0025, 1.0)
    assert e.value.code == capi.ERR_NOT_FINITE


def test_unsupported_shapes_fail_loudly():
    from qmf_b200 import capi
    from qmf_b200.bpr import BprEngineHandle
    from qmf_b200.wals import WalsEngineHandle
    with pytest.raises(capi.QmfbError) as e:
        WalsEngineHandle(10, 10, 257)
    assert e.value.code == -5 and "nfactors" in str(e.value)
    with pytest.raises(capi.QmfbError):
        BprEngineHandle(10, 10, 300)
    h = BprEngineHandle(3, 3, 4)
    with pytest.raises(capi.QmfbError):
        h.get_biases()              # "can't access bias when withBiases = false" (qmf/FactorData.h:45-48)
    with pytest.raises(capi.QmfbError):
        h.set_data([0, 5], [1, 1])  # user idx out of range
