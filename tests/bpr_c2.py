"""C2-sized BPR distributional parity: shared by tools/bpr_c2_spread.py (reference runs, here on CPU)
and tests/test_bpr_c2_gpu.py (GPU runs).  BASELINE.json configs[1]: bpr on 10k x 5k, nfactors=30, 3 negatives,
use_biases, 10 epochs, with AUC / p@10 - here on a PLANTED problem (a learnable low-rank preference; on
gen_uniform data every model has AUC 0.5 and parity would be vacuous)."""
import numpy as np

NU, NI, K = 10_000, 5_000, 30
NTRAIN, NTEST = 450_000, 50_000
HP = dict(lr=0.05, bias_lambda=1.0, user_lambda=0.025, item_lambda=0.0025, decay=0.9, init_bound=0.01, num_neg=3,
          nepochs=10, eval_num_neg=3, eval_seed=42)   # qmf/bpr.cpp:28-59 defaults


def planted_c2(seed=2024, rank=8):
    """(train_u, train_i, test_u, test_i): the NTRAIN + NTEST user/item cells with the highest planted score"""
    rng = np.random.default_rng(seed)
    A, B = rng.normal(size=(NU, rank)), rng.normal(size=(NI, rank))
    S = A @ B.T + 0.5 * rng.normal(size=(NU, NI))
    n = NTRAIN + NTEST
    flat = np.argpartition(-S.ravel(), n)[:n]
    flat = rng.permutation(flat)
    u, i = (flat // NI).astype(np.int64), (flat % NI).astype(np.int64)
    return u[:NTRAIN], i[:NTRAIN], u[NTRAIN:], i[NTRAIN:]


def rank_metrics(P, Q, b, test_u, test_i):
    """mean over test users of AUC and p@10 of the scores over ALL items against the user's test positives
    (qmf/Engine.cpp:73-96, qmf/metrics/Metrics.cpp; ties are measure-zero on trained factors)"""
    order = np.argsort(test_u, kind="stable")
    tu, ti = test_u[order], test_i[order]
    users, start = np.unique(tu, return_index=True)
    end = np.append(start[1:], len(tu))
    aucs, p10 = [], []
    for blk in range(0, len(users), 512):
        us = users[blk:blk + 512]
        S = P[us] @ Q.T
        if b is not None:
            S += b[None, :]
        ranks = np.argsort(np.argsort(S, axis=1), axis=1)          # 0 = lowest score
        top = np.argpartition(-S, 10, axis=1)[:, :10]
        for r, (s, e) in enumerate(zip(start[blk:blk + 512], end[blk:blk + 512])):
            pos = np.unique(ti[s:e])
            npos, nneg = len(pos), NI - len(pos)
            rk = np.sort(ranks[r, pos])
            # negatives below each positive = rank - (positives below it)
            aucs.append(float((rk - np.arange(npos)).sum()) / (npos * nneg))
            p10.append(len(np.intersect1d(top[r], pos)) / 10.0)
    return float(np.mean(aucs)), float(np.mean(p10))


def dense_index(ids):
    """IdIndex (qmf/utils/IdIndex.h): dense idx in order of first appearance.  Returns (idx per element, raw id per idx)."""
    uniq, first, inv = np.unique(ids, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")           # unique values by first appearance
    rank = np.empty(len(uniq), np.int64)
    rank[order] = np.arange(len(uniq))
    return rank[inv], uniq[order]


def pos_csr(du, di, nu):
    """per-user sorted positive sets (BPREngine::init userPositives_)"""
    o = np.lexsort((di, du))
    ptr = np.zeros(nu + 1, np.int64)
    np.cumsum(np.bincount(du, minlength=nu), out=ptr[1:])
    return ptr, np.ascontiguousarray(di[o])
