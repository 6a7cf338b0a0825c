#!/usr/bin/env python
"""Generates tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref/libqmf_ref.so,
compiled from /root/reference by `make -C oracle ref`) on seeded inputs.  Run it where
/root/reference is mounted:   OMP_NUM_THREADS=1 python tests/golden/make_golden.py
The fixtures are what travels to the GPU box; the reference itself does not."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("OMP_NUM_THREADS", "1")
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")

import oracle  # noqa: E402
from util import init_factors, uniform_dataset  # noqa: E402

R = oracle.ref()
R.ref_set_min_log_level(2)


def wals_case(name, nu, ni, nnz, k, nepochs, seed, dup=0, alpha=40.0, lam=0.05, nthreads=4):
    u, i, v = uniform_dataset(nu, ni, nnz, seed, id_scale=(7, 3), dup=dup)
    h = R.ref_wals_create(k, nepochs, lam, alpha, nthreads, None, 0, 0, 42)
    R.ref_wals_init(h, u, i, v, len(u))
    NU, NI = R.ref_wals_nusers(h), R.ref_wals_nitems(h)
    Y0 = init_factors(NI, k, seed + 1)
    R.ref_wals_set_factors(h, 1, Y0)
    n = len(u)
    out = dict(u=u, i=i, v=v, k=k, alpha=alpha, lam=lam, nthreads=nthreads, Y0=Y0)
    for side, nrows in ((0, NU), (1, NI)):
        rp, rid = np.zeros(nrows + 1, np.int64), np.zeros(nrows, np.int64)
        ci, cid, va = np.zeros(n, np.int32), np.zeros(n, np.int64), np.zeros(n)
        R.ref_wals_csr(h, side, rp, rid, ci, cid, va)
        out.update({"rp%d" % side: rp, "rid%d" % side: rid, "ci%d" % side: ci, "cid%d" % side: cid, "va%d" % side: va})
    losses, X, Y = [], [], []
    for _ in range(nepochs):
        lu = R.ref_wals_half_step(h, 0)
        li = R.ref_wals_half_step(h, 1)
        losses.append((lu, li))
        Xe, Ye = np.zeros((NU, k)), np.zeros((NI, k))
        R.ref_wals_get_factors(h, 0, Xe)
        R.ref_wals_get_factors(h, 1, Ye)
        X.append(Xe)
        Y.append(Ye)
    R.ref_wals_destroy(h)
    out.update(losses=np.array(losses), X=np.array(X), Y=np.array(Y))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "nusers", NU, "nitems", NI, "losses", losses[-1])


def bpr_case(name, nu, ni, npairs, k, nepochs, seed, biases=True):
    rng = np.random.default_rng(seed)
    cells = rng.choice(nu * ni, npairs, replace=False)
    u = (cells // ni + 1).astype(np.int64) * 5
    i = (cells % ni + 1).astype(np.int64) * 2
    v = np.ones(npairs)
    v[:: 17] = 0.0          # value < 1 lines are ignored (BPREngine.cpp:70-72)
    ntest = npairs // 10
    tu, ti, tv = u[-ntest:].copy(), i[-ntest:].copy(), np.ones(ntest)
    u, i, v = u[:-ntest], i[:-ntest], v[:-ntest]
    cfg = dict(k=k, lr=0.05, bias_lambda=1.0, user_lambda=0.025, item_lambda=0.0025, decay=0.9, biases=int(biases),
               bound=0.01, num_neg=3, eval_num_neg=3, eval_seed=42, nthreads=4, gen_seed=7)
    # nepochs = 1 per optimize() call so that the per-epoch state can be recorded
    h = R.ref_bpr_create(k, 1, cfg["lr"], cfg["bias_lambda"], cfg["user_lambda"], cfg["item_lambda"], cfg["decay"],
                         cfg["biases"], cfg["bound"], cfg["num_neg"], 1, 1, cfg["eval_num_neg"], cfg["eval_seed"],
                         cfg["nthreads"], b"auc,ap,p@5,r@5", 0, 1, cfg["gen_seed"])
    R.ref_bpr_init(h, u, i, v, len(u))
    R.ref_bpr_init_test(h, tu, ti, tv, len(tu))
    NU, NI, ND = R.ref_bpr_nusers(h), R.ref_bpr_nitems(h), R.ref_bpr_ndata(h)
    uid, iid = np.zeros(NU, np.int64), np.zeros(NI, np.int64)
    R.ref_bpr_ids(h, 0, uid)
    R.ref_bpr_ids(h, 1, iid)
    du, di = np.zeros(ND, np.int64), np.zeros(ND, np.int64)
    R.ref_bpr_data(h, du, di)
    ev = {}
    for test in (0, 1):
        n = R.ref_bpr_eval_size(h, test)
        a, b, c = np.zeros(n, np.int64), np.zeros(n, np.int64), np.zeros(n, np.int64)
        R.ref_bpr_eval_set(h, test, a, b, c)
        ev["eu%d" % test], ev["ei%d" % test], ev["ej%d" % test] = a, b, c
    P0, Q0, b0 = np.zeros((NU, k)), np.zeros((NI, k)), np.zeros(NI)
    R.ref_bpr_get_factors(h, 0, P0)
    R.ref_bpr_get_factors(h, 1, Q0)
    if biases:
        R.ref_bpr_get_biases(h, b0)
    ntu = R.ref_bpr_num_test_users(h)
    test_users = np.zeros(ntu, np.int64)
    R.ref_bpr_test_users(h, test_users)
    losses = []
    for _ in range(nepochs):
        R.ref_bpr_optimize(h)
        losses.append((R.ref_bpr_eval_loss(h, 0), R.ref_bpr_eval_loss(h, 1)))
    P, Q, b = np.zeros((NU, k)), np.zeros((NI, k)), np.zeros(NI)
    R.ref_bpr_get_factors(h, 0, P)
    R.ref_bpr_get_factors(h, 1, Q)
    if biases:
        R.ref_bpr_get_biases(h, b)
    # reference evaluation of the final factors: dense scores + per-user-average metrics
    scores = np.zeros((ntu, NI))
    R.ref_compute_test_scores(P, NU, Q, NI, k, oracle.ptr(b) if biases else None, test_users, ntu, 4, scores)
    labels = np.zeros((ntu, NI))
    tusers2 = np.zeros(ntu, np.int64)
    n2 = R.ref_init_avg_test_data(uid, NU, iid, NI, tu, ti, tv, len(tu), 0, 42, oracle.ptr(tusers2), oracle.ptr(labels))
    assert n2 == ntu and np.array_equal(tusers2, test_users)
    metrics = {m: R.ref_metric_avg(m.encode(), labels, scores, ntu, NI, 4) for m in ("auc", "ap", "p@5", "r@5")}
    R.ref_bpr_destroy(h)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), u=u, i=i, v=v, tu=tu, ti=ti, tv=tv, uid=uid, iid=iid, du=du,
                        di=di, P0=P0, Q0=Q0, b0=b0, P=P, Q=Q, b=b, losses=np.array(losses), test_users=test_users,
                        labels=labels, metric_names=np.array(list(metrics)), metric_values=np.array(list(metrics.values())),
                        cfg_keys=np.array(list(cfg)), cfg_vals=np.array([float(x) for x in cfg.values()]), **ev)
    print(name, "nusers", NU, "nitems", NI, "ndata", ND, "losses", losses[-1], metrics)


def misc_case(name):
    rng = np.random.default_rng(3)
    out = {}
    # linearSymmetricSolve on the MatrixTest-style random symmetric INDEFINITE system (MatrixTest.cpp:92-116)
    n = 50
    M = rng.uniform(-1, 1, (n, n))
    A = np.ascontiguousarray(M + M.T)
    b = rng.uniform(-1, 1, n)
    x = np.zeros(n)
    R.ref_linear_symmetric_solve(A, b.copy(), n, x)
    out.update(solve_A=A, solve_b=b, solve_x=x)
    # Gram, both reference variants
    Y = rng.uniform(-0.5, 0.5, (17, 5))
    G0, G1 = np.zeros((5, 5)), np.zeros((5, 5))
    R.ref_gram(Y, 17, 5, 4, 0, G0)
    R.ref_gram(Y, 17, 5, 4, 1, G1)
    out.update(gram_Y=Y, gram_racefree=G0, gram_used=G1)
    # saveFactors text
    F = np.round(rng.uniform(-2, 2, (4, 3)), 12)
    bias = np.round(rng.uniform(-2, 2, 4), 12)
    ids = np.array([10, -3, 77, 123456789012], np.int64)
    for tag, bb in (("nobias", None), ("bias", bias)):
        need = R.ref_save_factors(F, oracle.ptr(bb) if bb is not None else None, ids, 4, 3, None, 0)
        buf = bytes(need)
        import ctypes
        cbuf = ctypes.create_string_buffer(need)
        R.ref_save_factors(F, oracle.ptr(bb) if bb is not None else None, ids, 4, 3, cbuf, need)
        out["savetxt_" + tag] = np.frombuffer(cbuf.raw, dtype=np.uint8).copy()
    out.update(save_F=F, save_bias=bias, save_ids=ids)
    # BPREngine::update on explicit triplets (BPREngine.cpp:178-220), k = 9 with biases
    k, nu, ni = 9, 12, 10
    cells = rng.choice(nu * ni, 60, replace=False)
    u, i = (cells // ni).astype(np.int64), (cells % ni).astype(np.int64)
    h = R.ref_bpr_create(k, 1, 0.05, 1.0, 0.025, 0.0025, 0.9, 1, 0.01, 3, 1, 1, 3, 42, 2, None, 0, 0, 5)
    R.ref_bpr_init(h, u, i, np.ones(60), 60)
    NU, NI = R.ref_bpr_nusers(h), R.ref_bpr_nitems(h)
    P0, Q0, b0 = rng.uniform(-0.5, 0.5, (NU, k)), rng.uniform(-0.5, 0.5, (NI, k)), rng.uniform(-0.5, 0.5, NI)
    R.ref_bpr_set_factors(h, 0, P0)
    R.ref_bpr_set_factors(h, 1, Q0)
    R.ref_bpr_set_biases(h, b0)
    tu, ti = rng.integers(0, NU, 80), rng.integers(0, NI, 80)
    tj = (ti + rng.integers(1, NI, 80)) % NI
    xs = []
    for t in range(80):
        xs.append(R.ref_bpr_predict_difference(h, int(tu[t]), int(ti[t]), int(tj[t])))
        R.ref_bpr_update(h, int(tu[t]), int(ti[t]), int(tj[t]))
    P, Q, b = np.zeros((NU, k)), np.zeros((NI, k)), np.zeros(NI)
    R.ref_bpr_get_factors(h, 0, P)
    R.ref_bpr_get_factors(h, 1, Q)
    R.ref_bpr_get_biases(h, b)
    R.ref_bpr_destroy(h)
    out.update(step_P0=P0, step_Q0=Q0, step_b0=b0, step_u=tu, step_i=ti, step_j=tj, step_x=np.array(xs), step_P=P, step_Q=Q,
               step_b=b, step_cfg=np.array([0.05, 0.025, 0.0025, 1.0]))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "ok")


if __name__ == "__main__":
    wals_case("wals_k30", 300, 200, 6000, 30, 3, seed=11, dup=40)
    wals_case("wals_k128", 400, 300, 12000, 128, 2, seed=12)
    wals_case("wals_k64", 220, 160, 5000, 64, 2, seed=13)
    bpr_case("bpr_k30", 400, 250, 9000, 30, 4, seed=21, biases=True)
    bpr_case("bpr_k16_nobias", 200, 120, 3000, 16, 3, seed=22, biases=False)
    misc_case("misc")
