"""Ranking evaluation through the C ABI (qmfb_eval_rank) + the host-side step that turns the
GPU's integer rank statistics into AUC / AP / P@k / R@k with the reference's own arithmetic
(qmf/metrics/Metrics.cpp:65-164).  The C++ host (qmf_b200/host/qmf/metrics) does the same."""
import ctypes as C

import numpy as np

from .capi import check, lib


def labels_to_csr(label_rows):
    """label_rows: list (per test user) of ascending distinct positive item idx -> (ptr, items)"""
    ptr = np.zeros(len(label_rows) + 1, dtype=np.int64)
    np.cumsum([len(r) for r in label_rows], out=ptr[1:])
    items = np.concatenate([np.asarray(r, dtype=np.int32) for r in label_rows]) if ptr[-1] else np.zeros(0, np.int32)
    return ptr, np.ascontiguousarray(items, dtype=np.int32)


def eval_rank(U, V, biases, test_users, label_ptr, label_items, device=0):
    """returns (cnt, pos_scores): see include/qmf_b200.h qmfb_eval_rank"""
    U = np.ascontiguousarray(U, dtype=np.float64)
    V = np.ascontiguousarray(V, dtype=np.float64)
    tu = np.ascontiguousarray(test_users, dtype=np.int32)
    lp = np.ascontiguousarray(label_ptr, dtype=np.int64)
    li = np.ascontiguousarray(label_items, dtype=np.int32)
    nT, nl = len(tu), int(lp[-1])
    cnt = np.zeros(nl + nT, dtype=np.int32)
    sc = np.zeros(max(nl, 1), dtype=np.float64)
    if li.size == 0:
        li = np.zeros(1, np.int32)
    b = None if biases is None else np.ascontiguousarray(biases, dtype=np.float64).ctypes.data_as(C.c_void_p)
    check(lib.qmfb_eval_rank(device, U, U.shape[0], V, V.shape[0], U.shape[1], b, tu, nT, lp, li, cnt, sc))
    return cnt, sc[:nl]


def user_metrics(cnt_t, nitems, names):
    """Metrics of one test user from its bucket counts cnt_t[0..nP] (exact reference arithmetic:
    AUC accumulates (double)tp / pos / neg once per negative in rank order, Metrics.cpp:87-95;
    AP accumulates (double)pos / (i + 1) per positive in rank order, :156-162)."""
    nP = len(cnt_t) - 1
    nN = nitems - nP
    out = {}
    # position (0-based) of the q-th positive in ascending-score order
    greater = np.concatenate([np.cumsum(cnt_t[::-1])[::-1][1:], [0]])  # greater[q] = sum_{i>q} cnt[i]
    pos_index = greater[:nP] + (nP - 1 - np.arange(nP))
    for name in names:
        if name == "auc":
            if nP == 0 or nN == 0:
                out[name] = 1.0
                continue
            auc = 0.0
            for i in range(nP, -1, -1):
                term = float(nP - i) / nP / nN
                # cnt additions of the same term, exact result of the one-by-one loop
                auc = float(lib.qmfb_repeated_add(auc, term, int(cnt_t[i])))
            out[name] = auc
        elif name == "ap":
            ap = 0.0
            for q in range(nP - 1, -1, -1):
                ap += float(nP - q) / float(pos_index[q] + 1)
            out[name] = ap / nP
        elif name.startswith("p@") or name.startswith("r@"):
            k = int(name[2:])
            hits = int(np.sum(pos_index < k))
            out[name] = hits / float(k) if name[0] == "p" else hits / float(nP)
        else:
            raise ValueError(name)
    return out


def average_metric(values, nthreads):
    """Metric::compute(labels, scores, parallel) (Metrics.cpp:38-52): strided per-thread partial sums
    folded in thread order (ParallelExecutor::mapReduce), divided by the number of users;
    nthreads == 0 is the serial overload (:27-36)."""
    n = len(values)
    if nthreads <= 0:
        s = 0.0
        for v in values:
            s += v
        return s / n
    total = 0.0
    for th in range(nthreads):
        part = 0.0
        for t in range(th, n, nthreads):
            part = part + values[t]
        total = total + part
    return total / n


def rank_metrics(name, cnt, label_ptr, nitems, host_threads=0):
    """per-user values of one ranking metric for ALL test users at once (qmfb_rank_metrics, host threads)"""
    lp = np.ascontiguousarray(label_ptr, dtype=np.int64)
    c = np.ascontiguousarray(cnt, dtype=np.int32)
    out = np.empty(len(lp) - 1, dtype=np.float64)
    check(lib.qmfb_rank_metrics(name.encode(), c, lp, len(lp) - 1, int(nitems), host_threads, out))
    return out
