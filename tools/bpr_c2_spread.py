#!/usr/bin/env python
"""Spread of the REFERENCE's own BPR on the planted C2 problem (run here, where oracle/_ref exists):
>= 5 seeded BPREngine::optimize runs with num_hogwild_threads = 1 and = 16 -> train loss, test loss, AUC, p@10
after 10 epochs.  Writes tests/golden/bpr_c2_spread.json, the band tests/test_bpr_c2_gpu.py holds the GPU to."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle
from bpr_c2 import HP, K, NI, NU, planted_c2, rank_metrics

L = oracle.ref()
L.ref_set_min_log_level(2)
tr_u, tr_i, te_u, te_i = planted_c2()
runs = {}
for hog in (1, 16):
    rows = []
    for seed in range(5):
        h = L.ref_bpr_create(K, HP["nepochs"], HP["lr"], HP["bias_lambda"], HP["user_lambda"], HP["item_lambda"], HP["decay"], 1,
                             HP["init_bound"], HP["num_neg"], hog, 1, HP["eval_num_neg"], HP["eval_seed"], max(hog, 1), None, 0, 0,
                             1000 + seed)
        L.ref_bpr_init(h, tr_u + 1, tr_i + 1, np.ones(len(tr_u)), len(tr_u))
        L.ref_bpr_init_test(h, te_u + 1, te_i + 1, np.ones(len(te_u)), len(te_u))
        sec = L.ref_bpr_optimize(h)
        nu, ni = L.ref_bpr_nusers(h), L.ref_bpr_nitems(h)
        uid, iid = np.zeros(nu, np.int64), np.zeros(ni, np.int64)
        L.ref_bpr_ids(h, 0, uid); L.ref_bpr_ids(h, 1, iid)
        P, Q, b = np.zeros((nu, K)), np.zeros((ni, K)), np.zeros(ni)
        L.ref_bpr_get_factors(h, 0, P); L.ref_bpr_get_factors(h, 1, Q); L.ref_bpr_get_biases(h, b)
        # dense idx -> raw id - 1 (users / items that never occur keep zero factors; none on this problem)
        Pf, Qf, bf = np.zeros((NU, K)), np.zeros((NI, K)), np.zeros(NI)
        Pf[uid - 1], Qf[iid - 1], bf[iid - 1] = P, Q, b
        auc, p10 = rank_metrics(Pf, Qf, bf, te_u, te_i)
        row = dict(seed=1000 + seed, train_loss=L.ref_bpr_eval_loss(h, 0), test_loss=L.ref_bpr_eval_loss(h, 1), auc=auc, p10=p10,
                   seconds=sec)
        L.ref_bpr_destroy(h)
        print(hog, row, flush=True)
        rows.append(row)
    runs["hogwild_%d" % hog] = rows
summary = {}
for name, rows in runs.items():
    summary[name] = {m: dict(mean=float(np.mean([r[m] for r in rows])), min=float(min(r[m] for r in rows)),
                             max=float(max(r[m] for r in rows))) for m in ("train_loss", "test_loss", "auc", "p10")}
out = dict(problem="planted_c2(seed=2024): 10k x 5k, 450k train + 50k test pairs, k=30, biases, 3 negatives, 10 epochs, "
                   "lr 0.05 decay 0.9 (qmf/bpr.cpp defaults)", runs=runs, summary=summary)
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "bpr_c2_spread.json"), "w"), indent=1)
print(json.dumps(summary, indent=1))
