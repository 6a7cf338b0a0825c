#!/usr/bin/env python
"""End-to-end CLI fixtures: runs the reference's own `wals` binary (oracle/_ref/wals, unmodified
sources) on a small seeded dataset and stores inputs + outputs under tests/golden/cli/.
Run where /root/reference is mounted:  python tests/golden/make_cli_golden.py"""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import init_factors, uniform_dataset  # noqa: E402

out = os.path.join(HERE, "cli")
os.makedirs(out, exist_ok=True)
u, i, v = uniform_dataset(260, 180, 5200, seed=77, id_scale=(3, 7), dup=15)
ntest = 500
with open(os.path.join(out, "train.txt"), "w") as f:
    for a, b, c in zip(u[:-ntest], i[:-ntest], v[:-ntest]):
        f.write("%d %d %g\n" % (a, b, c))
with open(os.path.join(out, "test.txt"), "w") as f:
    for a, b, c in zip(u[-ntest:], i[-ntest:], v[-ntest:]):
        f.write("%d %d %g\n" % (a, b, c))
nitems = len(np.unique(i[:-ntest]))
k = 30
with open(os.path.join(out, "dist.txt"), "w") as f:
    for x in init_factors(nitems, k, seed=3).reshape(-1):
        f.write("%.9f\n" % x)
env = dict(os.environ, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1")
cmd = [os.path.join(ROOT, "oracle", "_ref", "wals"), "--nepochs=3", "--nfactors=%d" % k, "--regularization_lambda=0.05",
       "--confidence_weight=40", "--nthreads=4", "--train_dataset=" + os.path.join(out, "train.txt"),
       "--test_dataset=" + os.path.join(out, "test.txt"), "--distribution_file=" + os.path.join(out, "dist.txt"),
       "--test_avg_metrics=auc,ap,p@10,r@10", "--test_always",
       "--user_factors=" + os.path.join(out, "ref_user_factors.txt"), "--item_factors=" + os.path.join(out, "ref_item_factors.txt")]
r = subprocess.run(cmd, env=env, capture_output=True, text=True, check=True)
lines = [l.split("] ", 1)[1] for l in r.stderr.splitlines() if "epoch" in l and ("train loss" in l or "recorded metric" in l)]
open(os.path.join(out, "ref_log.txt"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines))

# --num_test_users: the reference subsamples the test users with std::shuffle(mt19937(eval_seed)) over an unordered_set's
# iteration order (qmf/Engine.cpp:35-50); which users are drawn decides the averaged metrics
for ntu, seed in ((40, 7), (120, 42)):
    cmd2 = cmd[:-2] + ["--num_test_users=%d" % ntu, "--eval_seed=%d" % seed]
    r = subprocess.run(cmd2, env=env, capture_output=True, text=True, check=True)
    lines = [l.split("] ", 1)[1] for l in r.stderr.splitlines() if "recorded metric" in l]
    open(os.path.join(out, "ref_log_numtest%d_seed%d.txt" % (ntu, seed)), "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[-4:]))
