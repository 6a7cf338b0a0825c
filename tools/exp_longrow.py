#!/usr/bin/env python
"""Item half-step with one blockbuster item (400k ratings) among 2M other entries, k = 128:
solve time with the long-row pre-pass and (QMFB_NO_LONG_ROWS=1) without it."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from qmf_b200 import WalsEngineHandle, csr_from_coo
rng = np.random.default_rng(1)
nu, ni, k = 500_000, 2_000, 128
u = np.concatenate([rng.choice(nu, 400_000, replace=False), rng.integers(0, nu, 2_000_000)]).astype(np.int64)
i = np.concatenate([np.full(400_000, 3), rng.integers(0, ni, 2_000_000)]).astype(np.int64)
cells = np.unique(u * ni + i)
u, i = cells // ni, cells % ni
v = rng.integers(1, 6, len(cells)).astype(np.float64)
uids, urp, uci, uv = csr_from_coo(u, i, v)
iids, irp, ici, iv = csr_from_coo(i, u, v)
h = WalsEngineHandle(len(uids), len(iids), k)
h.set_csr(0, urp, uci, uv); h.set_csr(1, irp, ici, iv)
h.set_factors(1, rng.uniform(-0.01, 0.01, (len(iids), k)))
best = 1e9
for _ in range(3):
    h.half_step(0, 40.0, 0.05)
    loss = h.half_step(1, 40.0, 0.05)
    best = min(best, h.last_timing()[1])
print("long rows %s: item solve %.3f ms, longest row %d nnz, loss %.15g" % (
    "IGNORED" if os.environ.get("QMFB_NO_LONG_ROWS") else "pre-built", best, int(np.diff(irp).max()), loss))
