#!/usr/bin/env python
"""Correctness (numpy normal equations on sampled rows) + timing of the solve kernel for the library in
$QMFB_LIB and the kernel choice in $QMFB_SOLVE (classic | ws) on the user-shaped and item-shaped halves of C4.
Run every variant under `timeout`: a kernel that deadlocks must cost seconds, not the call."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from qmf_b200.wals_dist import CudaKernels
dev = torch.device("cuda", 0)
K = CudaKernels()
k = int(os.environ.get("EXP_K", "128"))
kp = K.padded_k(k)
ALPHA, LAM = 40.0, 0.05


def run(nrows, nnz_row, ncols, reps=3, check=16):
    g = torch.Generator(device=dev).manual_seed(1)
    Y = torch.zeros(ncols, kp, device=dev, dtype=torch.float64)
    Y[:, :k] = (torch.rand(ncols, k, generator=g, device=dev, dtype=torch.float64) - 0.5) * 0.1
    X = torch.zeros(nrows, kp, device=dev, dtype=torch.float64)
    row_ptr = (torch.arange(nrows + 1, device=dev, dtype=torch.int64) * nnz_row)
    col = torch.randint(0, ncols, (nrows * nnz_row,), generator=g, device=dev, dtype=torch.int32)
    val = torch.randint(1, 6, (nrows * nnz_row,), generator=g, device=dev).to(torch.float64)
    order = torch.arange(nrows, device=dev, dtype=torch.int32)
    gram = torch.zeros(K.gram_packed_len(k), device=dev, dtype=torch.float64)
    ws = torch.empty(K.gram_workspace_len(k), device=dev, dtype=torch.float64)
    K.gram(Y, 0, ncols, k, ws, gram)
    row_loss = torch.zeros(nrows, device=dev, dtype=torch.float64)
    loss = torch.zeros(1, device=dev, dtype=torch.float64)
    scratch = torch.zeros(2, device=dev, dtype=torch.int32)
    best = 1e30
    for _ in range(reps + 1):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        K.solve(X, 0, Y, k, row_ptr, col, val, order, gram, ALPHA, LAM, row_loss, loss, scratch)
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    Yh = Y[:, :k].cpu().numpy()
    G = Yh.T @ Yh
    err = 0.0
    for r in np.linspace(0, nrows - 1, check).astype(int):
        sl = slice(r * nnz_row, (r + 1) * nnz_row)
        Ys, w = Yh[col[sl].cpu().numpy()], val[sl].cpu().numpy()
        A = G + (Ys * (ALPHA * w)[:, None]).T @ Ys + LAM * np.eye(k)
        b = ((1 + ALPHA * w)[:, None] * Ys).sum(0)
        x = np.linalg.solve(A, b)
        err = max(err, float(np.abs(X[r, :k].cpu().numpy() - x).max() / np.abs(x).max()))
    return best, err, int(scratch[1].item())


u, eu, fu = run(148 * 800, 208, 17770)
i, ei, fi = run(148 * 30, 5618, 480189)
print("%-34s %-8s user-like %.2f ms (err %.1e, flag %d)   item-like %.2f ms (err %.1e, flag %d)" % (
    os.path.basename(os.environ.get("QMFB_LIB", "default")), os.environ.get("QMFB_SOLVE", "auto"), u, eu, fu, i, ei, fi), flush=True)
