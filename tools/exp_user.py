#!/usr/bin/env python
"""Time wals_solve_kernel on the user-shaped and item-shaped halves of C4 for the library in $QMFB_LIB."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qmf_b200.wals_dist import CudaKernels
dev = torch.device("cuda", 0)
K = CudaKernels()
k = 128
kp = K.padded_k(k)
def run(nrows, nnz_row, ncols, reps=3):
    g = torch.Generator(device=dev).manual_seed(1)
    Y = (torch.rand(ncols, kp, generator=g, device=dev, dtype=torch.float64) - 0.5) * 0.1
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
        K.solve(X, 0, Y, k, row_ptr, col, val, order, gram, 40.0, 0.05, row_loss, loss, scratch)
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best
u = run(148 * 800, 208, 17770)
i = run(148 * 30, 5618, 480189)
print("%-28s user-like %.2f ms   item-like %.2f ms" % (os.environ.get("QMFB_LIB", "default"), u, i), flush=True)
