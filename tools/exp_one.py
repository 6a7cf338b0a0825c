#!/usr/bin/env python
"""Single-shape run of wals_solve_kernel for profiling: exp_one.py nrows nnz_per_row ncols"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qmf_b200.wals_dist import CudaKernels
dev = torch.device("cuda", 0)
K = CudaKernels()
k = 128
kp = K.padded_k(k)
nrows, nnz_row, ncols = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
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
for _ in range(2):
    K.solve(X, 0, Y, k, row_ptr, col, val, order, gram, 40.0, 0.05, row_loss, loss, scratch)
torch.cuda.synchronize()
print("ok", float(loss))
