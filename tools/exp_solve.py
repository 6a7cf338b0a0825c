#!/usr/bin/env python
"""Phase experiments for wals_solve_kernel on synthetic row shapes (GPU only).
usage: exp_solve.py  -> prints ms, cycles per row per CTA and DMMA efficiency for several shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qmf_b200.wals_dist import CudaKernels

dev = torch.device("cuda", 0)
K = CudaKernels()
k = int(os.environ.get("EXP_K", "128"))
kp = K.padded_k(k)


def run(name, nrows, nnz_row, ncols, reps=3):
    g = torch.Generator(device=dev).manual_seed(1)
    Y = (torch.rand(ncols, kp, generator=g, device=dev, dtype=torch.float64) - 0.5) * 0.1
    X = torch.zeros(nrows, kp, device=dev, dtype=torch.float64)
    row_ptr = (torch.arange(nrows + 1, device=dev, dtype=torch.int64) * nnz_row)
    col = torch.randint(0, ncols, (max(nrows * nnz_row, 1),), generator=g, device=dev, dtype=torch.int32)
    val = torch.randint(1, 6, (max(nrows * nnz_row, 1),), generator=g, device=dev).to(torch.float64)
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
    nnz = nrows * nnz_row
    flops = nnz * (k * (k + 1) + 2 * k) + nrows * (k ** 3 / 3 + 4 * k * k)
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    cyc_row_sm = best * 1e-3 * 1.965e9 / (nrows / sms)
    print("%-34s rows=%7d nnz/row=%6d ycols=%7d  %8.3f ms  %7.2f TF  %9.0f cyc/row/SM  err=%d" % (
        name, nrows, nnz_row, ncols, best, flops / best * 1e-9, cyc_row_sm, int(scratch[1])), flush=True)


run("solve-only (1 chunk)", 148 * 400, 16, 17800)
run("solve + 2 chunks", 148 * 400, 32, 17800)
run("user-like, L2-resident Y", 148 * 200, 208, 17800)
run("user-like x2 nnz", 148 * 100, 416, 17800)
run("item-like, DRAM Y (480k rows)", 148 * 8, 5618, 480000)
run("item-like, small Y (2k rows)", 148 * 8, 5618, 2000)
run("item-like, L2 Y (60k rows)", 148 * 8, 5618, 60000)
