#!/usr/bin/env python
"""Per-phase cycle breakdown of wals_solve_kernel (debug build: make -C qmf_b200/csrc prof).
usage: QMFB_LIB=qmf_b200/libqmf_b200_prof.so python tools/exp_phases.py nrows nnz_per_row ncols"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qmf_b200.wals_dist import CudaKernels
dev = torch.device("cuda", 0)
K = CudaKernels()
k = 128
kp = K.padded_k(k)
names = ["build", "tile store+b", "factor diag (warp0)", "panel (b)", "wait top-of-step", "wait after panel", "diag tile update",
         "cholesky total", "back substitution", "loss/store/next", "row total", "  build: issue_one", "  build: wait full",
         "  build: chunk_mma", "  build: b-acc + arrive", "(rows)", "  build: prologue issue", "  build: gram load", "  build: chunk loop",
         "  build: b partial store", "  build: end barrier / wait tile buffer", "  build: tile store", "ws solver: wait for tiles",
         "ws solver: solve+loss+store"]
K.lib.qmfb_debug_set_flags(int(os.environ.get("EXP_FLAGS", "0")))
for spec in sys.argv[1:]:
    nrows, nnz_row, ncols = (int(x) for x in spec.split(","))
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
    buf = (C.c_ulonglong * 24)()
    K.solve(X, 0, Y, k, row_ptr, col, val, order, gram, 40.0, 0.05, row_loss, loss, scratch)
    torch.cuda.synchronize()
    K.lib.qmfb_debug_phase_cycles(buf)
    K.solve(X, 0, Y, k, row_ptr, col, val, order, gram, 40.0, 0.05, row_loss, loss, scratch)
    torch.cuda.synchronize()
    K.lib.qmfb_debug_phase_cycles(buf)
    rows = buf[15]
    print("== rows=%d nnz/row=%d ycols=%d  (cycles per row, thread 0 of each CTA; %d rows)" % (nrows, nnz_row, ncols, rows))
    for i, n in enumerate(names):
        if i == 15:
            continue
        print("  %-24s %9.0f" % (n, buf[i] / max(rows, 1)))
