"""C4 epoch time with the plain and the warp-specialised solve kernel (user / item half-step ms)."""
import ctypes as C
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qmf_b200 import capi
from qmf_b200.datagen import CONFIGS, init_item_factors, uniform_csr_torch
from qmf_b200.wals_dist import ShardedWals

name = sys.argv[1] if len(sys.argv) > 1 else "c4"
nu, ni, nnz, k = CONFIGS[name]
dev = torch.device("cuda", 0)
csr_user, csr_item = uniform_csr_torch(nu, ni, nnz, seed=20240501, device=dev)
sw = ShardedWals(nu, ni, k, csr_user, csr_item, dev)
Y0 = init_item_factors(ni, k, seed=7)
res = {}
for mode, label in ((1, "plain"), (2, "ws"), (0, "auto")):
    capi.check(capi.lib.qmfb_wals_set_solve_kernel(mode))
    sw.set_factors(1, Y0)
    for _ in range(2):
        sw.epoch(40.0, 0.05)
    ev = [[{n: torch.cuda.Event(enable_timing=True) for n in ("gram0", "solve0", "solve1")} for _ in range(2)] for _ in range(3)]
    torch.cuda.synchronize()
    for s in range(3):
        loss = sw.epoch(40.0, 0.05, ev[s])
    torch.cuda.synchronize()
    sw.check_error()
    um = min(e[0]["solve0"].elapsed_time(e[0]["solve1"]) for e in ev)
    im = min(e[1]["solve0"].elapsed_time(e[1]["solve1"]) for e in ev)
    res[label] = {"user_ms": um, "item_ms": im, "loss": float(loss.item())}
    print(label, res[label], flush=True)
print(json.dumps(res))
