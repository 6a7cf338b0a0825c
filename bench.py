#!/usr/bin/env python
"""bench.py — WALS epoch throughput (BASELINE.json metric: "WALS s/epoch & nnz/s at 1/2/4/8 B200")
on the Netflix-shaped synthetic config C4 (480k users x 17.8k items, 100M nnz, k=128, FP64).

  python bench.py --gpus 1 --steps K --warmup W                 # our arm, one GPU
  torchrun ... bench.py --gpus N --steps K --warmup W           # our arm, N ranks (row-partitioned)
  python bench.py --impl reference --steps K --warmup W         # the reference's CPU path on host cores

A step = one WALS epoch (user half-step + item half-step: Gram, per-row build, solve, loss).
`value` = nnz per epoch / device seconds per epoch (each nnz is touched twice per epoch), inputs
resident in HBM.  `e2e` = the same through host buffers (H2D of the item factors, D2H of both
factor matrices and the loss every step).  The dataset (CSR+CSC 2.4 GB, factors 0.5 GB) is far
larger than the 126 MB L2, so no explicit L2 flush is needed between steps.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALPHA, LAMBDA = 40.0, 0.05          # reference defaults (qmf/wals.cpp:28-29)
L2_NOTE = "inputs (2.9 GB) larger than the 126 MB L2; no flush"


def fp64_peak_tflops():
    """FP64 DMMA peak, builder-measured on this pool's B200 (tools/fp64_peak.cu -> profiles/fp64_peak.json);
    MEASURED_PEAKS.json carries no FP64 entry."""
    with open(os.path.join(ROOT, "profiles", "fp64_peak.json")) as f:
        return float(json.load(f)["dmma_tflops"])



def algorithmic_flops(n_rows, nnz, k):
    """SURVEY.md §8(d), symmetric-aware, one half-step over n_rows rows with nnz signals."""
    build = nnz * (k * (k + 1) + 2 * k)
    solve = n_rows * (k ** 3 / 3.0 + 2 * k * k)
    loss = n_rows * (2 * k * k + 3 * k)
    return build + solve + loss


def algorithmic_bytes(n_rows, n_other, nnz, k):
    """CSR idx32+val64, one gathered row of k doubles per nnz, factor write, Gram read of the other side"""
    return nnz * 12 + nnz * 8 * k + n_rows * 8 * k + n_other * 8 * k


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 8:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            # median over samples taken under load (top half)
            load = sorted(sm)[len(sm) // 2:]
            out.update(sm_mhz=float(np.median(load)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------------
# reference / CPU baseline leg (the only place bench.py executes anything under oracle/)
# ------------------------------------------------------------------------------------------------
def cpu_sample_problem(cfg, su, si, sg, seed=123):
    from qmf_b200.datagen import uniform_row_sample
    nu, ni, nnz, k = cfg
    p = nnz / (nu * ni)
    rng = np.random.default_rng(seed)
    user_rows = uniform_row_sample(nu, ni, p, su, seed + 1)
    item_rows = uniform_row_sample(ni, nu, p, si, seed + 2)
    Yi = rng.uniform(-0.01, 0.01, size=(ni, k))     # item factors (fixed side of the user step)
    Yu = rng.uniform(-0.3, 0.3, size=(nu, k))       # user factors (fixed side of the item step)
    return dict(user_rows=user_rows, item_rows=item_rows, Yi=Yi, Yu=Yu, sg=sg)


def cpu_sample_step(cfg, prob, threads):
    """One bounded sample of the reference's epoch: `su` user rows + `si` item rows through the
    reference's own updateFactorsForOne on its ParallelExecutor (oracle/_ref), plus its serial
    Gram on `sg` rows; extrapolated to the full epoch by row counts.  Returns (epoch_seconds
    estimate, sample_seconds, kind)."""
    import oracle
    nu, ni, nnz, k = cfg
    if oracle.ref_available():
        L = oracle.ref()
        L.ref_set_min_log_level(2)
        sec = C.c_double()
        t_meas = 0.0
        est = 0.0
        for rows, Y, nright, nleft_total in ((prob["user_rows"], prob["Yi"], ni, nu), (prob["item_rows"], prob["Yu"], nu, ni)):
            rp, col, val = rows
            ns = len(rp) - 1
            # Gram of the fixed side: the reference's (serial under OMP_NUM_THREADS=1) computeXtX
            sg = min(prob["sg"], nright)
            G = np.zeros((k, k))
            t0 = time.perf_counter()
            L.ref_gram(np.ascontiguousarray(Y[:sg]), sg, k, threads, 1, G)
            tg = time.perf_counter() - t0
            X = np.zeros((ns, k))
            L.ref_wals_update_rows(X, ns, Y, nright, k, rp, col, val, G, ALPHA, LAMBDA, threads, C.byref(sec))
            t_meas += tg + sec.value
            est += tg * (nright / sg) + sec.value * (nleft_total / ns)
        return est, t_meas, "reference"
    # oracle port (single thread): smaller sample
    L = oracle.oracle()
    t_meas = est = 0.0
    for rows, Y, nright, nleft_total in ((prob["user_rows"], prob["Yi"], ni, nu), (prob["item_rows"], prob["Yu"], nu, ni)):
        rp, col, val = rows
        ns = max(1, (len(rp) - 1) // 16)
        G = np.zeros((k, k))
        sg = min(prob["sg"] // 4, nright)
        t0 = time.perf_counter()
        L.qmfo_gram(np.ascontiguousarray(Y[:sg]), sg, k, G)
        tg = time.perf_counter() - t0
        x = np.zeros(k)
        t0 = time.perf_counter()
        for r in range(ns):
            L.qmfo_wals_update_row(Y, k, col[rp[r]:rp[r + 1]], val[rp[r]:rp[r + 1]], rp[r + 1] - rp[r], G, ALPHA, LAMBDA, x)
        tr = time.perf_counter() - t0
        t_meas += tg + tr
        est += tg * (nright / sg) + tr * (nleft_total / ns)
    return est, t_meas, "port"


def sample_sizes(cfg):
    nu, ni, nnz, k = cfg
    # ~2.5 M nnz per side at k=128 (~1.7 s per side on 16 threads), scaled by k^2 for smaller k
    budget = 2.5e6 * (128.0 / k) ** 2
    su = int(min(nu, max(64, budget / (nnz / nu))))
    si = int(min(ni, max(16, budget / (nnz / ni))))
    return su, si, 8000


def run_reference(args, cfg, workload):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    os.environ.setdefault("OMP_NUM_THREADS", "1")       # the reference's OpenMP Gram is racy
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    nu, ni, nnz, k = cfg
    threads = os.cpu_count() or 1
    su, si, sg = sample_sizes(cfg)
    prob = cpu_sample_problem(cfg, su, si, sg)
    for _ in range(args.warmup):
        cpu_sample_step(cfg, prob, threads)
    ests, meas, kind = [], [], "reference"
    for _ in range(args.steps):
        e, m, kind = cpu_sample_step(cfg, prob, threads)
        ests.append(e)
        meas.append(m)
    epoch_s = float(np.mean(ests))
    value = nnz / epoch_s
    sample = ("%d of %d user rows + %d of %d item rows of the %s uniform problem through the reference's "
              "updateFactorsForOne on %d pool threads (+ its serial Gram on %d rows), extrapolated to one epoch by row "
              "count; OMP_NUM_THREADS=1" % (su, nu, si, ni, workload, threads, sg))
    line = {
        "impl": "reference", "metric": "wals_nnz_per_s", "value": value, "unit": "nnz/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(meas)) * 1e3,
        "est_epoch_s": epoch_s, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        # the same config object as our arm prints for this N (the last three keys describe OUR arm's run of the workload)
        "data": "synthetic", "config": {"workload": workload, "nusers": nu, "nitems": ni, "nnz": nnz, "nfactors": k,
                                        "alpha": ALPHA, "lambda": LAMBDA, "parallelism": "rows x%d" % args.gpus,
                                        "exchange": "p2p" if args.gpus > 1 else "none", "l2": L2_NOTE},
        "cpu_baseline": {"value": value, "unit": "nnz/s", "cores": threads if kind == "reference" else 1, "kind": kind,
                         "sample": sample},
        "e2e": {"value": value, "unit": "nnz/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# BPR (north_star part 2): triplet updates/s on 1 GPU, as achieved algorithmic HBM GB/s
# ------------------------------------------------------------------------------------------------
BPR_SHAPES = {
    # name: (nusers, nitems, npairs, nfactors)
    "c2": (10_000, 5_000, 450_000, 30),          # BASELINE.json configs[1] (L2-resident: 3.6 MB of factors)
    "large": (4_000_000, 1_000_000, 40_000_000, 64),   # 2.6 GB of factors: HBM-resident gather/scatter
}


def bpr_pairs(nu, ni, npairs, seed):
    rng = np.random.default_rng(seed)
    u = rng.integers(0, nu, size=npairs, dtype=np.int64).astype(np.int32)
    i = rng.integers(0, ni, size=npairs, dtype=np.int64).astype(np.int32)
    return u, i


def run_bpr_ours(shape, epochs=5, warmup=2, device=0):
    from qmf_b200.bpr import BprEngineHandle
    nu, ni, npairs, k = BPR_SHAPES[shape]
    u, i = bpr_pairs(nu, ni, npairs, 11)
    h = BprEngineHandle(nu, ni, k, use_biases=True, device=device)
    h.set_data(u, i)
    rng = np.random.default_rng(1)
    h.set_factors(0, rng.uniform(-0.01, 0.01, (nu, k)))
    h.set_factors(1, rng.uniform(-0.01, 0.01, (ni, k)))
    h.set_biases(rng.uniform(-0.01, 0.01, ni))
    lr, ms = 0.05, []
    for e in range(warmup + epochs):
        h.epoch(lr, 0.025, 0.0025, 1.0, 3, seed=7, epoch=e, shuffle=True)
        if e >= warmup:
            ms.append(h.last_epoch_ms())
        lr *= 0.9
    upd = npairs * 3
    per_triplet = 6 * 8 * k + 32 + 24
    sec = float(np.mean(ms)) * 1e-3
    out = {"shape": {"nusers": nu, "nitems": ni, "npairs": npairs, "nfactors": k, "num_neg": 3, "use_biases": True},
           "updates_per_s": upd / sec, "ms_per_epoch": sec * 1e3, "algorithmic_bytes_per_triplet": per_triplet,
           "achieved_gbs": upd * per_triplet / sec * 1e-9, "launches_per_epoch": 1}
    h.close()
    return out


def run_bpr_reference(threads):
    """the reference's own BPREngine::optimize (Hogwild on all host threads) on the C2 shape"""
    import oracle
    if not oracle.ref_available():
        return None
    L = oracle.ref()
    L.ref_set_min_log_level(2)
    nu, ni, npairs, k = BPR_SHAPES["c2"]
    u, i = bpr_pairs(nu, ni, npairs, 11)
    h = L.ref_bpr_create(k, 2, 0.05, 1.0, 0.025, 0.0025, 0.9, 1, 0.01, 3, threads, 1, 3, 42, threads, None, 0, 0, 7)
    L.ref_bpr_init(h, u.astype(np.int64) + 1, i.astype(np.int64) + 1, np.ones(npairs), npairs)
    sec = L.ref_bpr_optimize(h)     # 2 epochs incl. the eval-loss passes the reference always runs
    L.ref_bpr_destroy(h)
    return {"updates_per_s": 2 * npairs * 3 / sec, "cores": threads, "kind": "reference",
            "sample": "2 epochs of BPREngine::optimize on the C2 shape, num_hogwild_threads = nthreads = %d" % threads}


EVAL_SHAPES = {
    # name: (test users, nitems, nfactors, positives per user)
    "c2": (10_000, 5_000, 30, 5),              # BASELINE.json configs[1]: all users of the 10k x 5k problem
    "large": (100_000, 1_000_000, 128, 10),    # 100k test users x 1M items (C5-shaped catalogue), k=128
}


def run_eval_ours(shape, device=0, reps=4):
    """north_star part 3: all-item scoring (DMMA GEMM) fused with the rank statistics, device-resident factors"""
    import torch
    from qmf_b200 import capi
    nu, ni, k, npos = EVAL_SHAPES[shape]
    kp = capi.check(capi.lib.qmfb_padded_k(k))
    dev = torch.device("cuda", device)
    g = torch.Generator(device=dev).manual_seed(3)
    U = torch.zeros(nu, kp, device=dev, dtype=torch.float64)
    V = torch.zeros(ni, kp, device=dev, dtype=torch.float64)
    U[:, :k] = torch.rand(nu, k, generator=g, device=dev, dtype=torch.float64) - 0.5
    V[:, :k] = torch.rand(ni, k, generator=g, device=dev, dtype=torch.float64) - 0.5
    bias = torch.rand(ni, generator=g, device=dev, dtype=torch.float64) - 0.5
    tu = torch.arange(nu, device=dev, dtype=torch.int32)
    lp = torch.arange(nu + 1, device=dev, dtype=torch.int64) * npos
    li = torch.sort(torch.randint(0, ni // npos, (nu, npos), generator=g, device=dev) +
                    torch.arange(npos, device=dev) * (ni // npos), dim=1).values.to(torch.int32).reshape(-1).contiguous()
    cnt = torch.zeros(nu * npos + nu, device=dev, dtype=torch.int32)
    sc = torch.zeros(nu * npos, device=dev, dtype=torch.float64)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    ms = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        capi.check(capi.lib.qmfb_eval_rank_dev(st, U.data_ptr(), kp, V.data_ptr(), kp, ni, k, bias.data_ptr(), tu.data_ptr(), nu,
                                               lp.data_ptr(), li.data_ptr(), nu * npos, npos, cnt.data_ptr(), sc.data_ptr()))
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    t = min(ms[1:]) * 1e-3
    assert int(cnt.sum()) == nu * (ni - npos)
    tf = 2.0 * nu * ni * k / t * 1e-12
    return {"shape": {"test_users": nu, "nitems": ni, "nfactors": k, "positives_per_user": npos}, "ms": t * 1e3,
            "user_item_scores_per_s": nu * ni / t, "tflops": tf,
            "roofline": {"bound": "tensor", "achieved": tf, "peak": fp64_peak_tflops(), "unit": "TFLOP/s",
                         "frac": tf / fp64_peak_tflops(), "algorithmic_flops": 2.0 * nu * ni * k},
            "note": "DMMA score GEMM fused with bucket counting; only pairs within the proven rounding bound of a "
                    "positive's score are re-scored in the reference's exact order; 5 launches (2 memsets, item norms, "
                    "positives, score kernel); algorithmic flops 2 nT ni k"}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args, cfg, workload):
    import torch
    import torch.distributed as dist
    from qmf_b200 import capi
    from qmf_b200.datagen import init_item_factors, uniform_csr_torch
    from qmf_b200.wals_dist import ShardedWals

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    nu, ni, nnz, k = cfg

    csr_user, csr_item = uniform_csr_torch(nu, ni, nnz, seed=20240501, device=device)
    sw = ShardedWals(nu, ni, k, csr_user, csr_item, device, rank, world, exchange=args.exchange)
    Y0 = init_item_factors(ni, k, seed=7)
    sw.set_factors(1, Y0)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        sw.epoch(ALPHA, LAMBDA)
    barrier()
    sw.check_error()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [[{n: torch.cuda.Event(enable_timing=True) for n in ("gram0", "solve0", "solve1")} for _ in range(2)]
          for _ in range(args.steps)]
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = sw.launches
    barrier()
    t0.record()
    loss = None
    for s in range(args.steps):
        loss = sw.epoch(ALPHA, LAMBDA, ev[s])
    t1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    ms_per_step = ms_total / args.steps
    value = nnz / (ms_per_step * 1e-3)
    launches = sw.launches - launches0
    loss_value = float(loss.item())
    sw.check_error()

    # dominant kernel: wals_solve_kernel (2 launches per epoch) on this rank's shards
    solve_ms = [[e[h]["solve0"].elapsed_time(e[h]["solve1"]) for h in range(2)] for e in ev]
    gram_ms = [[e[h]["gram0"].elapsed_time(e[h]["solve0"]) for h in range(2)] for e in ev]
    fl = by = 0.0
    for side in (0, 1):
        sh = sw.shard[side]
        fl += algorithmic_flops(sh["end"] - sh["begin"], sh["nnz"], k)
        by += algorithmic_bytes(sh["end"] - sh["begin"], sw.n[1 - side], sh["nnz"], k)
    solve_s_per_epoch = float(np.mean([sum(x) for x in solve_ms])) * 1e-3
    achieved_tf = fl / solve_s_per_epoch * 1e-12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    traffic, traffic_src = None, None
    if world == 1:
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            key = workload.split()[0].lower()
            ent = [tj[key + "_user"]["launches"][0], tj[key + "_item"]["launches"][0]]
            traffic = [float(e["dram_bytes"]) for e in ent]
            traffic_src = "profiles/" + tj[key + "_user"]["source"] + ", profiles/" + tj[key + "_item"]["source"]
        except Exception:
            traffic = None
    traffic_total = sum(traffic) if traffic and all(t == t for t in traffic) else None
    roofline = {
        "kernel": "wals_solve_kernel<%d> (2 launches/epoch: user rows, item rows)" % (sw.kp // 8),
        "bound": "tensor", "achieved": achieved_tf, "peak": fp64_peak_tflops(), "unit": "TFLOP/s",
        "frac": achieved_tf / fp64_peak_tflops(),
        # dram__bytes_read.sum + dram__bytes_write.sum of the two launches of one epoch (user rows + item rows), from
        # the ncu --set full captures summarised in profiles/traffic.json (tools/capture_profiles.sh, tools/ncu_summary.py)
        "traffic": traffic_total, "traffic_per_launch": traffic, "traffic_source": traffic_src,
        "peak_source": "builder-measured FP64 DMMA peak on this pool's B200 (profiles/fp64_peak.json, profiles/r01_fp64_peak.txt); MEASURED_PEAKS.json "
                       "has no FP64 entry (its bf16 figure does not apply to an FP64 kernel)",
        "algorithmic_flops_per_epoch": fl, "solve_ms_user_item": [float(np.mean([x[0] for x in solve_ms])),
                                                                  float(np.mean([x[1] for x in solve_ms]))],
        "gram_ms_user_item": [float(np.mean([x[0] for x in gram_ms])), float(np.mean([x[1] for x in gram_ms]))],
        "hbm": {"achieved": by / solve_s_per_epoch * 1e-9, "peak": hbm_peak, "unit": "GB/s",
                "frac": by / solve_s_per_epoch * 1e-9 / hbm_peak, "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"},
    }

    # ---- e2e: host buffers in/out every step --------------------------------------------------
    e2e = None
    if not args.no_e2e:
        item_in = torch.from_numpy(Y0).pin_memory()
        ub, ue = sw.shard[0]["begin"], sw.shard[0]["end"]
        ib, ie = sw.shard[1]["begin"], sw.shard[1]["end"]
        user_out = torch.empty((ue - ub, k), dtype=torch.float64).pin_memory()
        item_out = torch.empty((ie - ib, k), dtype=torch.float64).pin_memory()
        loss_host = torch.empty(1, dtype=torch.float64).pin_memory()

        def e2e_step():
            sw.epoch_host(ALPHA, LAMBDA, item_in, user_out, item_out, loss_host)

        if world == 1:
            # through the engine-level C ABI (qmfb_wals_epoch_host), the call qmf::WALSEngine binds
            from qmf_b200 import WalsEngineHandle
            h = WalsEngineHandle(nu, ni, k, device=local_rank)
            for side, (rp, col, val) in enumerate((csr_user, csr_item)):
                h.set_csr(side, rp.cpu().numpy(), col.cpu().numpy(), val.cpu().numpy())
            uo = torch.empty((nu, k), dtype=torch.float64).pin_memory()
            io = torch.empty((ni, k), dtype=torch.float64).pin_memory()
            h.epoch_host(ALPHA, LAMBDA, item_in.numpy(), uo.numpy(), io.numpy())
            torch.cuda.synchronize()
            tt = time.perf_counter()
            for _ in range(args.steps):
                l2 = h.epoch_host(ALPHA, LAMBDA, item_in.numpy(), uo.numpy(), io.numpy())
            e2e_ms = (time.perf_counter() - tt) * 1e3 / args.steps
            h2d, d2h = ni * k * 8, (nu + ni) * k * 8 + 8 + 16
            launches_e2e = h.launch_count()
            h.close()
            e2e = {"value": nnz / (e2e_ms * 1e-3), "unit": "nnz/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                   "ms_per_step": e2e_ms, "api": "qmfb_wals_epoch_host (C ABI, pinned host buffers, wall clock around the "
                   "synchronous call)", "loss": l2, "launches": launches_e2e}
        else:
            e2e_step()
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(args.steps):
                e2e_step()
            b.record()
            barrier()
            t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_ms = float(t.item()) / args.steps
            e2e = {"value": nnz / (e2e_ms * 1e-3), "unit": "nnz/s", "h2d_bytes_per_step": world * ni * k * 8,
                   "d2h_bytes_per_step": (nu + ni) * k * 8 + 8 * world, "ms_per_step": e2e_ms,
                   "api": "ShardedWals.epoch_host (one process per GPU over the kernel-level C ABI: qmfb_gram_dev, "
                          "qmfb_wals_solve_peers_dev): pinned host factors in/out on every rank (CUDA events, max over ranks)"}
            # The same epoch through the engine-level C ABI that `wals --ngpus N` / qmf::WALSEngine bind: ONE process
            # (rank 0) drives all N GPUs with qmfb_wals_sharded_epoch_host while the other ranks sleep on a CPU barrier.
            cpu_group = dist.new_group(backend="gloo")
            if rank == 0:
                try:
                    from qmf_b200.wals import ShardedWalsHandle
                    hs = ShardedWalsHandle(nu, ni, k, list(range(world)))
                    for side, (rp, col, val) in enumerate((csr_user, csr_item)):
                        hs.set_csr(side, rp.cpu().numpy(), col.cpu().numpy(), val.cpu().numpy())
                    uo = torch.empty((nu, k), dtype=torch.float64).pin_memory()
                    io = torch.empty((ni, k), dtype=torch.float64).pin_memory()
                    for _ in range(2):
                        l2 = hs.epoch_host(ALPHA, LAMBDA, item_in.numpy(), uo.numpy(), io.numpy())
                    tt = time.perf_counter()
                    for _ in range(args.steps):
                        l2 = hs.epoch_host(ALPHA, LAMBDA, item_in.numpy(), uo.numpy(), io.numpy())
                    cabi_ms = (time.perf_counter() - tt) * 1e3 / args.steps
                    e2e["c_abi"] = {"value": nnz / (cabi_ms * 1e-3), "unit": "nnz/s", "ms_per_step": cabi_ms, "loss": l2,
                                    "h2d_bytes_per_step": world * ni * k * 8, "d2h_bytes_per_step": (nu + ni) * k * 8 + 8,
                                    "launches": hs.launch_count(),
                                    "api": "qmfb_wals_sharded_epoch_host: one process, %d GPUs, pinned host buffers, wall clock "
                                           "around the synchronous call (what `wals --ngpus %d` binds)" % (world, world)}
                    hs.close()
                except Exception as ex:  # noqa: BLE001 - reported in the line, the device-timed numbers stand
                    e2e["c_abi"] = {"error": str(ex)[:300]}
            dist.barrier(group=cpu_group)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        su, si, sg = sample_sizes(cfg)
        prob = cpu_sample_problem(cfg, su, si, sg)
        est, meas, kind = cpu_sample_step(cfg, prob, threads)
        cpu_baseline = {"value": nnz / est, "unit": "nnz/s", "cores": threads if kind == "reference" else 1, "kind": kind,
                        "est_epoch_s": est, "sample_s": meas,
                        "sample": "%d of %d user rows + %d of %d item rows (+ serial Gram on %d rows) through the "
                                  "reference's updateFactorsForOne, extrapolated by row count" % (su, nu, si, ni, sg)}

    bpr = None
    if rank == 0 and world == 1 and not args.no_bpr:
        hbm = float(peaks.get("hbm_gbs", 6650.0))
        bpr = {}
        for shape in ("c2", "large"):
            r = run_bpr_ours(shape, device=local_rank)
            r["roofline"] = {"bound": "hbm", "achieved": r["achieved_gbs"], "peak": hbm, "unit": "GB/s",
                             "frac": r["achieved_gbs"] / hbm}
            bpr[shape] = r
        if not args.no_cpu_baseline:
            bpr["cpu_baseline_c2"] = run_bpr_reference(os.cpu_count() or 1)
    evalr = None
    if rank == 0 and world == 1 and not args.no_bpr:
        evalr = {shape: run_eval_ours(shape, local_rank) for shape in ("c2", "large")}

    if rank == 0:
        line = {
            "metric": "wals_nnz_per_s", "value": value, "unit": "nnz/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "s_per_epoch": ms_per_step * 1e-3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "nusers": nu, "nitems": ni, "nnz": nnz, "nfactors": k, "alpha": ALPHA,
                       "lambda": LAMBDA, "parallelism": "rows x%d" % world, "exchange": sw.exchange if world > 1 else "none",
                       "l2": L2_NOTE},
            "loss": loss_value, "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "e2e": e2e,
            "cpu_baseline": cpu_baseline, "bpr": bpr, "eval": evalr, "lib": os.path.relpath(capi.LIB_PATH, ROOT),
        }
        print(json.dumps(line), flush=True)
    sw.close()
    if world > 1:
        dist.destroy_process_group()


def run_c5(args):
    """BASELINE configs[4]: WALS + all-user p@10 / AUC evaluation on the 10M x 1M x 1B power-law problem, one
    process per GPU, every rank generating ONLY its own shard (qmf_b200.datagen.powerlaw_shard_torch).
    --c5-scale s shrinks users, items and draws by s (1 GPU: 0.125)."""
    import torch
    import torch.distributed as dist
    from qmf_b200 import capi
    from qmf_b200.datagen import CONFIGS, powerlaw_shard_torch
    from qmf_b200.wals_dist import ShardedWals

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    nu, ni, draws, k = CONFIGS["c5"]
    sc = float(args.c5_scale)
    nu, ni, draws = int(nu * sc), int(ni * sc), int(draws * sc)
    t_gen = time.perf_counter()
    prob = powerlaw_shard_torch(nu, ni, draws, seed=20240505, device=device, rank=rank, world=world)
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t_gen
    nnz = prob["nnz"]
    sw = ShardedWals(nu, ni, k, prob["csr_user"], prob["csr_item"], device, rank, world, exchange=args.exchange,
                     ranges=prob["ranges"])
    g = torch.Generator(device=device).manual_seed(7)
    Y0 = (torch.rand(ni, k, generator=g, device=device, dtype=torch.float64) * 0.02 - 0.01)
    sw.set_factors(1, Y0)
    del Y0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        sw.epoch(ALPHA, LAMBDA)
    barrier()
    sw.check_error()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [[{n: torch.cuda.Event(enable_timing=True) for n in ("gram0", "solve0", "solve1")} for _ in range(2)]
          for _ in range(args.steps)]
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = sw.launches
    barrier()
    t0.record()
    loss = None
    for s in range(args.steps):
        loss = sw.epoch(ALPHA, LAMBDA, ev[s])
    t1.record()
    barrier()
    ms = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms.item()) / args.steps
    sw.check_error()
    loss_value = float(loss.item())
    launches = sw.launches - launches0
    solve_ms = [[e[h]["solve0"].elapsed_time(e[h]["solve1"]) for h in range(2)] for e in ev]
    fl_local = sum(algorithmic_flops(sw.shard[s]["end"] - sw.shard[s]["begin"], sw.shard[s]["nnz"], k) for s in (0, 1))
    solve_s = float(np.mean([sum(x) for x in solve_ms])) * 1e-3
    stats = torch.tensor([fl_local, solve_s, float(sw.shard[0]["nnz"]), float(sw.shard[1]["nnz"])], dtype=torch.float64, device=device)
    allst = [torch.zeros_like(stats) for _ in range(world)]
    if world > 1:
        dist.all_gather(allst, stats)
    else:
        allst = [stats]
    allst = [x.tolist() for x in allst]
    fl_total = sum(x[0] for x in allst)
    achieved_tf = fl_total / max(x[1] for x in allst) * 1e-12   # whole job: all ranks' flops / slowest rank's kernel time

    # ---- all-user evaluation: every rank scores ITS users against all items on its full replicas ----------
    ub, ue = sw.shard[0]["begin"], sw.shard[0]["end"]
    nT = ue - ub
    tu = torch.arange(ub, ue, device=device, dtype=torch.int32)
    lp = torch.arange(nT + 1, device=device, dtype=torch.int64)
    li = prob["test_items"][ub:ue].contiguous()
    cnt = torch.zeros(2 * nT, device=device, dtype=torch.int32)
    psc = torch.zeros(nT, device=device, dtype=torch.float64)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    kp = sw.kp
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    capi.check(capi.lib.qmfb_eval_rank_dev(st, sw.F[0].data_ptr(), kp, sw.F[1].data_ptr(), kp, ni, k, None, tu.data_ptr(), nT,
                                           lp.data_ptr(), li.data_ptr(), nT, 1, cnt.data_ptr(), psc.data_ptr()))
    b.record()
    barrier()
    ems = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(ems, op=dist.ReduceOp.MAX)
    eval_s = float(ems.item()) * 1e-3
    t_host = time.perf_counter()
    cnt_h = cnt.cpu().numpy()
    lp_h = np.arange(nT + 1, dtype=np.int64)
    sums = []
    for name in ("auc", "p@10"):
        out = np.empty(nT, dtype=np.float64)
        # host threads: this rank's share of the box's cores (8 ranks x all cores oversubscribed the host: 41.7 s -> see line)
        capi.check(capi.lib.qmfb_rank_metrics(name.encode(), cnt_h, lp_h, nT, ni, max(1, (os.cpu_count() or 1) // world), out))
        sums.append(float(out.sum()))
    t_host = time.perf_counter() - t_host
    tot = torch.tensor(sums + [float(cnt_h.sum())], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(tot)
    auc, p10 = float(tot[0].item()) / nu, float(tot[1].item()) / nu
    assert int(tot[2].item()) == nu * (ni - 1), "every (user, negative item) pair must land in exactly one bucket"
    clocks = sampler.stop() if rank == 0 else None
    eval_tf = 2.0 * nu * ni * k / eval_s * 1e-12
    if rank == 0:
        peak = fp64_peak_tflops() * world
        line = {
            "metric": "wals_nnz_per_s", "value": nnz / (ms_per_step * 1e-3), "unit": "nnz/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "s_per_epoch": ms_per_step * 1e-3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C5 power-law 10M x 1M, 1B draws, k=128 (scale %g)" % sc, "nusers": nu, "nitems": ni,
                       "nnz": nnz, "draws": draws, "nfactors": k, "alpha": ALPHA, "lambda": LAMBDA,
                       "parallelism": "rows x%d, each rank generates and holds only its shard" % world,
                       "exchange": sw.exchange if world > 1 else "none", "max_item_len": prob["max_item_len"],
                       "max_user_len": prob["max_user_len"], "l2": "inputs larger than the 126 MB L2; no flush"},
            "loss": loss_value, "gpu_launches": launches, "clocks": clocks, "generate_s": t_gen,
            "roofline": {"kernel": "wals_solve(_ws)_kernel<16> + long-row pre-pass (between the solve events)", "bound": "tensor",
                         "achieved": achieved_tf, "peak": peak, "unit": "TFLOP/s", "frac": achieved_tf / peak, "traffic": None,
                         "algorithmic_flops_per_epoch": fl_total,
                         "per_rank": [{"flops": x[0], "solve_s": x[1], "user_nnz": x[2], "item_nnz": x[3]} for x in allst],
                         "peak_source": "builder-measured FP64 DMMA peak x n_gpus (profiles/fp64_peak.json)"},
            "eval": {"test_users": nu, "nitems": ni, "seconds": eval_s, "host_metric_seconds": t_host, "auc": auc, "p@10": p10,
                     "tflops": eval_tf, "roofline": {"bound": "tensor", "achieved": eval_tf, "peak": peak, "unit": "TFLOP/s",
                                                     "frac": eval_tf / peak, "algorithmic_flops": 2.0 * nu * ni * k},
                     "note": "one held-out item per user from the popularity law; every rank scores its own users against all "
                             "items on its replicas (qmfb_eval_rank_dev), metric sums allreduced"},
            "cpu_baseline": None, "e2e": None,
            "note": "the reference cannot run C5 (1 B-line text parse, ~80 GB of AoS signals, dense nT x ni score matrix): "
                    "no CPU arm; SURVEY.md 8d",
            "lib": os.path.relpath(capi.LIB_PATH, ROOT),
        }
        print(json.dumps(line), flush=True)
    sw.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=["c1", "c3", "c4", "c5"])
    ap.add_argument("--c5-scale", type=float, default=1.0, help="c5 only: shrink users, items and draws by this factor")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-bpr", action="store_true")
    ap.add_argument("--exchange", default="auto", choices=["auto", "p2p", "nccl"],
                    help="N > 1: how the solved shards reach the other ranks (p2p = peer stores fused into the solve kernel)")
    args = ap.parse_args()
    if args.workload == "c5":
        if args.impl == "reference":
            if int(os.environ.get("RANK", "0")) == 0:
                print(json.dumps({"impl": "reference", "unavailable": "the reference cannot run C5 (1 B-line text parse, ~80 GB of "
                                  "AoS signals, dense score matrix; SURVEY.md 8d)"}), flush=True)
            return
        return run_c5(args)
    from qmf_b200.datagen import CONFIGS
    cfg = CONFIGS[args.workload]
    names = {"c1": "C1 uniform 10k x 5k, 500k nnz, k=30", "c3": "C3 MovieLens-20M-shaped 138k x 27k, 20M nnz, k=64",
             "c4": "C4 Netflix-shaped 480k x 17.8k, 100M nnz, k=128"}
    if args.impl == "reference":
        run_reference(args, cfg, names[args.workload])
    else:
        run_ours(args, cfg, names[args.workload])


if __name__ == "__main__":
    main()
