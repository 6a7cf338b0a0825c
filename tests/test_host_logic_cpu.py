"""CPU tests of the host-side logic around the kernels: CSR construction, the metric arithmetic
that consumes the GPU's rank statistics, shard balancing, and the 2-rank exchange logic of the
sharded WALS driver (gloo, with the CPU oracle standing in for the CUDA kernels)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from util import init_factors, rel_err, rel_err_rows, uniform_dataset

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_csr_from_coo_matches_reference_fixture():
    from qmf_b200.wals import csr_from_coo
    g = np.load(os.path.join(GOLD, "wals_k30.npz"))
    for side, (r, c) in enumerate(((g["u"], g["i"]), (g["i"], g["u"]))):
        ids, rp, ci, va = csr_from_coo(r, c, g["v"])
        assert np.array_equal(ids, g["rid%d" % side]) and np.array_equal(rp, g["rp%d" % side])
        assert np.array_equal(ci, g["ci%d" % side])
        for row in range(len(ids)):   # duplicates may be permuted by the reference's unstable sort
            sl = slice(rp[row], rp[row + 1])
            assert sorted(zip(ci[sl], va[sl])) == sorted(zip(g["ci%d" % side][sl], g["va%d" % side][sl]))


def _cnt_cpu(labels, scores):
    pos = np.sort(scores[labels > 0])
    neg = scores[labels <= 0]
    idx = np.searchsorted(pos, neg, side="left")       # positives strictly below each negative
    return np.bincount(idx, minlength=len(pos) + 1).astype(np.int32)


def test_metric_arithmetic_from_rank_statistics(oracle_lib):
    from qmf_b200.evalrank import user_metrics
    rng = np.random.default_rng(5)
    names = ["auc", "ap", "p@1", "p@7", "r@7", "r@20"]
    for trial in range(60):
        n = int(rng.integers(25, 200))
        labels = (rng.uniform(size=n) < rng.uniform(0.02, 0.5)).astype(float)
        labels[rng.integers(0, n)] = 1.0
        scores = np.round(rng.normal(size=n), 1 if trial % 2 else 6)
        got = user_metrics(_cnt_cpu(labels, scores), n, names)
        for name in names:
            kind, k = oracle.metric_kind(name)
            assert got[name] == oracle_lib.qmfo_metric_one(kind, k, labels, scores, n), (trial, name)
    # AUC degenerate classes return 1.0 (Metrics.cpp:80-83)
    assert user_metrics(np.array([5], np.int32), 5, ["auc"])["auc"] == 1.0
    assert user_metrics(np.zeros(4, np.int32), 3, ["auc"])["auc"] == 1.0


def test_balanced_row_ranges():
    from qmf_b200.wals_dist import balanced_row_ranges
    rng = np.random.default_rng(1)
    lens = rng.zipf(1.5, size=5000).clip(1, 4000)
    rp = np.zeros(5001, np.int64)
    np.cumsum(lens, out=rp[1:])
    for world in (1, 2, 3, 8):
        rr = balanced_row_ranges(rp, world)
        assert rr[0][0] == 0 and rr[-1][1] == 5000 and all(a[1] == b[0] for a, b in zip(rr, rr[1:]))
        nnz = [rp[e] - rp[b] for b, e in rr]
        assert max(nnz) <= rp[-1] / world + lens.max()


class OracleKernels:
    """CPU stand-in for the CUDA kernels (dense k x k Gram as the 'packed' form) — exercises only
    the sharding / exchange logic of ShardedWals."""
    launches_per_half_step = 4

    def __init__(self):
        self.O = oracle.oracle()

    def padded_k(self, k):
        return k

    def gram_packed_len(self, k):
        return k * k

    def gram_workspace_len(self, k):
        return 1

    def gram(self, Y, rb, re, k, ws, out):
        G = np.zeros((k, k))
        if re > rb:
            self.O.qmfo_gram(np.ascontiguousarray(Y[rb:re].numpy()), re - rb, k, G)
        out.copy_(torch.from_numpy(G.reshape(-1)))

    def solve(self, X, row_offset, Y, k, row_ptr, col, val, order, gram, alpha, lam, row_loss, loss_sum, scratch, nnz=-1):
        G = np.ascontiguousarray(gram.numpy().reshape(k, k))
        Yn = np.ascontiguousarray(Y.numpy())
        rp, c, v = row_ptr.numpy(), col.numpy(), val.numpy()
        total = 0.0
        x = np.zeros(k)
        for r in range(len(rp) - 1):
            sl = slice(rp[r], rp[r + 1])
            total += self.O.qmfo_wals_update_row(Yn, k, np.ascontiguousarray(c[sl]), np.ascontiguousarray(v[sl]),
                                                 rp[r + 1] - rp[r], G, alpha, lam, x)
            X[row_offset + r] = torch.from_numpy(x.copy())
        loss_sum[0] = total


def _problem():
    from qmf_b200.wals import csr_from_coo
    u, i, v = uniform_dataset(90, 60, 1500, 8, id_scale=(2, 5))
    uids, urp, uci, uv = csr_from_coo(u, i, v)
    iids, irp, ici, iv = csr_from_coo(i, u, v)
    return (len(uids), len(iids), 12, (urp, uci, uv), (irp, ici, iv), init_factors(len(iids), 12, 2))


def _rank_main(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from qmf_b200.wals_dist import ShardedWals
    NU, NI, k, ucsr, icsr, Y0 = _problem()
    t = lambda a: tuple(torch.from_numpy(np.ascontiguousarray(x)) for x in a)
    sw = ShardedWals(NU, NI, k, t(ucsr), t(icsr), torch.device("cpu"), rank, world, kernels=OracleKernels())
    sw.set_factors(1, Y0)
    losses = [float(sw.epoch(40.0, 0.05)) for _ in range(2)]
    if rank == 0:
        np.savez(out_path, X=sw.get_factors(0).numpy(), Y=sw.get_factors(1).numpy(), losses=np.array(losses),
                 ranges=np.array(sw.ranges[0]))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_wals_two_ranks_gloo(tmp_path, oracle_lib):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "r0.npz")
    mp.spawn(_rank_main, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    assert 0 < got["ranges"][0][1] < 90                # really split in two shards
    NU, NI, k, ucsr, icsr, Y0 = _problem()
    X, Y = np.zeros((NU, k)), Y0.copy()
    losses = []
    for _ in range(2):
        oracle_lib.qmfo_wals_half_step(X, NU, Y, NI, k, *ucsr, 40.0, 0.05, NU, NI, 1)
        losses.append(oracle_lib.qmfo_wals_half_step(Y, NI, X, NU, k, *icsr, 40.0, 0.05, NU, NI, 1))
    # the 2-rank Gram is the sum of two partial Grams (different association): ~1e-16 relative
    assert rel_err_rows(got["X"], X) < 1e-11 and rel_err_rows(got["Y"], Y) < 1e-11
    assert np.allclose(got["losses"], losses, rtol=1e-12, atol=0)


# ---- C5 path: every rank generates only its shard (qmf_b200.datagen.powerlaw_shard_torch) ---------------------
_PL = dict(nusers=3000, nitems=400, draws=40_000, seed=11, chunk=9_000, k=8)


def test_powerlaw_shards_concatenate_to_the_full_problem():
    """the per-rank shards of the streamed power-law generator are exactly the row ranges of the 1-rank problem,
    both orientations are transposes of each other, and the popularity law is heavy-tailed"""
    from qmf_b200.datagen import powerlaw_shard_torch
    a = _PL
    full = powerlaw_shard_torch(a["nusers"], a["nitems"], a["draws"], a["seed"], "cpu", 0, 1, chunk_draws=a["chunk"])
    parts = [powerlaw_shard_torch(a["nusers"], a["nitems"], a["draws"], a["seed"], "cpu", r, 3, chunk_draws=a["chunk"])
             for r in range(3)]
    for key in ("csr_user", "csr_item"):
        for j in (1, 2):
            assert torch.equal(torch.cat([p[key][j] for p in parts]), full[key][j])
    assert all(p["ranges"] == parts[0]["ranges"] and p["nnz"] == full["nnz"] for p in parts)
    assert torch.equal(parts[2]["test_items"], full["test_items"])
    urp, ucol, uval = full["csr_user"]
    irp, icol, ival = full["csr_item"]
    u = torch.repeat_interleave(torch.arange(a["nusers"]), urp[1:] - urp[:-1])
    i = torch.repeat_interleave(torch.arange(a["nitems"]), irp[1:] - irp[:-1])
    ka, kb = u * a["nitems"] + ucol.long(), icol.long() * a["nitems"] + i
    oa, ob = torch.argsort(ka), torch.argsort(kb)
    assert torch.equal(ka[oa], kb[ob]) and torch.equal(uval[oa], ival[ob])
    assert len(torch.unique(ka)) == len(ka)                    # distinct cells
    ilen = (irp[1:] - irp[:-1]).sort(descending=True).values
    assert ilen[0] > 20 * ilen[len(ilen) // 2]                 # blockbusters vs the median item
    assert set(uval.unique().tolist()) <= {1.0, 2.0, 3.0, 4.0, 5.0}


def _rank_main_presharded(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from qmf_b200.datagen import powerlaw_shard_torch
    from qmf_b200.wals_dist import ShardedWals
    a = _PL
    p = powerlaw_shard_torch(a["nusers"], a["nitems"], a["draws"], a["seed"], "cpu", rank, world, chunk_draws=a["chunk"])
    sw = ShardedWals(a["nusers"], a["nitems"], a["k"], p["csr_user"], p["csr_item"], torch.device("cpu"), rank, world,
                     kernels=OracleKernels(), ranges=p["ranges"])
    sw.set_factors(1, init_factors(a["nitems"], a["k"], 2))
    losses = [float(sw.epoch(40.0, 0.05)) for _ in range(2)]
    sw.check_error()
    if rank == 0:
        np.savez(out_path, X=sw.get_factors(0).numpy(), Y=sw.get_factors(1).numpy(), losses=np.array(losses))
    dist.barrier()
    dist.destroy_process_group()


def test_presharded_wals_two_ranks_gloo(tmp_path, oracle_lib):
    """ShardedWals(ranges=...) on shards each rank generated for itself == the oracle on the full problem"""
    from qmf_b200.datagen import powerlaw_shard_torch
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "r0.npz")
    mp.spawn(_rank_main_presharded, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    a = _PL
    full = powerlaw_shard_torch(a["nusers"], a["nitems"], a["draws"], a["seed"], "cpu", 0, 1, chunk_draws=a["chunk"])
    ucsr = tuple(np.ascontiguousarray(x.numpy()) for x in full["csr_user"])
    icsr = tuple(np.ascontiguousarray(x.numpy()) for x in full["csr_item"])
    NU, NI, k = a["nusers"], a["nitems"], a["k"]
    X, Y = np.zeros((NU, k)), init_factors(NI, k, 2).copy()
    losses = []
    for _ in range(2):
        oracle_lib.qmfo_wals_half_step(X, NU, Y, NI, k, *ucsr, 40.0, 0.05, NU, NI, 1)
        losses.append(oracle_lib.qmfo_wals_half_step(Y, NI, X, NU, k, *icsr, 40.0, 0.05, NU, NI, 1))
    assert rel_err_rows(got["X"], X) < 1e-11 and rel_err_rows(got["Y"], Y) < 1e-11
    assert np.allclose(got["losses"], losses, rtol=1e-12, atol=0)


class _FailingKernels(OracleKernels):
    """raises the NOT_SPD flag in the USER half-step on rank 1 only (the launcher clears scratch before every solve)"""

    def __init__(self, rank):
        super().__init__()
        self.rank, self.calls = rank, 0

    def solve(self, X, row_offset, Y, k, row_ptr, col, val, order, gram, alpha, lam, row_loss, loss_sum, scratch, nnz=-1):
        scratch.zero_()
        super().solve(X, row_offset, Y, k, row_ptr, col, val, order, gram, alpha, lam, row_loss, loss_sum, scratch, nnz)
        if self.rank == 1 and self.calls == 0:
            scratch[1] = 1
        self.calls += 1


def _rank_main_not_spd(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from qmf_b200.wals_dist import ShardedWals
    NU, NI, k, ucsr, icsr, Y0 = _problem()
    t = lambda a: tuple(torch.from_numpy(np.ascontiguousarray(x)) for x in a)
    sw = ShardedWals(NU, NI, k, t(ucsr), t(icsr), torch.device("cpu"), rank, world, kernels=_FailingKernels(rank))
    sw.set_factors(1, Y0)
    sw.epoch(40.0, 0.05)           # the flag is raised in the user half-step and must survive the item half-step
    raised = False
    try:
        sw.check_error()
    except RuntimeError:
        raised = True
    again = False
    try:
        sw.check_error()           # cleared by the first check
    except RuntimeError:
        again = True
    open(os.path.join(out_dir, "r%d.txt" % rank), "w").write("%d %d" % (raised, again))
    dist.barrier()
    dist.destroy_process_group()


def test_not_spd_flag_is_sticky_and_raised_on_every_rank(tmp_path, oracle_lib):
    """a non-positive pivot in the user half-step on ONE rank must not be lost when the launcher clears the
    flag for the item half-step, and every rank must raise (the flag travels in the loss allreduce)"""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_rank_main_not_spd, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert open(str(tmp_path / ("r%d.txt" % r))).read() == "1 0"


def test_powerlaw_degrees_follow_the_stated_law():
    """rank-size law with exponent 1/1.1, clipped to [1, 1e5], scaled to the requested number of draws, seeded"""
    from qmf_b200.datagen import powerlaw_degrees
    d = powerlaw_degrees(200_000, 20_000_000, seed=3)
    assert d.min() >= 1 and d.max() <= 100_000 and abs(int(d.sum()) - 20_000_000) < 200_000 // 2
    assert np.array_equal(d, powerlaw_degrees(200_000, 20_000_000, seed=3))
    s = np.sort(d)[::-1].astype(np.float64)
    j = np.array([1000, 10_000, 100_000])                     # unclipped part of the law: d(j) ~ j^(-1/1.1)
    slope = np.polyfit(np.log(j), np.log(s[j - 1]), 1)[0]
    assert abs(slope + 1 / 1.1) < 0.02, slope
