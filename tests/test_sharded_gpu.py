"""The single-process multi-GPU WALS engine (qmfb_wals_sharded_*, what `wals --ngpus N` binds) against
the one-GPU engine: factors and loss must be BIT-IDENTICAL for every shard count (fixed Gram parts summed
in part order, row losses summed in global row order).  On a one-GPU box the shards share that GPU
(a device may be listed several times), which exercises the cuts, the cross-shard Gram reduce and the
fused peer stores; with >= 2 GPUs the same tests run over NVLink peer memory."""
import os
import subprocess

import numpy as np
import pytest

from util import init_factors, uniform_dataset

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "qmf_b200", "host", "bin")
CLI = os.path.join(ROOT, "tests", "golden", "cli")
ALPHA, LAMBDA = 40.0, 0.05


def _ngpus():
    from qmf_b200 import capi
    return max(int(capi.lib.qmfb_device_count()), 0)


def _problem(nu, ni, nnz, k, seed):
    from qmf_b200 import csr_from_coo
    u, i, v = uniform_dataset(nu, ni, nnz, seed, id_scale=(3, 5), dup=7)
    uids, urp, uci, uv = csr_from_coo(u, i, v)
    iids, irp, ici, iv = csr_from_coo(i, u, v)
    return (u, i, v), (urp, uci, uv), (irp, ici, iv), len(uids), len(iids), init_factors(len(iids), k, seed=5)


def _run_single(NU, NI, k, ucsr, icsr, Y0, epochs):
    from qmf_b200 import WalsEngineHandle
    h = WalsEngineHandle(NU, NI, k)
    h.set_csr(0, *ucsr)
    h.set_csr(1, *icsr)
    h.set_factors(1, Y0)
    out = []
    for _ in range(epochs):
        lu = h.half_step(0, ALPHA, LAMBDA)
        li = h.half_step(1, ALPHA, LAMBDA)
        out.append((lu, li, h.get_factors(0), h.get_factors(1)))
    h.close()
    return out


@pytest.mark.parametrize("nu,ni,nnz,k,nshards", [(700, 500, 30000, 30, 3), (900, 400, 40000, 128, 2), (500, 450, 20000, 128, 8),
                                                 (300, 260, 9000, 160, 3), (40, 30, 300, 64, 5)])
def test_shards_sharing_one_gpu_are_bit_identical_to_one_gpu(nu, ni, nnz, k, nshards):
    from qmf_b200.wals import ShardedWalsHandle
    _, ucsr, icsr, NU, NI, Y0 = _problem(nu, ni, nnz, k, seed=nu + k)
    want = _run_single(NU, NI, k, ucsr, icsr, Y0, 2)
    sw = ShardedWalsHandle(NU, NI, k, [0] * nshards)
    sw.set_csr(0, *ucsr)
    sw.set_csr(1, *icsr)
    # the shards tile the rows, balanced by nnz
    for side, n in ((0, NU), (1, NI)):
        pos = 0
        for slot in range(nshards):
            _, b, nr, _ = sw.shard(slot, side)
            assert b == pos
            pos += nr
        assert pos == n
    sw.set_factors(1, Y0)
    for e in range(2):
        lu = sw.half_step(0, ALPHA, LAMBDA)
        li = sw.half_step(1, ALPHA, LAMBDA)
        assert (lu, li) == want[e][:2]
        for slot in range(nshards):                      # every replica
            assert np.array_equal(sw.get_factors(0, slot), want[e][2]), (e, slot)
            assert np.array_equal(sw.get_factors(1, slot), want[e][3]), (e, slot)
    sw.close()


def test_sharded_from_gpu_ingest_and_epoch_host():
    """set_signals (device-to-device shard copies from the ingest handle) + the host-buffer epoch call"""
    from qmf_b200.wals import ShardedWalsHandle, Signals
    k = 64
    (u, i, v), ucsr, icsr, NU, NI, Y0 = _problem(800, 600, 50000, k, seed=77)
    want = _run_single(NU, NI, k, ucsr, icsr, Y0, 2)
    sig = Signals(u, i, v)
    n = max(1, min(_ngpus(), 4))
    devices = list(range(n)) if n > 1 else [0, 0, 0]
    sw = ShardedWalsHandle(NU, NI, k, devices)
    sw.set_signals(sig)
    sig.close()
    X, Y = np.empty((NU, k)), np.empty((NI, k))
    loss = sw.epoch_host(ALPHA, LAMBDA, Y0, X, Y)
    assert loss == want[0][1] and np.array_equal(X, want[0][2]) and np.array_equal(Y, want[0][3])
    loss = sw.epoch_host(ALPHA, LAMBDA, Y.copy(), X, Y)
    assert loss == want[1][1] and np.array_equal(X, want[1][2]) and np.array_equal(Y, want[1][3])
    assert sw.launch_count() > 0
    sw.close()


def test_replicas_on_every_gpu_are_bit_identical_to_one_gpu():
    """>= 2 GPUs: peer stores over NVLink, Gram parts read from peer memory"""
    n = _ngpus()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    from qmf_b200.wals import ShardedWalsHandle
    k = 128
    _, ucsr, icsr, NU, NI, Y0 = _problem(6000, 2500, 400000, k, seed=9)
    want = _run_single(NU, NI, k, ucsr, icsr, Y0, 3)
    sw = ShardedWalsHandle(NU, NI, k, list(range(n)))
    sw.set_csr(0, *ucsr)
    sw.set_csr(1, *icsr)
    sw.set_factors(1, Y0)
    for e in range(3):
        lu = sw.half_step(0, ALPHA, LAMBDA)
        li = sw.half_step(1, ALPHA, LAMBDA)
        assert (lu, li) == want[e][:2]
        for slot in range(n):
            assert np.array_equal(sw.get_factors(0, slot), want[e][2]), (e, slot)
            assert np.array_equal(sw.get_factors(1, slot), want[e][3]), (e, slot)
    sw.close()


def test_wals_binary_ngpus_writes_the_same_files(tmp_path):
    """`wals --ngpus N` reproduces the one-GPU factor files byte for byte (and its log lines)"""
    n = _ngpus()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    outs = []
    for ngpus in (1, n):
        uf, itf = str(tmp_path / ("u%d.txt" % ngpus)), str(tmp_path / ("i%d.txt" % ngpus))
        cmd = [os.path.join(BIN, "wals"), "--nepochs=3", "--nfactors=30", "--train_dataset=" + os.path.join(CLI, "train.txt"),
               "--test_dataset=" + os.path.join(CLI, "test.txt"), "--distribution_file=" + os.path.join(CLI, "dist.txt"),
               "--test_avg_metrics=auc,p@10", "--ngpus=%d" % ngpus, "--user_factors=" + uf, "--item_factors=" + itf]
        r = subprocess.run(cmd, env=dict(os.environ, QMF_LOG_PRECISION="17"), capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        lines = [ln.split("] ", 1)[-1] for ln in r.stderr.splitlines() if "train loss" in ln or "recorded metric" in ln]
        outs.append((open(uf, "rb").read(), open(itf, "rb").read(), lines))
    assert outs[0][0] == outs[1][0] and outs[0][1] == outs[1][1]
    assert outs[0][2] == outs[1][2] and len(outs[0][2]) == 5
