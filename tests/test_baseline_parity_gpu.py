"""Parity at the BASELINE.json config sizes against the LIVE reference (oracle/_ref, the unmodified
reference sources compiled by oracle/Makefile; it travels to the GPU box with the snapshot).

  C1  10k x 5k, 500k nnz, k=30, 10 epochs : every epoch's loss (1e-12) and both factor matrices
      (per-row relative 1e-9) against WALSEngine::iterate of the reference; indexing / CSR bit-exact.
  C3  138k x 27k, 20M nnz, k=64           : sampled-row parity (below).
  C4  480k x 17.8k, 100M nnz, k=128       : sampled-row parity.

Sampled-row parity (the reference needs ~40 core-minutes for one C4 epoch): after the GPU's half-step,
2 000 random rows plus the longest and the shortest rows of that side go through the reference's own
static WALSEngine::updateFactorsForOne (qmf/wals/WALSEngine.cpp:266-310) with the SAME fixed-side factors
and the reference's own Gram of them (computeXtX, :246-264); the solved rows must agree per row to
1e-9 relative and the row loss terms to 1e-11 relative.
"""
import numpy as np
import pytest

from util import init_factors, rel_err_rows, uniform_dataset

pytestmark = pytest.mark.gpu

ALPHA, LAMBDA = 40.0, 0.05   # reference defaults, qmf/wals.cpp:28-29
FACTOR_TOL = 1e-9            # per row, relative (north_star)
LOSS_TOL = 1e-12             # objective, relative (north_star)
ROW_LOSS_TOL = 1e-11         # single row loss terms, relative


def test_c1_full_size_ten_epochs_match_reference(ref_lib):
    from qmf_b200 import WalsEngineHandle
    from qmf_b200.wals import Signals
    nu, ni, nnz, k, nepochs = 10_000, 5_000, 500_000, 30, 10
    u, i, v = uniform_dataset(nu, ni, nnz, seed=20240501 + 1)
    R = ref_lib
    rh = R.ref_wals_create(k, nepochs, LAMBDA, ALPHA, 16, None, 0, 0, 42)
    R.ref_wals_init(rh, u, i, v, nnz)
    NU, NI = R.ref_wals_nusers(rh), R.ref_wals_nitems(rh)
    sig = Signals(u, i, v)
    assert (sig.nusers, sig.nitems, sig.nnz) == (NU, NI, nnz)
    # raw-id order and both CSR orientations bit-exact against the reference's own groups
    for side, n in ((0, NU), (1, NI)):
        ids = np.empty(n, np.int64)
        R.ref_wals_ids(rh, side, ids)
        assert np.array_equal(sig.ids(side), ids)
        rp, rid, ci, cid, val = np.empty(n + 1, np.int64), np.empty(n, np.int64), np.empty(nnz, np.int32), np.empty(nnz, np.int64), np.empty(nnz)
        R.ref_wals_csr(rh, side, rp, rid, ci, cid, val)
        grp, gci, gval, _ = sig.csr(side)
        assert np.array_equal(grp, rp) and np.array_equal(gci, ci) and np.array_equal(gval, val)
    Y0 = init_factors(NI, k, seed=7)
    R.ref_wals_set_factors(rh, 1, Y0)
    h = WalsEngineHandle(NU, NI, k)
    h.set_signals(sig)
    h.set_factors(1, Y0)
    Xr, Yr = np.empty((NU, k)), np.empty((NI, k))
    for epoch in range(nepochs):
        lu, lu_r = h.half_step(0, ALPHA, LAMBDA), R.ref_wals_half_step(rh, 0)
        li, li_r = h.half_step(1, ALPHA, LAMBDA), R.ref_wals_half_step(rh, 1)
        R.ref_wals_get_factors(rh, 0, Xr)
        R.ref_wals_get_factors(rh, 1, Yr)
        assert abs(lu - lu_r) <= LOSS_TOL * abs(lu_r), (epoch, lu, lu_r)
        assert abs(li - li_r) <= LOSS_TOL * abs(li_r), (epoch, li, li_r)
        ex, ey = rel_err_rows(h.get_factors(0), Xr), rel_err_rows(h.get_factors(1), Yr)
        assert ex < FACTOR_TOL and ey < FACTOR_TOL, (epoch, ex, ey)
    R.ref_wals_destroy(rh)
    h.close()
    sig.close()


def _sample_rows(row_ptr_dev, nsample, seed):
    """2 000 random rows + the 4 longest + the 4 shortest (host int64 array, ascending)"""
    import torch
    lens = (row_ptr_dev[1:] - row_ptr_dev[:-1])
    n = lens.numel()
    g = np.random.default_rng(seed)
    pick = set(g.choice(n, size=min(nsample, n), replace=False).tolist())
    order = torch.argsort(lens, stable=True)
    pick.update(order[:4].tolist())
    pick.update(order[-4:].tolist())
    return np.array(sorted(pick), dtype=np.int64)


def _check_side(R, sw, side, csr, rows, k, threads):
    """rows of `side` just solved by the GPU vs the reference's updateFactorsForOne on the same inputs"""
    import torch
    rp, col, val = csr
    Y = sw.get_factors(1 - side).cpu().numpy().copy()          # the fixed side the GPU solved against
    G = np.zeros((k, k))
    R.ref_gram(Y, Y.shape[0], k, 1, 1, G)                      # the reference's own computeXtX (serial)
    rows_d = torch.as_tensor(rows, device=rp.device)
    starts, ends = rp[rows_d], rp[rows_d + 1]
    lens = (ends - starts)
    lrp = np.zeros(len(rows) + 1, np.int64)
    np.cumsum(lens.cpu().numpy(), out=lrp[1:])
    idx = torch.cat([torch.arange(int(s), int(e), device=rp.device) for s, e in zip(starts.tolist(), ends.tolist())])
    lcol, lval = col[idx].cpu().numpy().astype(np.int32), val[idx].cpu().numpy()
    Xr, lr = np.zeros((len(rows), k)), np.zeros(len(rows))
    R.ref_wals_update_rows_losses(Xr, len(rows), Y, Y.shape[0], k, lrp, lcol, lval, G, ALPHA, LAMBDA, threads, lr)
    Xg = sw.get_factors(side)[rows_d].cpu().numpy()
    lg = sw.row_loss[rows_d - sw.shard[side]["begin"]].cpu().numpy()
    ex = rel_err_rows(Xg, Xr)
    el = float(np.max(np.abs(lg - lr) / np.maximum(np.abs(lr), 1e-300)))
    return ex, el, int(lens.max()), int(lens.min())


@pytest.mark.parametrize("name", ["c3", "c4"])
def test_sampled_rows_match_reference_at_full_size(ref_lib, name):
    import os
    import torch
    from qmf_b200.datagen import CONFIGS, init_item_factors, uniform_csr_torch
    from qmf_b200.wals_dist import ShardedWals
    nu, ni, nnz, k = CONFIGS[name]
    dev = torch.device("cuda", 0)
    csr_user, csr_item = uniform_csr_torch(nu, ni, nnz, seed=20240501, device=dev)
    sw = ShardedWals(nu, ni, k, csr_user, csr_item, dev)
    sw.set_factors(1, init_item_factors(ni, k, seed=7))
    threads = os.cpu_count() or 1
    # second epoch: both sides then carry trained (not freshly initialised) factors
    sw.epoch(ALPHA, LAMBDA)
    sw.half_step(0, ALPHA, LAMBDA)
    torch.cuda.synchronize()
    sw.check_error()
    ex, el, lmax, lmin = _check_side(ref_lib, sw, 0, csr_user, _sample_rows(csr_user[0], 2000, 1), k, threads)
    assert ex < FACTOR_TOL and el < ROW_LOSS_TOL, ("user rows", ex, el, lmax, lmin)
    sw.half_step(1, ALPHA, LAMBDA)
    torch.cuda.synchronize()
    sw.check_error()
    ex, el, lmax, lmin = _check_side(ref_lib, sw, 1, csr_item, _sample_rows(csr_item[0], 2000, 2), k, threads)
    assert ex < FACTOR_TOL and el < ROW_LOSS_TOL, ("item rows", ex, el, lmax, lmin)
    sw.close()
