"""GPU parity of the WALS half-step (Gram, per-row build + solve, loss) against the CPU oracle,
through the C ABI.  Tolerances are the north_star's: 1e-9 relative on factors, 1e-12 relative on
the objective."""
import numpy as np
import pytest

from util import init_factors, rel_err, rel_err_rows, uniform_dataset

pytestmark = pytest.mark.gpu

FACTOR_TOL = 1e-9
LOSS_TOL = 1e-12


def _setup(nu, ni, nnz, k, seed, dup=0):
    from qmf_b200 import WalsEngineHandle, csr_from_coo
    u, i, v = uniform_dataset(nu, ni, nnz, seed, id_scale=(7, 3), dup=dup)
    uids, urp, uci, uv = csr_from_coo(u, i, v)
    iids, irp, ici, iv = csr_from_coo(i, u, v)
    h = WalsEngineHandle(len(uids), len(iids), k)
    h.set_csr(0, urp, uci, uv)
    h.set_csr(1, irp, ici, iv)
    return h, (urp, uci, uv), (irp, ici, iv), len(uids), len(iids)


@pytest.mark.parametrize("n,k", [(1, 30), (17, 5), (1000, 30), (5000, 64), (3001, 100), (20000, 128), (40, 1), (2500, 80), (999, 96),
                                 (3000, 129), (777, 160), (5000, 200), (4097, 256)])
def test_gram_matches_oracle(oracle_lib, n, k):
    from qmf_b200 import WalsEngineHandle
    Y = init_factors(n, k, seed=n + k, bound=0.5)
    h = WalsEngineHandle(n, 1, k)
    h.set_factors(0, Y)
    G = h.gram(0)
    Go = np.zeros((k, k))
    oracle_lib.qmfo_gram(Y, n, k, Go)
    assert rel_err(G, Go) < 1e-13
    assert np.array_equal(G, G.T)


@pytest.mark.parametrize("nu,ni,nnz,k,dup", [(300, 200, 6000, 30, 0), (200, 150, 4000, 64, 25), (400, 300, 12000, 128, 0),
                                              (260, 210, 5000, 100, 0), (50, 40, 300, 8, 5),
                                              (240, 200, 5000, 80, 0), (250, 200, 5000, 96, 3), (100, 90, 2000, 33, 0),
                                              # 128 < k <= 256: workspace-resident tiles (wals_big.cuh)
                                              (420, 340, 9000, 160, 0), (600, 520, 16000, 256, 10), (450, 400, 9000, 200, 0)])
def test_epochs_match_oracle(oracle_lib, nu, ni, nnz, k, dup):
    h, ucsr, icsr, NU, NI = _setup(nu, ni, nnz, k, seed=11 * k, dup=dup)
    Y0 = init_factors(NI, k, seed=5)
    h.set_factors(1, Y0)
    X = np.zeros((NU, k))
    Y = Y0.copy()
    alpha, lam = 40.0, 0.05
    for epoch in range(3):
        lu = h.half_step(0, alpha, lam)
        li = h.half_step(1, alpha, lam)
        lu_o = oracle_lib.qmfo_wals_half_step(X, NU, Y, NI, k, ucsr[0], ucsr[1], ucsr[2], alpha, lam, NU, NI, 16)
        li_o = oracle_lib.qmfo_wals_half_step(Y, NI, X, NU, k, icsr[0], icsr[1], icsr[2], alpha, lam, NU, NI, 16)
        Xg, Yg = h.get_factors(0), h.get_factors(1)
        assert rel_err_rows(Xg, X) < FACTOR_TOL, (epoch, "user factors")
        assert rel_err_rows(Yg, Y) < FACTOR_TOL, (epoch, "item factors")
        assert abs(lu - lu_o) <= LOSS_TOL * abs(lu_o), (epoch, lu, lu_o)
        assert abs(li - li_o) <= LOSS_TOL * abs(li_o), (epoch, li, li_o)


def test_rank_deficient_gram_backward_error(oracle_lib):
    """Fewer rows than factors: G = Y^T Y is rank deficient, A = G + ... + lambda I has condition
    number ~1e6, so the Cholesky (GPU) and Bunch-Kaufman (reference dsysv) solutions agree only to
    cond * eps.  Parity bar here: forward error 1e-6 and a normal-equation residual no worse
    than 4x the oracle's."""
    h, ucsr, icsr, NU, NI = _setup(120, 90, 3000, 128, seed=1408)
    k, alpha, lam = 128, 40.0, 0.05
    Y0 = init_factors(NI, k, seed=5)
    h.set_factors(1, Y0)
    X = np.zeros((NU, k))
    Y = Y0.copy()
    h.half_step(0, alpha, lam)
    oracle_lib.qmfo_wals_half_step(X, NU, Y, NI, k, ucsr[0], ucsr[1], ucsr[2], alpha, lam, NU, NI, 16)
    Xg = h.get_factors(0)
    assert rel_err_rows(Xg, X) < 1e-9
    h.half_step(1, alpha, lam)
    oracle_lib.qmfo_wals_half_step(Y, NI, X, NU, k, icsr[0], icsr[1], icsr[2], alpha, lam, NU, NI, 16)
    Yg = h.get_factors(1)
    assert rel_err_rows(Yg, Y) < 1e-6
    G = X.T @ X
    worst_g = worst_o = 0.0
    for r in range(NI):
        sl = slice(icsr[0][r], icsr[0][r + 1])
        Ys, w = X[icsr[1][sl]], icsr[2][sl]
        A = G + (Ys * (alpha * w)[:, None]).T @ Ys + lam * np.eye(k)
        b = ((1 + alpha * w)[:, None] * Ys).sum(0)
        worst_g = max(worst_g, np.abs(A @ Yg[r] - b).max() / np.abs(b).max())
        worst_o = max(worst_o, np.abs(A @ Y[r] - b).max() / np.abs(b).max())
    assert worst_g < max(4 * worst_o, 1e-12), (worst_g, worst_o)


def test_row_shapes_edge_cases(oracle_lib):
    """rows with 1 signal, exactly 16/17/32 signals (chunk boundaries), an empty row, a very long row"""
    from qmf_b200 import WalsEngineHandle
    k, ni = 30, 400
    rng = np.random.default_rng(3)
    lens = [1, 16, 17, 32, 0, 15, 33, 400, 2]
    row_ptr = np.zeros(len(lens) + 1, np.int64)
    np.cumsum(lens, out=row_ptr[1:])
    col = np.concatenate([np.sort(rng.choice(ni, n, replace=False)) for n in lens]).astype(np.int32)
    val = rng.integers(1, 6, size=len(col)).astype(np.float64)
    Y = init_factors(ni, k, seed=9, bound=0.3)
    h = WalsEngineHandle(len(lens), ni, k)
    h.set_csr(0, row_ptr, col, val)
    h.set_factors(1, Y)
    loss = h.half_step(0, 40.0, 0.05)
    X = np.zeros((len(lens), k))
    loss_o = oracle_lib.qmfo_wals_half_step(X, len(lens), Y, ni, k, row_ptr, col, val, 40.0, 0.05, len(lens), ni, 4)
    assert rel_err_rows(h.get_factors(0), X) < FACTOR_TOL
    assert abs(loss - loss_o) <= LOSS_TOL * abs(loss_o)
    assert np.all(h.get_factors(0)[4] == 0.0)  # empty row: A = G + lambda I, b = 0


def test_not_spd_is_reported():
    """lambda = 0 and fewer signals than factors with a zero Gram: singular system.  The reference
    aborts (CHECK_EQ(result, 0), qmf/Matrix.cpp:94); the C ABI returns QMFB_ERR_NOT_SPD."""
    from qmf_b200 import WalsEngineHandle, capi
    k, ni = 30, 50
    h = WalsEngineHandle(1, ni, k)
    h.set_csr(0, np.array([0, 2], np.int64), np.array([1, 3], np.int32), np.array([1.0, 2.0]))
    Y = np.zeros((ni, k))
    Y[1, 0] = 1.0
    Y[3, 1] = 1.0
    h.set_factors(1, Y)
    with pytest.raises(capi.QmfbError) as e:
        h.half_step(0, 40.0, 0.0)
    assert e.value.code == capi.ERR_NOT_SPD


def test_epoch_host_roundtrip(oracle_lib):
    h, ucsr, icsr, NU, NI = _setup(150, 100, 2500, 30, seed=77)
    Y0 = init_factors(NI, 30, seed=6)
    Xo, Yo = np.empty((NU, 30)), np.empty((NI, 30))
    loss = h.epoch_host(40.0, 0.05, Y0, Xo, Yo)
    X = np.zeros((NU, 30))
    Y = Y0.copy()
    oracle_lib.qmfo_wals_half_step(X, NU, Y, NI, 30, ucsr[0], ucsr[1], ucsr[2], 40.0, 0.05, NU, NI, 16)
    lo = oracle_lib.qmfo_wals_half_step(Y, NI, X, NU, 30, icsr[0], icsr[1], icsr[2], 40.0, 0.05, NU, NI, 16)
    assert rel_err_rows(Xo, X) < FACTOR_TOL and rel_err_rows(Yo, Y) < FACTOR_TOL
    assert abs(loss - lo) <= LOSS_TOL * abs(lo)


@pytest.mark.parametrize("k,npeers", [(30, 1), (128, 3), (128, 7), (160, 2)])
def test_fused_peer_stores_replicate_the_solved_rows(k, npeers):
    """qmfb_wals_solve_peers_dev: every solved row also lands in row (row_offset + r) of each peer
    replica (here: further buffers on the same GPU; across GPUs they are IPC-mapped peer memory)"""
    import torch
    from qmf_b200 import csr_from_coo
    from qmf_b200.wals_dist import CudaKernels
    K = CudaKernels()
    dev = torch.device("cuda", 0)
    u, i, v = uniform_dataset(3 * k + 40, 2 * k + 30, 30 * k, 5 + k, id_scale=(1, 1))
    uids, urp, uci, uv = csr_from_coo(u, i, v)
    NU, NI, kp = len(uids), int(uci.max()) + 1, K.padded_k(k)
    Y = torch.zeros(NI, kp, dtype=torch.float64, device=dev)
    Y[:, :k] = torch.from_numpy(init_factors(NI, k, seed=3)).to(dev)
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(dev).to(dt)
    rp, col, val = t(urp, torch.int64), t(uci, torch.int32), t(uv, torch.float64)
    order = torch.argsort(rp[1:] - rp[:-1], descending=True, stable=True).to(torch.int32)
    gram = torch.zeros(K.gram_packed_len(k), dtype=torch.float64, device=dev)
    ws = torch.empty(K.gram_workspace_len(k), dtype=torch.float64, device=dev)
    K.gram(Y, 0, NI, k, ws, gram)
    off = 11  # the shard's rows sit at [off, off + NU) of the replicas
    bufs = [torch.full((NU + 2 * off, kp), -7.0, dtype=torch.float64, device=dev) for _ in range(npeers + 1)]
    row_loss = torch.zeros(NU, dtype=torch.float64, device=dev)
    loss, scratch = torch.zeros(1, dtype=torch.float64, device=dev), torch.zeros(2, dtype=torch.int32, device=dev)
    K.solve(bufs[0], off, Y, k, rp, col, val, order, gram, 40.0, 0.05, row_loss, loss, scratch,
            peers=tuple(b.data_ptr() for b in bufs[1:]))
    torch.cuda.synchronize()
    assert int(scratch[1]) == 0
    ref = bufs[0].cpu().numpy()
    assert np.all(ref[:off] == -7.0) and np.all(ref[off + NU:] == -7.0) and np.all(ref[off:off + NU, k:] == 0.0)
    assert np.abs(ref[off:off + NU, :k]).max() > 0
    for b in bufs[1:]:
        assert np.array_equal(b.cpu().numpy(), ref)


def test_sharded_driver_single_rank_matches_oracle(oracle_lib):
    """the kernel-level ABI (device pointers, caller-owned torch tensors) used by the multi-GPU driver"""
    import torch
    from qmf_b200 import csr_from_coo
    from qmf_b200.wals_dist import ShardedWals
    u, i, v = uniform_dataset(220, 140, 4000, 31, id_scale=(5, 2), dup=12)
    uids, urp, uci, uv = csr_from_coo(u, i, v)
    iids, irp, ici, iv = csr_from_coo(i, u, v)
    NU, NI, k = len(uids), len(iids), 64
    dev = torch.device("cuda", 0)
    t = lambda a: tuple(torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in a)
    sw = ShardedWals(NU, NI, k, t((urp, uci, uv)), t((irp, ici, iv)), dev)
    Y0 = init_factors(NI, k, seed=8)
    sw.set_factors(1, Y0)
    X, Y = np.zeros((NU, k)), Y0.copy()
    for _ in range(2):
        loss = float(sw.epoch(40.0, 0.05))
        sw.check_error()
        oracle_lib.qmfo_wals_half_step(X, NU, Y, NI, k, urp, uci, uv, 40.0, 0.05, NU, NI, 16)
        lo = oracle_lib.qmfo_wals_half_step(Y, NI, X, NU, k, irp, ici, iv, 40.0, 0.05, NU, NI, 16)
        assert rel_err_rows(sw.get_factors(0).cpu().numpy(), X) < FACTOR_TOL
        assert rel_err_rows(sw.get_factors(1).cpu().numpy(), Y) < FACTOR_TOL
        assert abs(loss - lo) <= LOSS_TOL * abs(lo)


def test_full_size_properties_c3():
    """BASELINE config C3 (138k x 27k, 20M nnz, k=64) at full size: properties that do not need the
    oracle - (1) every sampled row satisfies its normal equations (G + sum alpha r y y^T + lambda I) x = b
    to 1e-10, (2) the kernel is deterministic (a repeated half-step is bit-identical), (3) the loss
    equals sum_rows [c + x^T(A - lambda I)x - 2 x^T b] evaluated on the host for the sampled rows."""
    import torch
    from qmf_b200.datagen import CONFIGS, init_item_factors, uniform_csr_torch
    from qmf_b200.wals_dist import ShardedWals
    nu, ni, nnz, k = CONFIGS["c3"]
    dev = torch.device("cuda", 0)
    csr_user, csr_item = uniform_csr_torch(nu, ni, nnz, seed=99, device=dev)
    sw = ShardedWals(nu, ni, k, csr_user, csr_item, dev)
    sw.set_factors(1, init_item_factors(ni, k, seed=1))
    alpha, lam = 40.0, 0.05
    sw.epoch(alpha, lam)
    sw.check_error()
    Y = sw.get_factors(1).clone()            # fixed side of the next user half-step
    sw.half_step(0, alpha, lam)
    row_loss = sw.row_loss[:nu].clone()
    X1 = sw.get_factors(0).clone()
    sw.half_step(0, alpha, lam)              # same inputs (Y unchanged) -> identical bits
    assert torch.equal(X1, sw.get_factors(0)) and torch.equal(row_loss, sw.row_loss[:nu])
    Yh = Y.cpu().numpy()
    G = Yh.T @ Yh
    rp, col, val = (t.cpu().numpy() for t in csr_user)
    rng = np.random.default_rng(0)
    rows = np.concatenate([rng.integers(0, nu, 40), [int(np.argmax(np.diff(rp))), int(np.argmin(np.diff(rp)))]])
    Xh, rl = X1.cpu().numpy(), row_loss.cpu().numpy()
    for r in rows:
        sl = slice(rp[r], rp[r + 1])
        Ys, w = Yh[col[sl]], val[sl]
        B = G + (Ys * (alpha * w)[:, None]).T @ Ys
        b = ((1 + alpha * w)[:, None] * Ys).sum(0)
        x = Xh[r]
        assert np.abs((B + lam * np.eye(k)) @ x - b).max() <= 1e-10 * np.abs(b).max(), r
        want = (1 + alpha * w).sum() + x @ B @ x - 2 * x @ b
        assert abs(rl[r] - want) <= 1e-11 * abs(want), (r, rl[r], want)


def test_extremely_long_row_is_built_by_many_ctas(oracle_lib):
    """one item with > kLongRow (32768) ratings: long_row_partial/reduce kernels + the solve kernel's
    no-gather path must give the oracle's factors and loss; the user side (no long row) is unaffected"""
    from qmf_b200 import WalsEngineHandle, csr_from_coo
    rng = np.random.default_rng(99)
    nu, ni, k = 60000, 200, 64   # nu, ni >= 2k: well-conditioned Gram matrices
    # item 7 is rated by 50 000 distinct users, everything else is sparse
    u_long = rng.choice(nu, size=50000, replace=False)
    u_rest = rng.integers(0, nu, size=120000)
    i_rest = rng.integers(0, ni, size=120000)
    u = np.concatenate([u_long, u_rest]).astype(np.int64)
    i = np.concatenate([np.full(50000, 7), i_rest]).astype(np.int64)
    cells = np.unique(u * ni + i)
    u, i = cells // ni, cells % ni
    v = rng.integers(1, 6, size=len(cells)).astype(np.float64)
    uids, urp, uci, uv = csr_from_coo(u, i, v)
    iids, irp, ici, iv = csr_from_coo(i, u, v)
    NU, NI = len(uids), len(iids)
    assert np.diff(irp).max() >= 50000
    h = WalsEngineHandle(NU, NI, k)
    h.set_csr(0, urp, uci, uv)
    h.set_csr(1, irp, ici, iv)
    Y0 = init_factors(NI, k, seed=5)
    h.set_factors(1, Y0)
    X, Y = np.zeros((NU, k)), Y0.copy()
    lu = h.half_step(0, 40.0, 0.05)
    li = h.half_step(1, 40.0, 0.05)
    lu_o = oracle_lib.qmfo_wals_half_step(X, NU, Y, NI, k, urp, uci, uv, 40.0, 0.05, NU, NI, 16)
    li_o = oracle_lib.qmfo_wals_half_step(Y, NI, X, NU, k, irp, ici, iv, 40.0, 0.05, NU, NI, 16)
    assert rel_err_rows(h.get_factors(0), X) < FACTOR_TOL and rel_err_rows(h.get_factors(1), Y) < FACTOR_TOL
    assert abs(lu - lu_o) <= LOSS_TOL * abs(lu_o) and abs(li - li_o) <= LOSS_TOL * abs(li_o)


_POWER_LAW = {}


def _power_law_problem():
    """Zipf-like item popularity: one item above kLongMaxParts * kLongSegMin entries (its segments are longer
    than the minimum), a dozen more above kLongRow with very different lengths, a long tail"""
    if "p" not in _POWER_LAW:
        from qmf_b200 import csr_from_coo
        rng = np.random.default_rng(2024)
        nu, ni = 1_200_000, 300
        lens = np.maximum((400_000 / np.arange(1, ni + 1)).astype(np.int64), 40)
        lens[0] = 1_100_000
        assert (lens >= 32768).sum() >= 12 and lens[0] > 256 * 4096
        u = np.concatenate([rng.choice(nu, size=int(n), replace=False) for n in lens]).astype(np.int64)
        i = np.repeat(np.arange(ni), lens).astype(np.int64)
        v = rng.integers(1, 6, size=len(u)).astype(np.float64)
        _POWER_LAW["p"] = (csr_from_coo(u, i, v), csr_from_coo(i, u, v))
    return _POWER_LAW["p"]


@pytest.mark.parametrize("k,mode", [(128, 1), (128, 2), (30, 0)])
def test_power_law_items_many_long_rows(oracle_lib, k, mode):
    """the (row, segment) work list of the long-row pre-pass, both solve kernels: the item half-step from the
    GPU's own user factors against the oracle's (factors per row, loss), long rows also against numpy"""
    from qmf_b200 import WalsEngineHandle, capi
    (uids, urp, uci, uv), (iids, irp, ici, iv) = _power_law_problem()
    NU, NI = len(uids), len(iids)
    alpha, lam = 40.0, 0.05
    try:
        capi.check(capi.lib.qmfb_wals_set_solve_kernel(mode))
        h = WalsEngineHandle(NU, NI, k)
        h.set_csr(0, urp, uci, uv)
        h.set_csr(1, irp, ici, iv)
        h.set_factors(1, init_factors(NI, k, seed=5))
        h.half_step(0, alpha, lam)
        X = h.get_factors(0)
        li = h.half_step(1, alpha, lam)
        Y = h.get_factors(1)
        h.close()
    finally:
        capi.check(capi.lib.qmfb_wals_set_solve_kernel(0))
    G = X.T @ X
    for r in list(range(14)) + [50, NI - 1]:
        sl = slice(irp[r], irp[r + 1])
        Xs, w = X[ici[sl]], iv[sl]
        B = G + (Xs * (alpha * w)[:, None]).T @ Xs
        b = ((1 + alpha * w)[:, None] * Xs).sum(0)
        want = np.linalg.solve(B + lam * np.eye(k), b)
        assert np.abs(Y[r] - want).max() <= 1e-9 * np.abs(want).max(), r
    key = ("oracle", k)
    if key not in _POWER_LAW:   # X is the same for both kernels (identical per-row arithmetic): one oracle run per k
        Yo = np.zeros((NI, k))
        _POWER_LAW[key] = (X.copy(), Yo, oracle_lib.qmfo_wals_half_step(Yo, NI, X, NU, k, irp, ici, iv, alpha, lam, NU, NI, 16))
    Xo, Yo, li_o = _POWER_LAW[key]
    assert np.array_equal(X, Xo)
    assert rel_err_rows(Y, Yo) < FACTOR_TOL
    # The loss of a row is c + x^T B x - 2 x^T b with c = sum_s (1 + alpha r_s) (WALSEngine.cpp:286,295-304): for a row of
    # 1.1e6 signals c ~ 1.3e8 and the three terms cancel to ~1e4, so the reference's own left-to-right sums carry an
    # absolute error of ~1e-16 * sqrt(n) * c that no other summation order reproduces.  The 1e-12 gate is therefore
    # applied relative to the magnitude the loss is summed FROM (sum of c over all rows), not to the cancelled result.
    c_total = float((1.0 + alpha * iv).sum()) / NU / NI
    assert abs(li - li_o) <= LOSS_TOL * max(abs(li_o), c_total), (li, li_o, c_total)


@pytest.mark.parametrize("nu,ni,nnz,k,dup", [(700, 500, 20000, 128, 0), (900, 300, 9000, 128, 11), (400, 300, 12000, 100, 0),
                                              (500, 400, 8000, 64, 0), (300, 200, 6000, 30, 5), (260, 210, 5000, 96, 0),
                                              (4000, 300, 60000, 128, 0)])
def test_warp_specialised_kernel_matches_the_plain_kernel_and_the_oracle(oracle_lib, nu, ni, nnz, k, dup):
    """wals_solve_ws_kernel (builder warps + two solver groups per SM) against the plain kernel: identical
    factors (same per-row arithmetic: gather order, DMMA order, Cholesky order) and loss; and against the
    oracle at the usual tolerances.  Forced through qmfb_wals_set_solve_kernel for every tile count."""
    from qmf_b200 import capi
    alpha, lam = 40.0, 0.05
    out = {}
    try:
        for mode in (1, 2):
            capi.check(capi.lib.qmfb_wals_set_solve_kernel(mode))
            h, ucsr, icsr, NU, NI = _setup(nu, ni, nnz, k, seed=3 * k + nu, dup=dup)
            h.set_factors(1, init_factors(NI, k, seed=5))
            res = []
            for _ in range(2):
                lu = h.half_step(0, alpha, lam)
                li = h.half_step(1, alpha, lam)
                res.append((lu, li, h.get_factors(0), h.get_factors(1)))
            out[mode] = res
            h.close()
    finally:
        capi.check(capi.lib.qmfb_wals_set_solve_kernel(0))
    X, Y = np.zeros((NU, k)), init_factors(NI, k, seed=5)
    for e in range(2):
        assert out[1][e][0] == out[2][e][0] and out[1][e][1] == out[2][e][1]
        assert np.array_equal(out[1][e][2], out[2][e][2]) and np.array_equal(out[1][e][3], out[2][e][3])
        lu_o = oracle_lib.qmfo_wals_half_step(X, NU, Y, NI, k, ucsr[0], ucsr[1], ucsr[2], alpha, lam, NU, NI, 16)
        li_o = oracle_lib.qmfo_wals_half_step(Y, NI, X, NU, k, icsr[0], icsr[1], icsr[2], alpha, lam, NU, NI, 16)
        assert rel_err_rows(out[2][e][2], X) < FACTOR_TOL and rel_err_rows(out[2][e][3], Y) < FACTOR_TOL
        assert abs(out[2][e][0] - lu_o) <= LOSS_TOL * abs(lu_o) and abs(out[2][e][1] - li_o) <= LOSS_TOL * abs(li_o)
