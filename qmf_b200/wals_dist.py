"""One-process-per-GPU WALS driver over the kernel-level C ABI (qmfb_gram_dev / qmfb_wals_solve_dev).

Sharding: users and items are each split into `world` contiguous row ranges
balanced by nnz.  Every rank keeps full replicas of both factor matrices plus the CSR rows of
its user range and the CSC rows of its item range.  One half-step "update side S" is

  1. partial Gram over THIS rank's rows of the other side   (qmfb_gram_dev)
  2. allreduce(sum) of the packed Gram (<= 68 KB)            (NCCL)
  3. solve this rank's rows of S                             (qmfb_wals_solve_dev)
  4. every solved row also goes into the other ranks' replicas, FROM the solve kernel, as
     peer-memory stores over NVLink (qmfb_wals_solve_peers_dev; replicas are CUDA-IPC buffers) -
     there is no separate all-gather.  (exchange="nccl": one broadcast per rank instead.)
  5. allreduce(sum) of the loss scalar                       (NCCL) - which also orders the next
     half-step's reads after every rank's peer stores

torch is used for device memory, streams and torch.distributed only.  With world == 1 no
collective is issued.  The same class runs on CPU tensors with the `gloo` backend when a
`kernels` object is injected (tests/test_host_logic_cpu.py::test_sharded_wals_two_ranks_gloo exercises the sharding/exchange logic
that way); the product path always uses the CUDA library.
"""
import ctypes as C

import numpy as np
import torch

try:
    import torch.distributed as dist
except Exception:  # pragma: no cover
    dist = None


def balanced_row_ranges(row_ptr, world):
    """Split rows into `world` contiguous ranges with ~equal nnz (prefix sum over row_ptr).
    Returns a list of (begin, end)."""
    rp = row_ptr.detach().cpu().numpy() if isinstance(row_ptr, torch.Tensor) else np.asarray(row_ptr)
    nrows = len(rp) - 1
    nnz = int(rp[-1])
    cuts = [0]
    for r in range(1, world):
        target = nnz * r // world
        c = int(np.searchsorted(rp, target, side="left"))
        c = min(max(c, cuts[-1]), nrows)
        cuts.append(c)
    cuts.append(nrows)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


class CudaKernels:
    """The product path: hand-written sm_100a kernels behind the C ABI."""

    def __init__(self):
        from . import capi
        self.capi = capi
        self.lib = capi.lib

    def padded_k(self, k):
        return self.capi.check(self.lib.qmfb_padded_k(k))

    def gram_packed_len(self, k):
        return int(self.lib.qmfb_gram_packed_len(k))

    def gram_workspace_len(self, k):
        return int(self.lib.qmfb_gram_workspace_len(k))

    @staticmethod
    def _stream():
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def gram(self, Y, row_begin, row_end, k, ws, out):
        self.capi.check(self.lib.qmfb_gram_dev(self._stream(), Y.data_ptr(), Y.stride(0), row_begin, row_end, k,
                                               ws.data_ptr(), out.data_ptr()))

    def solve(self, X, row_offset, Y, k, row_ptr, col, val, order, gram, alpha, lam, row_loss, loss_sum, scratch,
              peers=(), nnz=-1):
        """peers: raw device pointers of the other ranks' replicas of X (fused all-gather)"""
        arr = (C.c_void_p * max(len(peers), 1))(*peers)
        self.capi.check(self.lib.qmfb_wals_solve_peers_dev(
            self._stream(), X.data_ptr(), X.stride(0), row_offset, Y.data_ptr(), Y.stride(0), k, row_ptr.data_ptr(),
            col.data_ptr(), val.data_ptr(), order.data_ptr(), order.numel(), nnz, gram.data_ptr(), alpha, lam,
            row_loss.data_ptr(), loss_sum.data_ptr(), scratch.data_ptr(), arr, len(peers)))

    # ---- replicas shareable between the ranks of one box (CUDA IPC) ----
    def ipc_alloc(self, device_index, nbytes):
        ptr, handle = C.c_void_p(), C.create_string_buffer(64)
        self.capi.check(self.lib.qmfb_ipc_alloc(device_index, nbytes, C.byref(ptr), handle))
        return ptr.value, handle.raw

    def ipc_open(self, device_index, handle):
        ptr = C.c_void_p()
        self.capi.check(self.lib.qmfb_ipc_open(device_index, handle, C.byref(ptr)))
        return ptr.value

    def ipc_close(self, ptr):
        self.lib.qmfb_ipc_close(C.c_void_p(ptr))

    def ipc_free(self, ptr):
        self.lib.qmfb_ipc_free(C.c_void_p(ptr))

    launches_per_half_step = 7  # gram_partial, gram_reduce, long_row_plan, long_row_partial, long_row_reduce, wals_solve, sum


class _DeviceBuffer:
    """Zero-copy torch view of a raw device allocation (CUDA array interface)."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f8", "data": (int(ptr), False), "version": 2,
                                         "strides": None}


class ShardedWals:
    """State of one rank.  `csr[side]` = (row_ptr int64 [n+1], col int32 [nnz], val f64 [nnz]) of
    the FULL problem for that orientation, on `device`; the rank keeps only its slice."""

    def __init__(self, nusers, nitems, k, csr_user, csr_item, device, rank=0, world=1, kernels=None, exchange="auto",
                 ranges=None):
        """ranges=None: csr_* hold the FULL problem and are cut here into nnz-balanced contiguous row ranges.
        ranges=(user_ranges, item_ranges), each a list of `world` (begin, end): csr_* are ALREADY this rank's
        rows only (row_ptr local, starting at 0) - the form for problems no single GPU should materialise
        (C5: qmf_b200.datagen.powerlaw_shard_torch)."""
        self.n = (int(nusers), int(nitems))
        self.k = int(k)
        self.rank, self.world = rank, world
        self.device = device
        self.kern = kernels if kernels is not None else CudaKernels()
        self.kp = self.kern.padded_k(self.k)
        # exchange of the solved shards: "p2p" = peer stores from the solve kernel into IPC-mapped replicas
        # (CUDA product path, world > 1), "nccl" = one broadcast per rank after the kernel
        auto = exchange == "auto"
        if auto:
            exchange = "p2p" if (world > 1 and kernels is None and torch.device(device).type == "cuda") else "nccl"
        self.exchange = exchange
        self._ipc_own, self._ipc_peers, self.peers = [], [], [(), ()]
        if exchange == "p2p" and not self._init_p2p(torch.device(device), required=not auto):
            self.exchange = exchange = "nccl"   # auto: CUDA IPC is not available between these processes
        if exchange != "p2p":
            self.F = [torch.zeros(self.n[s], self.kp, dtype=torch.float64, device=device) for s in (0, 1)]
        self.ranges, self.shard = [], []
        for side, (rp, col, val) in enumerate((csr_user, csr_item)):
            if ranges is not None:
                rr = [(int(b), int(e)) for b, e in ranges[side]]
                b, e = rr[rank]
                assert rp.numel() == e - b + 1, "local row_ptr does not match this rank's range"
                lrp, lcol, lval, p0, p1 = rp.contiguous(), col.contiguous(), val.contiguous(), 0, int(rp[-1])
                if p1 == 0:
                    lcol = torch.zeros(1, dtype=torch.int32, device=device)
                    lval = torch.zeros(1, dtype=torch.float64, device=device)
            else:
                rr = balanced_row_ranges(rp, world)
                b, e = rr[rank]
                p0, p1 = int(rp[b]), int(rp[e])
                lrp = (rp[b:e + 1] - rp[b]).contiguous()
                lcol = col[p0:p1].contiguous() if p1 > p0 else torch.zeros(1, dtype=torch.int32, device=device)
                lval = val[p0:p1].contiguous() if p1 > p0 else torch.zeros(1, dtype=torch.float64, device=device)
            lens = lrp[1:] - lrp[:-1]
            order = torch.argsort(lens, descending=True, stable=True).to(torch.int32).contiguous()
            self.ranges.append(rr)
            self.shard.append(dict(begin=b, end=e, row_ptr=lrp, col=lcol, val=lval, order=order, nnz=p1 - p0))
        self.gram_packed = torch.zeros(self.kern.gram_packed_len(self.k), dtype=torch.float64, device=device)
        self.gram_ws = torch.empty(self.kern.gram_workspace_len(self.k), dtype=torch.float64, device=device)
        self.row_loss = torch.zeros(max(max(s["end"] - s["begin"] for s in self.shard), 1), dtype=torch.float64,
                                    device=device)
        # [0] = loss sum of the last half-step (written by the kernel), [1] = number of half-steps, on ANY rank and
        # since the last check_error(), in which a pivot was not positive: both travel in the one allreduce
        # that follows a half-step, so every rank sees the same flag and raises together
        self.loss_err = torch.zeros(2, dtype=torch.float64, device=device)
        self.loss_sum = self.loss_err[:1]
        self.scratch = torch.zeros(2, dtype=torch.int32, device=device)
        self.launches = 0
        self.timing = None  # optional dict of torch.cuda.Event pairs filled by half_step(record=True)

    def _init_p2p(self, device, required=True):
        """Factor replicas as CUDA-IPC allocations; every rank maps every other rank's replicas.
        Returns False (after releasing everything, on EVERY rank) if some rank could not export or map
        a replica and `required` is False - e.g. processes in different IPC namespaces."""
        idx = device.index if device.index is not None else torch.cuda.current_device()
        self.F, handles, err = [], [], None
        try:
            import os
            if os.environ.get("QMFB_FORCE_NO_IPC"):  # test hook for the fallback below
                raise RuntimeError("CUDA IPC disabled by QMFB_FORCE_NO_IPC")
            for s in (0, 1):
                ptr, handle = self.kern.ipc_alloc(idx, self.n[s] * self.kp * 8)
                self._ipc_own.append(ptr)
                handles.append(handle)
        except Exception as e:  # noqa: BLE001 - reported below, collectively
            err, handles = e, None
        gathered = [None] * self.world
        dist.all_gather_object(gathered, handles)
        peers = [[], []]
        if err is None and all(g is not None for g in gathered):
            try:
                for r in range(self.world):
                    if r == self.rank:
                        continue
                    for s in (0, 1):
                        p = self.kern.ipc_open(idx, gathered[r][s])
                        self._ipc_peers.append(p)
                        peers[s].append(p)
            except Exception as e:  # noqa: BLE001
                err = e
        elif err is None:
            err = RuntimeError("another rank could not allocate its shareable replicas")
        ok = torch.tensor([0 if err is not None else 1], dtype=torch.int32, device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            for p in self._ipc_peers:
                self.kern.ipc_close(p)
            dist.barrier()   # nobody frees a replica another rank still has mapped
            for p in self._ipc_own:
                self.kern.ipc_free(p)
            self._ipc_own, self._ipc_peers, self.F = [], [], []
            if required:
                raise RuntimeError("p2p exchange unavailable: %s" % (err if err is not None else "failed on another rank"))
            return False
        self.F = [torch.as_tensor(_DeviceBuffer(self._ipc_own[s], (self.n[s], self.kp)), device=device) for s in (0, 1)]
        self.peers = [tuple(peers[0]), tuple(peers[1])]
        return True

    def close(self):
        """Unmap the peers' replicas and free our own (p2p exchange); collective: call on every rank."""
        if self._ipc_own:
            torch.cuda.synchronize()
            if self.world > 1:
                dist.barrier()
            for p in self._ipc_peers:
                self.kern.ipc_close(p)
            self.F = []
            for p in self._ipc_own:
                self.kern.ipc_free(p)
            self._ipc_own, self._ipc_peers, self.peers = [], [], [(), ()]

    def set_factors(self, side, F):
        """F: [n, k] tensor/ndarray (host or device)"""
        F = torch.as_tensor(F, dtype=torch.float64)
        self.F[side][:, :self.k].copy_(F, non_blocking=True)

    def get_factors(self, side):
        return self.F[side][:, :self.k]

    def half_step(self, side, alpha, lam, events=None):
        """Returns the device scalar holding the summed loss of ALL rows of `side` (all ranks)."""
        other = 1 - side
        sh, osh = self.shard[side], self.shard[other]
        if events is not None:
            events["gram0"].record()
        self.kern.gram(self.F[other], osh["begin"], osh["end"], self.k, self.gram_ws, self.gram_packed)
        if self.world > 1:
            dist.all_reduce(self.gram_packed)
        if events is not None:
            events["solve0"].record()
        # leftData.setFactors(0) (WALSEngine.cpp:170-171) — rows of this shard; the others arrive in step 4
        self.F[side][sh["begin"]:sh["end"]].zero_()
        if self.exchange == "p2p":
            self.kern.solve(self.F[side], sh["begin"], self.F[other], self.k, sh["row_ptr"], sh["col"], sh["val"],
                            sh["order"], self.gram_packed, alpha, lam, self.row_loss, self.loss_sum, self.scratch,
                            peers=self.peers[side], nnz=sh["nnz"])
        else:
            self.kern.solve(self.F[side], sh["begin"], self.F[other], self.k, sh["row_ptr"], sh["col"], sh["val"],
                            sh["order"], self.gram_packed, alpha, lam, self.row_loss, self.loss_sum, self.scratch,
                            nnz=sh["nnz"])
        if events is not None:
            events["solve1"].record()
        self.launches += self.kern.launches_per_half_step
        # the launcher clears scratch (the NOT_SPD flag) at the start of every solve: fold it into the sticky count
        self.loss_err[1:2] += self.scratch[1:2]
        if self.world > 1:
            if self.exchange != "p2p":
                for r, (b, e) in enumerate(self.ranges[side]):
                    if e > b:
                        dist.broadcast(self.F[side][b:e], src=r)
            # also the cross-rank ordering point of the p2p exchange: it completes on a rank only after every
            # rank's solve kernel (and with it that rank's peer stores) has completed
            dist.all_reduce(self.loss_err)
        return self.loss_sum

    def epoch(self, alpha, lam, events=None):
        """user half-step, then item half-step; returns the item-step loss / nusers / nitems as a
        device tensor (WALSEngine::optimize, WALSEngine.cpp:86-92)"""
        self.half_step(0, alpha, lam, events[0] if events else None)
        loss = self.half_step(1, alpha, lam, events[1] if events else None)
        return loss / self.n[0] / self.n[1]

    def epoch_host(self, alpha, lam, item_in, user_out, item_out, loss_out=None):
        """One epoch with (pinned) host buffers: the item factors come from `item_in` [nitems, k], this
        rank's solved user / item rows go to `user_out` / `item_out`.  The copy of the user rows runs on
        a side stream underneath the item half-step (which only reads them).  Asynchronous."""
        k = self.k
        ub, ue = self.shard[0]["begin"], self.shard[0]["end"]
        ib, ie = self.shard[1]["begin"], self.shard[1]["end"]
        main = torch.cuda.current_stream()
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream()
        self.F[1][:, :k].copy_(item_in, non_blocking=True)
        self.half_step(0, alpha, lam)
        self._copy_stream.wait_stream(main)
        with torch.cuda.stream(self._copy_stream):
            user_out.copy_(self.F[0][ub:ue, :k], non_blocking=True)
        loss = self.half_step(1, alpha, lam)
        item_out.copy_(self.F[1][ib:ie, :k], non_blocking=True)
        loss = loss / self.n[0] / self.n[1]
        if loss_out is not None:
            loss_out.copy_(loss, non_blocking=True)
        main.wait_stream(self._copy_stream)
        return loss

    def check_error(self):
        """Raises on EVERY rank if any rank hit a non-positive pivot in any half-step since the last call."""
        bad = float(self.loss_err[1].item()) != 0.0
        self.loss_err[1:2].zero_()
        if bad:
            raise RuntimeError("normal equations not positive definite (reference: dsysv failed, qmf/Matrix.cpp:94)")
