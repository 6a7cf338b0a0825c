"""Python drivers over the WALS part of the C ABI.

``WalsEngineHandle`` wraps the host-buffer engine level (what the C++ ``qmf::WALSEngine`` binds);
``csr_from_coo`` builds the CSR exactly as WALSEngine::groupSignals does
(qmf/wals/WALSEngine.cpp:130-163): rows and entries in ascending raw-id order, duplicates kept.
"""
import ctypes as C

import numpy as np

from . import capi
from .capi import SIDE_ITEM, SIDE_USER, check, lib


def csr_from_coo(row_id, col_id, val, col_ids_sorted=None):
    """(row ids sorted unique, row_ptr int64, col_idx int32, val f64) with rows = ascending raw
    row id, entries within a row ascending raw col id (stable for duplicates), col_idx = rank of
    the raw col id among ``col_ids_sorted`` (default: the distinct col ids of the input)."""
    row_id = np.asarray(row_id, dtype=np.int64)
    col_id = np.asarray(col_id, dtype=np.int64)
    val = np.asarray(val, dtype=np.float64)
    perm = np.lexsort((col_id, row_id))
    r, c, v = row_id[perm], col_id[perm], val[perm]
    ids, counts = np.unique(r, return_counts=True)
    row_ptr = np.zeros(len(ids) + 1, dtype=np.int64)
    np.cumsum(counts, out=row_ptr[1:])
    if col_ids_sorted is None:
        col_ids_sorted = np.unique(col_id)
    col_idx = np.searchsorted(col_ids_sorted, c).astype(np.int32)
    return ids, row_ptr, col_idx, np.ascontiguousarray(v)


def _eval_rank_call(fn, handle, test_users, label_ptr, label_items):
    """shared by the engines' eval_rank methods: (cnt, pos_scores), layout of qmfb_eval_rank"""
    tu = np.ascontiguousarray(test_users, dtype=np.int32)
    lp = np.ascontiguousarray(label_ptr, dtype=np.int64)
    li = np.ascontiguousarray(label_items, dtype=np.int32)
    nT, nl = len(tu), int(lp[-1])
    cnt = np.zeros(nl + nT, dtype=np.int32)
    sc = np.zeros(max(nl, 1), dtype=np.float64)
    if li.size == 0:
        li = np.zeros(1, np.int32)
    check(fn(handle, tu, nT, lp, li, cnt, sc))
    return cnt, sc[:nl]


class Signals:
    """GPU-built dense indexing + both CSR orientations of a raw (user id, item id, value) dataset
    (qmfb_signals_*: IdIndex + WALSEngine::groupSignals on the device)."""

    def __init__(self, user_ids, item_ids, values, device=0):
        u = np.ascontiguousarray(user_ids, dtype=np.int64)
        i = np.ascontiguousarray(item_ids, dtype=np.int64)
        v = np.ascontiguousarray(values, dtype=np.float64)
        assert u.shape == i.shape == v.shape and u.ndim == 1
        self._h = None
        h = C.c_void_p()
        check(lib.qmfb_signals_build(device, u.size, u, i, v, C.byref(h)))
        self._h = h
        nu, ni, nnz = C.c_int64(), C.c_int64(), C.c_int64()
        check(lib.qmfb_signals_dims(h, C.byref(nu), C.byref(ni), C.byref(nnz)))
        self.nusers, self.nitems, self.nnz = nu.value, ni.value, nnz.value

    def close(self):
        if self._h:
            lib.qmfb_signals_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def ids(self, side):
        out = np.empty(self.nusers if side == SIDE_USER else self.nitems, dtype=np.int64)
        check(lib.qmfb_signals_ids(self._h, side, out))
        return out

    def csr(self, side):
        """(row_ptr, col_idx, val, order) host copies of one orientation."""
        n = self.nusers if side == SIDE_USER else self.nitems
        rp, col = np.empty(n + 1, np.int64), np.empty(self.nnz, np.int32)
        val, order = np.empty(self.nnz, np.float64), np.empty(n, np.int32)
        check(lib.qmfb_signals_csr(self._h, side, rp.ctypes.data, col.ctypes.data, val.ctypes.data, order.ctypes.data))
        return rp, col, val, order


class WalsEngineHandle:
    """Engine-level handle: host buffers in, host buffers out (single GPU)."""

    def __init__(self, nusers, nitems, nfactors, device=0):
        self.nusers, self.nitems, self.k = int(nusers), int(nitems), int(nfactors)
        self._h = None
        h = C.c_void_p()
        check(lib.qmfb_wals_create(device, self.nusers, self.nitems, self.k, C.byref(h)))
        self._h = h

    def close(self):
        if self._h:
            lib.qmfb_wals_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def _n(self, side):
        return self.nusers if side == SIDE_USER else self.nitems

    def set_signals(self, signals):
        check(lib.qmfb_wals_set_signals(self._h, signals._h))

    def set_csr(self, side, row_ptr, col_idx, val, row_begin=0):
        row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int64)
        col_idx = np.ascontiguousarray(col_idx, dtype=np.int32)
        val = np.ascontiguousarray(val, dtype=np.float64)
        if col_idx.size == 0:
            col_idx, val = np.zeros(1, np.int32), np.zeros(1, np.float64)
        check(lib.qmfb_wals_set_csr(self._h, side, row_begin, len(row_ptr) - 1, row_ptr, col_idx, val))

    def set_factors(self, side, F):
        F = np.ascontiguousarray(F, dtype=np.float64)
        assert F.shape == (self._n(side), self.k)
        check(lib.qmfb_wals_set_factors(self._h, side, F))

    def get_factors(self, side):
        F = np.empty((self._n(side), self.k), dtype=np.float64)
        check(lib.qmfb_wals_get_factors(self._h, side, F))
        return F

    def gram(self, side):
        G = np.empty((self.k, self.k), dtype=np.float64)
        check(lib.qmfb_wals_gram(self._h, side, G))
        return G

    def half_step(self, side, alpha, lam):
        """returns the summed row losses divided by nusers*nitems (WALSEngine.cpp:215)"""
        loss = C.c_double()
        check(lib.qmfb_wals_half_step(self._h, side, alpha, lam, C.byref(loss)))
        return loss.value / self.nusers / self.nitems

    def epoch_host(self, alpha, lam, item_in=None, user_out=None, item_out=None):
        loss = C.c_double()
        pin = item_in.ctypes.data_as(C.c_void_p) if item_in is not None else None
        pu = user_out.ctypes.data_as(C.c_void_p) if user_out is not None else None
        pi = item_out.ctypes.data_as(C.c_void_p) if item_out is not None else None
        check(lib.qmfb_wals_epoch_host(self._h, alpha, lam, pin, pu, pi, C.byref(loss)))
        return loss.value

    def eval_rank(self, test_users, label_ptr, label_items):
        """ranking statistics of the test users on the RESIDENT factors (qmfb_wals_eval_rank)"""
        return _eval_rank_call(lib.qmfb_wals_eval_rank, self._h, test_users, label_ptr, label_items)

    def launch_count(self):
        return int(lib.qmfb_wals_launch_count(self._h))

    def last_timing(self):
        g, s = C.c_float(), C.c_float()
        check(lib.qmfb_wals_last_timing(self._h, C.byref(g), C.byref(s)))
        return g.value, s.value


class ShardedWalsHandle:
    """WALS across several GPUs of one box from ONE process (qmfb_wals_sharded_*: what `wals --ngpus N`
    binds).  `devices` lists CUDA ordinals; a device may repeat (its shards then share that GPU).  Factors
    and loss are bit-identical to WalsEngineHandle for every device count."""

    def __init__(self, nusers, nitems, nfactors, devices):
        self.nusers, self.nitems, self.k = int(nusers), int(nitems), int(nfactors)
        self.devices = [int(d) for d in devices]
        self._h = None
        h = C.c_void_p()
        arr = (C.c_int * len(self.devices))(*self.devices)
        check(lib.qmfb_wals_sharded_create(len(self.devices), arr, self.nusers, self.nitems, self.k, C.byref(h)))
        self._h = h

    def close(self):
        if self._h:
            lib.qmfb_wals_sharded_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def _n(self, side):
        return self.nusers if side == SIDE_USER else self.nitems

    def set_csr(self, side, row_ptr, col_idx, val):
        row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int64)
        col_idx = np.ascontiguousarray(col_idx, dtype=np.int32)
        val = np.ascontiguousarray(val, dtype=np.float64)
        assert len(row_ptr) == self._n(side) + 1
        if col_idx.size == 0:
            col_idx, val = np.zeros(1, np.int32), np.zeros(1, np.float64)
        check(lib.qmfb_wals_sharded_set_csr(self._h, side, row_ptr, col_idx, val))

    def set_signals(self, signals):
        check(lib.qmfb_wals_sharded_set_signals(self._h, signals._h))

    def shard(self, slot, side):
        dev, b, n, nnz = C.c_int(), C.c_int64(), C.c_int64(), C.c_int64()
        check(lib.qmfb_wals_sharded_shard(self._h, slot, side, C.byref(dev), C.byref(b), C.byref(n), C.byref(nnz)))
        return dev.value, b.value, n.value, nnz.value

    def set_factors(self, side, F):
        F = np.ascontiguousarray(F, dtype=np.float64)
        assert F.shape == (self._n(side), self.k)
        check(lib.qmfb_wals_sharded_set_factors(self._h, side, F))

    def get_factors(self, side, slot=0):
        F = np.empty((self._n(side), self.k), dtype=np.float64)
        check(lib.qmfb_wals_sharded_get_factors(self._h, side, slot, F))
        return F

    def half_step(self, side, alpha, lam):
        loss = C.c_double()
        check(lib.qmfb_wals_sharded_half_step(self._h, side, alpha, lam, C.byref(loss)))
        return loss.value / self.nusers / self.nitems

    def epoch_host(self, alpha, lam, item_in=None, user_out=None, item_out=None):
        loss = C.c_double()
        pin = item_in.ctypes.data_as(C.c_void_p) if item_in is not None else None
        pu = user_out.ctypes.data_as(C.c_void_p) if user_out is not None else None
        pi = item_out.ctypes.data_as(C.c_void_p) if item_out is not None else None
        check(lib.qmfb_wals_sharded_epoch_host(self._h, alpha, lam, pin, pu, pi, C.byref(loss)))
        return loss.value

    def eval_rank(self, test_users, label_ptr, label_items):
        """the test users cut over the GPUs, each slice against that GPU's replicas (qmfb_wals_sharded_eval_rank)"""
        return _eval_rank_call(lib.qmfb_wals_sharded_eval_rank, self._h, test_users, label_ptr, label_items)

    def launch_count(self):
        return int(lib.qmfb_wals_sharded_launch_count(self._h))

    def last_timing(self):
        g, s = C.c_float(), C.c_float()
        check(lib.qmfb_wals_sharded_last_timing(self._h, C.byref(g), C.byref(s)))
        return g.value, s.value
