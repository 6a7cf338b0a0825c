"""Python driver over the BPR part of the C ABI (host buffers; what the C++ qmf::BPREngine binds)."""
import ctypes as C

import numpy as np

from .capi import SIDE_ITEM, SIDE_USER, check, lib


class BprEngineHandle:
    def __init__(self, nusers, nitems, nfactors, use_biases=False, device=0):
        self.nusers, self.nitems, self.k, self.use_biases = int(nusers), int(nitems), int(nfactors), bool(use_biases)
        self._h = None
        h = C.c_void_p()
        check(lib.qmfb_bpr_create(device, self.nusers, self.nitems, self.k, int(self.use_biases), C.byref(h)))
        self._h = h

    def close(self):
        if self._h:
            lib.qmfb_bpr_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def _n(self, side):
        return self.nusers if side == SIDE_USER else self.nitems

    def set_data(self, user_idx, item_idx):
        u = np.ascontiguousarray(user_idx, dtype=np.int32)
        i = np.ascontiguousarray(item_idx, dtype=np.int32)
        check(lib.qmfb_bpr_set_data(self._h, u, i, len(u)))

    def set_factors(self, side, F):
        F = np.ascontiguousarray(F, dtype=np.float64)
        assert F.shape == (self._n(side), self.k)
        check(lib.qmfb_bpr_set_factors(self._h, side, F))

    def get_factors(self, side):
        F = np.empty((self._n(side), self.k), dtype=np.float64)
        check(lib.qmfb_bpr_get_factors(self._h, side, F))
        return F

    def set_biases(self, b):
        check(lib.qmfb_bpr_set_biases(self._h, np.ascontiguousarray(b, dtype=np.float64)))

    def get_biases(self):
        b = np.empty(self.nitems, dtype=np.float64)
        check(lib.qmfb_bpr_get_biases(self._h, b))
        return b

    def epoch(self, lr, user_lambda, item_lambda, bias_lambda, num_neg, seed, epoch, shuffle=True):
        n = C.c_int64()
        check(lib.qmfb_bpr_epoch(self._h, lr, user_lambda, item_lambda, bias_lambda, num_neg, seed, epoch, int(shuffle),
                                 C.byref(n)))
        return n.value

    def update_triplets(self, u, i, j, lr, user_lambda, item_lambda, bias_lambda):
        u, i, j = (np.ascontiguousarray(a, dtype=np.int32) for a in (u, i, j))
        check(lib.qmfb_bpr_update_triplets(self._h, u, i, j, len(u), lr, user_lambda, item_lambda, bias_lambda))

    def eval_loss_sum(self, u, i, j, n=None):
        u, i, j = (np.ascontiguousarray(a, dtype=np.int32) for a in (u, i, j))
        s = C.c_double()
        check(lib.qmfb_bpr_eval_loss(self._h, u, i, j, len(u) if n is None else n, C.byref(s)))
        return s.value

    def eval_loss(self, u, i, j, nthreads=1):
        """mean loss with the reference's block/tail-drop summation (ParallelExecutor-inl.h:60-85):
        only nthreads * floor(n / nthreads) triplets are summed, the mean divides by n"""
        n = len(u)
        if n == 0:
            return -1.0
        used = (n // nthreads) * nthreads
        return self.eval_loss_sum(u, i, j, used) / n

    def eval_rank(self, test_users, label_ptr, label_items):
        """ranking statistics of the test users on the RESIDENT factors and biases (qmfb_bpr_eval_rank)"""
        from .wals import _eval_rank_call
        return _eval_rank_call(lib.qmfb_bpr_eval_rank, self._h, test_users, label_ptr, label_items)

    def set_concurrency(self, max_pairs_in_flight):
        check(lib.qmfb_bpr_set_concurrency(self._h, int(max_pairs_in_flight)))

    def set_hogwild_blocks(self, n):
        check(lib.qmfb_bpr_set_hogwild_blocks(self._h, int(n)))

    def last_epoch_ms(self):
        ms = C.c_float()
        check(lib.qmfb_bpr_last_epoch_ms(self._h, C.byref(ms)))
        return ms.value

    def launch_count(self):
        return int(lib.qmfb_bpr_launch_count(self._h))
