"""ctypes binding of the C ABI in include/qmf_b200.h (libqmf_b200.so, built in-tree by
qmf_b200/csrc/Makefile).  There is NO CPU fallback: if the library is missing, import fails
loudly; if a call fails, :class:`QmfbError` carries qmfb_last_error()."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QMFB_LIB", os.path.join(_HERE, "libqmf_b200.so"))  # QMFB_LIB: debug builds only

SIDE_USER = 0
SIDE_ITEM = 1
ERR_NOT_SPD = -3
ERR_NOT_FINITE = -4


class QmfbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("qmf_b200 error %d: %s" % (code, msg))
        self.code = code


if not os.path.exists(LIB_PATH):
    raise ImportError(
        "qmf_b200: %s is missing — build it with `make -C qmf_b200/csrc` (or __graft_entry__.build()); "
        "there is no CPU fallback" % LIB_PATH)

lib = C.CDLL(LIB_PATH)

c_i64 = C.c_int64
c_f64 = C.c_double
vp = C.c_void_p
p_i64 = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
p_i32 = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
p_f64 = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")

_SIG = {
    "qmfb_last_error": (C.c_char_p, []),
    "qmfb_version": (C.c_int, []),
    "qmfb_device_count": (C.c_int, []),
    "qmfb_padded_k": (C.c_int, [C.c_int]),
    "qmfb_gram_packed_len": (c_i64, [C.c_int]),
    "qmfb_gram_workspace_len": (c_i64, [C.c_int]),
    "qmfb_gram_dev": (C.c_int, [vp, vp, c_i64, c_i64, c_i64, C.c_int, vp, vp]),
    "qmfb_gram_parts_count": (C.c_int, [c_i64, C.c_int]),
    "qmfb_gram_parts_dev": (C.c_int, [vp, vp, c_i64, c_i64, c_i64, C.c_int, C.c_int, C.c_int, vp]),
    "qmfb_gram_reduce_parts_dev": (C.c_int, [vp, C.POINTER(vp), C.POINTER(C.c_int), C.c_int, C.c_int, vp]),
    "qmfb_gram_unpack_dev": (C.c_int, [vp, vp, C.c_int, vp]),
    "qmfb_wals_solve_dev": (C.c_int, [vp, vp, c_i64, c_i64, vp, c_i64, C.c_int, vp, vp, vp, vp, c_i64, c_i64, vp, c_f64, c_f64,
                                      vp, vp, vp]),
    "qmfb_wals_solve_peers_dev": (C.c_int, [vp, vp, c_i64, c_i64, vp, c_i64, C.c_int, vp, vp, vp, vp, c_i64, c_i64, vp, c_f64,
                                            c_f64, vp, vp, vp, C.POINTER(vp), C.c_int]),
    "qmfb_wals_set_solve_kernel": (C.c_int, [C.c_int]),
    "qmfb_ipc_alloc": (C.c_int, [C.c_int, c_i64, C.POINTER(vp), C.c_char_p]),
    "qmfb_ipc_open": (C.c_int, [C.c_int, C.c_char_p, C.POINTER(vp)]),
    "qmfb_ipc_close": (C.c_int, [vp]),
    "qmfb_ipc_free": (C.c_int, [vp]),
    "qmfb_wals_create": (C.c_int, [C.c_int, c_i64, c_i64, C.c_int, C.POINTER(vp)]),
    "qmfb_wals_destroy": (C.c_int, [vp]),
    "qmfb_wals_set_csr": (C.c_int, [vp, C.c_int, c_i64, c_i64, p_i64, p_i32, p_f64]),
    "qmfb_wals_set_factors": (C.c_int, [vp, C.c_int, p_f64]),
    "qmfb_wals_get_factors": (C.c_int, [vp, C.c_int, p_f64]),
    "qmfb_wals_gram": (C.c_int, [vp, C.c_int, p_f64]),
    "qmfb_wals_half_step": (C.c_int, [vp, C.c_int, c_f64, c_f64, C.POINTER(c_f64)]),
    "qmfb_wals_epoch_host": (C.c_int, [vp, c_f64, c_f64, vp, vp, vp, C.POINTER(c_f64)]),
    "qmfb_wals_factors_device": (vp, [vp, C.c_int]),
    "qmfb_wals_stream": (vp, [vp]),
    "qmfb_wals_launch_count": (c_i64, [vp]),
    "qmfb_wals_last_timing": (C.c_int, [vp, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "qmfb_wals_sharded_create": (C.c_int, [C.c_int, C.POINTER(C.c_int), c_i64, c_i64, C.c_int, C.POINTER(vp)]),
    "qmfb_wals_sharded_destroy": (C.c_int, [vp]),
    "qmfb_wals_sharded_ndev": (C.c_int, [vp]),
    "qmfb_wals_sharded_set_csr": (C.c_int, [vp, C.c_int, p_i64, p_i32, p_f64]),
    "qmfb_wals_sharded_set_signals": (C.c_int, [vp, vp]),
    "qmfb_wals_sharded_shard": (C.c_int, [vp, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(c_i64), C.POINTER(c_i64),
                                          C.POINTER(c_i64)]),
    "qmfb_wals_sharded_set_factors": (C.c_int, [vp, C.c_int, p_f64]),
    "qmfb_wals_sharded_get_factors": (C.c_int, [vp, C.c_int, C.c_int, p_f64]),
    "qmfb_wals_sharded_half_step": (C.c_int, [vp, C.c_int, c_f64, c_f64, C.POINTER(c_f64)]),
    "qmfb_wals_sharded_epoch_host": (C.c_int, [vp, c_f64, c_f64, vp, vp, vp, C.POINTER(c_f64)]),
    "qmfb_wals_sharded_factors_device": (vp, [vp, C.c_int, C.c_int]),
    "qmfb_wals_sharded_launch_count": (c_i64, [vp]),
    "qmfb_wals_sharded_last_timing": (C.c_int, [vp, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "qmfb_bpr_create": (C.c_int, [C.c_int, c_i64, c_i64, C.c_int, C.c_int, C.POINTER(vp)]),
    "qmfb_bpr_destroy": (C.c_int, [vp]),
    "qmfb_bpr_set_data": (C.c_int, [vp, p_i32, p_i32, c_i64]),
    "qmfb_bpr_set_factors": (C.c_int, [vp, C.c_int, p_f64]),
    "qmfb_bpr_get_factors": (C.c_int, [vp, C.c_int, p_f64]),
    "qmfb_bpr_set_biases": (C.c_int, [vp, p_f64]),
    "qmfb_bpr_get_biases": (C.c_int, [vp, p_f64]),
    "qmfb_bpr_epoch": (C.c_int, [vp, c_f64, c_f64, c_f64, c_f64, C.c_int, C.c_uint64, C.c_uint64, C.c_int,
                                 C.POINTER(c_i64)]),
    "qmfb_bpr_update_triplets": (C.c_int, [vp, p_i32, p_i32, p_i32, c_i64, c_f64, c_f64, c_f64, c_f64]),
    "qmfb_bpr_eval_loss": (C.c_int, [vp, p_i32, p_i32, p_i32, c_i64, C.POINTER(c_f64)]),
    "qmfb_bpr_set_concurrency": (C.c_int, [vp, c_i64]),
    "qmfb_bpr_set_hogwild_blocks": (C.c_int, [vp, c_i64]),
    "qmfb_bpr_last_epoch_ms": (C.c_int, [vp, C.POINTER(C.c_float)]),
    "qmfb_bpr_factors_device": (vp, [vp, C.c_int]),
    "qmfb_bpr_biases_device": (vp, [vp]),
    "qmfb_bpr_launch_count": (c_i64, [vp]),
    "qmfb_eval_rank": (C.c_int, [C.c_int, p_f64, c_i64, p_f64, c_i64, C.c_int, vp, p_i32, c_i64, p_i64, p_i32, p_i32,
                                 p_f64]),
    "qmfb_eval_rank_dev": (C.c_int, [vp, vp, c_i64, vp, c_i64, c_i64, C.c_int, vp, vp, c_i64, vp, vp, c_i64, c_i64, vp,
                                     vp]),
    "qmfb_wals_eval_rank": (C.c_int, [vp, p_i32, c_i64, p_i64, p_i32, p_i32, p_f64]),
    "qmfb_wals_sharded_eval_rank": (C.c_int, [vp, p_i32, c_i64, p_i64, p_i32, p_i32, p_f64]),
    "qmfb_bpr_eval_rank": (C.c_int, [vp, p_i32, c_i64, p_i64, p_i32, p_i32, p_f64]),
    "qmfb_repeated_add": (c_f64, [c_f64, c_f64, c_i64]),
    "qmfb_rank_metrics": (C.c_int, [C.c_char_p, p_i32, p_i64, c_i64, c_i64, C.c_int, p_f64]),
    "qmfb_signals_build": (C.c_int, [C.c_int, c_i64, p_i64, p_i64, p_f64, C.POINTER(vp)]),
    "qmfb_signals_destroy": (C.c_int, [vp]),
    "qmfb_signals_device_ordinal": (C.c_int, [vp]),
    "qmfb_signals_dims": (C.c_int, [vp, C.POINTER(c_i64), C.POINTER(c_i64), C.POINTER(c_i64)]),
    "qmfb_signals_ids": (C.c_int, [vp, C.c_int, p_i64]),
    "qmfb_signals_csr": (C.c_int, [vp, C.c_int, vp, vp, vp, vp]),
    "qmfb_signals_device": (C.c_int, [vp, C.c_int, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]),
    "qmfb_wals_set_signals": (C.c_int, [vp, vp]),
}

EXPORTS = tuple(_SIG)
for _name, (_res, _args) in _SIG.items():
    _f = getattr(lib, _name)
    _f.restype = _res
    _f.argtypes = _args


def last_error():
    return lib.qmfb_last_error().decode()


def check(rc):
    if rc < 0:
        raise QmfbError(rc, last_error())
    return rc
