"""TEST INFRASTRUCTURE — ctypes loaders for the CPU checker libraries.

* ``oracle()``  -> ``libqmf_oracle.so``: plain-C restatement of the reference hot path
  (``oracle/qmf_oracle.c``).
* ``ref()``     -> ``oracle/_ref/libqmf_ref.so``: the UNMODIFIED reference sources compiled from
  ``/root/reference`` plus ``oracle/ref_harness.cpp`` (built by ``make -C oracle ref``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import this module.  Nothing under ``qmf_b200/`` does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REFERENCE = os.environ.get("QMF_REFERENCE_DIR", "/root/reference")

c_i64 = C.c_int64
c_f64 = C.c_double
p_i64 = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
p_i32 = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
p_f64 = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
p_opt = C.c_void_p  # nullable pointer


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def build(ref=True):
    """Compile the checker libraries (building the checker is not using it)."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])
    if ref and os.path.isdir(os.path.join(_REFERENCE, "qmf")):
        subprocess.check_call(["make", "-s", "-j8", "-C", _HERE, "ref", "REF=" + _REFERENCE])


_oracle = None
_ref = None


def oracle():
    global _oracle
    if _oracle is not None:
        return _oracle
    path = os.path.join(_HERE, "libqmf_oracle.so")
    if not os.path.exists(path):
        build(ref=False)
    L = C.CDLL(path)
    L.qmfo_group_signals.restype = c_i64
    L.qmfo_group_signals.argtypes = [p_i64, p_i64, c_i64, p_i64, p_i64, p_i64]
    L.qmfo_gram.restype = None
    L.qmfo_gram.argtypes = [p_f64, c_i64, c_i64, p_f64]
    L.qmfo_sysv_upper.restype = C.c_int
    L.qmfo_sysv_upper.argtypes = [p_f64, c_i64, p_f64, p_i32]
    L.qmfo_wals_update_row.restype = c_f64
    L.qmfo_wals_update_row.argtypes = [p_f64, c_i64, p_i32, p_f64, c_i64, p_f64, c_f64, c_f64, p_f64]
    L.qmfo_wals_half_step.restype = c_f64
    L.qmfo_wals_half_step.argtypes = [p_f64, c_i64, p_f64, c_i64, c_i64, p_i64, p_i32, p_f64, c_f64, c_f64, c_i64,
                                      c_i64, c_i64]
    L.qmfo_bpr_predict_difference.restype = c_f64
    L.qmfo_bpr_predict_difference.argtypes = [p_f64, p_f64, p_opt, c_i64, c_i64, c_i64, c_i64]
    L.qmfo_bpr_update.restype = c_f64
    L.qmfo_bpr_update.argtypes = [p_f64, p_f64, p_opt, c_i64, c_i64, c_i64, c_i64, c_f64, c_f64, c_f64, c_f64]
    L.qmfo_bpr_eval_loss.restype = c_f64
    L.qmfo_bpr_eval_loss.argtypes = [p_f64, p_f64, p_opt, c_i64, p_i64, p_i64, p_i64, c_i64, c_i64]
    L.qmfo_bpr_sample_negatives.restype = None
    L.qmfo_bpr_sample_negatives.argtypes = [p_i64, c_i64, c_i64, c_i64, p_i64, p_i64, C.c_uint32, p_i64]
    L.qmfo_compute_test_scores.restype = None
    L.qmfo_compute_test_scores.argtypes = [p_f64, p_f64, p_opt, c_i64, c_i64, p_i64, c_i64, p_f64]
    L.qmfo_metric_one.restype = c_f64
    L.qmfo_metric_one.argtypes = [C.c_int, c_i64, p_f64, p_f64, c_i64]
    L.qmfo_metric_avg.restype = c_f64
    L.qmfo_metric_avg.argtypes = [C.c_int, c_i64, p_f64, p_f64, c_i64, c_i64, c_i64]
    L.qmfo_rank_stats.restype = None
    L.qmfo_rank_stats.argtypes = [p_f64, p_f64, c_i64, p_i64, c_i64, p_i64, p_f64]
    L.qmfo_save_factors.restype = c_i64
    L.qmfo_save_factors.argtypes = [p_f64, p_opt, p_i64, c_i64, c_i64, C.c_char_p, c_i64]
    _oracle = L
    return L


def ref_available():
    return os.path.exists(os.path.join(_HERE, "_ref", "libqmf_ref.so"))


def ref():
    """The unmodified reference behind oracle/ref_harness.cpp.  Needs OMP_NUM_THREADS=1 in the
    environment BEFORE first use (the reference's OpenMP Gram loop is racy, SURVEY.md header)."""
    global _ref
    if _ref is not None:
        return _ref
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    path = os.path.join(_HERE, "_ref", "libqmf_ref.so")
    if not os.path.exists(path):
        raise FileNotFoundError(path + " (run `make -C oracle ref` where /root/reference is mounted)")
    L = C.CDLL(path)
    vp = C.c_void_p
    sig = {
        "ref_wals_create": (vp, [c_i64, c_i64, c_f64, c_f64, C.c_int, C.c_char_p, c_i64, C.c_int, C.c_int32]),
        "ref_wals_destroy": (None, [vp]),
        "ref_wals_init": (None, [vp, p_i64, p_i64, p_f64, c_i64]),
        "ref_wals_init_test": (None, [vp, p_i64, p_i64, p_f64, c_i64]),
        "ref_wals_nusers": (c_i64, [vp]),
        "ref_wals_nitems": (c_i64, [vp]),
        "ref_wals_nnz": (c_i64, [vp, C.c_int]),
        "ref_wals_ids": (None, [vp, C.c_int, p_i64]),
        "ref_wals_csr": (None, [vp, C.c_int, p_i64, p_i64, p_i32, p_i64, p_f64]),
        "ref_wals_set_factors": (None, [vp, C.c_int, p_f64]),
        "ref_wals_get_factors": (None, [vp, C.c_int, p_f64]),
        "ref_wals_half_step": (c_f64, [vp, C.c_int]),
        "ref_wals_evaluate": (None, [vp, c_i64]),
        "ref_wals_optimize": (None, [vp]),
        "ref_wals_save": (None, [vp, C.c_char_p, C.c_char_p]),
        "ref_gram": (None, [p_f64, c_i64, c_i64, C.c_int, C.c_int, p_f64]),
        "ref_wals_update_one": (c_f64, [p_f64, c_i64, c_i64, p_f64, c_i64, c_i64, p_i32, p_f64, c_i64, p_f64, c_f64,
                                        c_f64]),
        "ref_wals_update_rows": (c_f64, [p_f64, c_i64, p_f64, c_i64, c_i64, p_i64, p_i32, p_f64, p_f64, c_f64, c_f64,
                                         C.c_int, C.POINTER(c_f64)]),
        "ref_wals_update_rows_losses": (c_f64, [p_f64, c_i64, p_f64, c_i64, c_i64, p_i64, p_i32, p_f64, p_f64, c_f64,
                                                c_f64, C.c_int, p_f64]),
        "ref_linear_symmetric_solve": (None, [p_f64, p_f64, c_i64, p_f64]),
        "ref_bpr_create": (vp, [c_i64, c_i64, c_f64, c_f64, c_f64, c_f64, c_f64, C.c_int, c_f64, c_i64, c_i64, C.c_int,
                                c_i64, C.c_int32, C.c_int, C.c_char_p, c_i64, C.c_int, c_i64]),
        "ref_bpr_destroy": (None, [vp]),
        "ref_bpr_init": (None, [vp, p_i64, p_i64, p_f64, c_i64]),
        "ref_bpr_init_test": (None, [vp, p_i64, p_i64, p_f64, c_i64]),
        "ref_bpr_nusers": (c_i64, [vp]),
        "ref_bpr_nitems": (c_i64, [vp]),
        "ref_bpr_ids": (None, [vp, C.c_int, p_i64]),
        "ref_bpr_ndata": (c_i64, [vp]),
        "ref_bpr_data": (None, [vp, p_i64, p_i64]),
        "ref_bpr_eval_size": (c_i64, [vp, C.c_int]),
        "ref_bpr_eval_set": (None, [vp, C.c_int, p_i64, p_i64, p_i64]),
        "ref_bpr_get_factors": (None, [vp, C.c_int, p_f64]),
        "ref_bpr_set_factors": (None, [vp, C.c_int, p_f64]),
        "ref_bpr_get_biases": (None, [vp, p_f64]),
        "ref_bpr_set_biases": (None, [vp, p_f64]),
        "ref_bpr_update": (None, [vp, c_i64, c_i64, c_i64]),
        "ref_bpr_predict_difference": (c_f64, [vp, c_i64, c_i64, c_i64]),
        "ref_bpr_learning_rate": (c_f64, [vp]),
        "ref_bpr_set_learning_rate": (None, [vp, c_f64]),
        "ref_bpr_eval_loss": (c_f64, [vp, C.c_int]),
        "ref_bpr_optimize": (c_f64, [vp]),
        "ref_bpr_num_test_users": (c_i64, [vp]),
        "ref_bpr_test_users": (None, [vp, p_i64]),
        "ref_compute_test_scores": (None, [p_f64, c_i64, p_f64, c_i64, c_i64, p_opt, p_i64, c_i64, C.c_int, p_f64]),
        "ref_init_avg_test_data": (c_i64, [p_i64, c_i64, p_i64, c_i64, p_i64, p_i64, p_f64, c_i64, c_i64, C.c_int32,
                                           p_opt, p_opt]),
        "ref_metric_one": (c_f64, [C.c_char_p, p_f64, p_f64, c_i64]),
        "ref_metric_avg": (c_f64, [C.c_char_p, p_f64, p_f64, c_i64, c_i64, C.c_int]),
        "ref_save_factors": (c_i64, [p_f64, p_opt, p_i64, c_i64, c_i64, C.c_char_p, c_i64]),
        "ref_read_dataset": (c_i64, [C.c_char_p, p_opt, p_opt, p_opt, c_i64]),
        "ref_set_min_log_level": (None, [C.c_int]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _ref = L
    return L


METRIC_KIND = {"mse": 0, "auc": 1, "ap": 2, "p": 3, "r": 4}


def metric_kind(name):
    """'auc' -> (1, 0); 'p@10' -> (3, 10)"""
    if "@" in name:
        m, k = name.split("@")
        return METRIC_KIND[m], int(k)
    return METRIC_KIND[name], 0


ptr = _ptr
