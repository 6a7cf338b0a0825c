"""CPU checks of the drop-in boundary: libqmf_b200.so loads and exports every symbol that
include/qmf_b200.h declares; the pure-host helpers work; compute entry points fail LOUDLY (no
silent CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "qmf_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qmfb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from qmf_b200 import capi
    names = declared_symbols()
    assert len(names) >= 35
    raw = C.CDLL(capi.LIB_PATH)
    missing = [n for n in names if not hasattr(raw, n)]
    assert not missing, missing
    assert sorted(capi.EXPORTS) == names, set(names) ^ set(capi.EXPORTS)


def test_layout_helpers():
    from qmf_b200 import capi
    lib = capi.lib
    assert [lib.qmfb_padded_k(k) for k in (1, 30, 32, 33, 64, 100, 128)] == [32, 32, 32, 64, 64, 128, 128]
    assert [lib.qmfb_padded_k(k) for k in (129, 160, 200, 256)] == [160, 160, 224, 256]
    assert lib.qmfb_padded_k(0) < 0
    assert lib.qmfb_padded_k(257) < 0 and "nfactors" in capi.last_error()
    assert lib.qmfb_gram_packed_len(256) == 528 * 64
    assert lib.qmfb_gram_packed_len(128) == 136 * 64
    assert lib.qmfb_gram_packed_len(30) == 10 * 64
    assert lib.qmfb_gram_workspace_len(64) == 36 * 64 * 296
    assert lib.qmfb_version() >= 100


def test_no_cpu_fallback_without_a_device():
    from qmf_b200 import capi
    if capi.lib.qmfb_device_count() > 0:
        pytest.skip("a CUDA device is present")
    h = C.c_void_p()
    rc = capi.lib.qmfb_wals_create(0, 10, 10, 8, C.byref(h))
    assert rc == -2 and "failed" in capi.last_error()        # QMFB_ERR_CUDA, loud
    rc = capi.lib.qmfb_bpr_create(0, 10, 10, 8, 0, C.byref(h))
    assert rc == -2
    # ingest and the shareable replicas: same contract
    u = np.array([1, 2, 3], np.int64)
    rc = capi.lib.qmfb_signals_build(0, 3, u, u, np.ones(3), C.byref(h))
    assert rc == -2 and "failed" in capi.last_error()
    rc = capi.lib.qmfb_ipc_alloc(0, 1024, C.byref(h), C.create_string_buffer(64))
    assert rc == -2
    with pytest.raises(capi.QmfbError):
        from qmf_b200.wals import Signals
        Signals(u, u, np.ones(3))
    with pytest.raises(capi.QmfbError):
        from qmf_b200.wals import WalsEngineHandle
        WalsEngineHandle(4, 4, 8)


def test_product_never_imports_the_oracle():
    """the oracle is test infrastructure: nothing under qmf_b200/ may reference it"""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "qmf_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"\bimport oracle\b|from oracle\b|qmf_oracle|libqmf_ref|oracle/", text):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad
