"""qmf_b200 — B200-native (sm_100a) replacement for the training hot path of taozhijiang/qmf.

The product is the shared library ``libqmf_b200.so`` (hand-written CUDA kernels behind the C ABI
of ``include/qmf_b200.h``) and the C++ host mirror of the reference interface in
``qmf_b200/host``.  The Python modules here are thin drivers over that ABI used by tests,
``bench.py`` and the one-process-per-GPU WALS driver.  Submodules that touch the ABI import
``qmf_b200.capi``, which raises ImportError when the CUDA library has not been built — there is
no CPU fallback.  ``qmf_b200.datagen`` (synthetic dataset shapes) has no such dependency."""

_LAZY = {
    "WalsEngineHandle": ("qmf_b200.wals", "WalsEngineHandle"),
    "csr_from_coo": ("qmf_b200.wals", "csr_from_coo"),
    "Signals": ("qmf_b200.wals", "Signals"),
    "ShardedWals": ("qmf_b200.wals_dist", "ShardedWals"),
    "BprEngineHandle": ("qmf_b200.bpr", "BprEngineHandle"),
}


def __getattr__(name):
    if name in _LAZY:
        import importlib
        mod, attr = _LAZY[name]
        return getattr(importlib.import_module(mod), attr)
    if name in ("capi", "wals", "wals_dist", "datagen", "bpr", "evalrank"):
        import importlib
        return importlib.import_module("qmf_b200." + name)
    raise AttributeError(name)
