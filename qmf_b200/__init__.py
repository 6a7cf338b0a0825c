"""qmf_b200 — B200-native (sm_100a) replacement for the training hot path of taozhijiang/qmf.

The product is the shared library ``libqmf_b200.so`` (hand-written CUDA kernels behind the C ABI
of ``include/qmf_b200.h``) and the C++ host mirror of the reference interface in
``qmf_b200/host``.  The Python modules here are thin drivers over that ABI used by tests,
``bench.py`` and the one-process-per-GPU WALS driver."""
from . import capi  # noqa: F401  (fails loudly when the CUDA library is missing)
from .wals import WalsEngineHandle, csr_from_coo  # noqa: F401
