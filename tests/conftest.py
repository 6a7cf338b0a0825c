import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# the reference's OpenMP Gram loop is racy (SURVEY.md header): pin it before any library loads
os.environ.setdefault("OMP_NUM_THREADS", "1")
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "ref: needs oracle/_ref (the compiled reference; built where /root/reference exists)")


@pytest.fixture(scope="session")
def oracle_lib():
    import oracle
    oracle.build(ref=False)
    return oracle.oracle()


@pytest.fixture(scope="session")
def ref_lib():
    import oracle
    if not oracle.ref_available():
        try:
            oracle.build(ref=True)
        except Exception:
            pass
    if not oracle.ref_available():
        pytest.skip("oracle/_ref/libqmf_ref.so not built (no /root/reference on this box)")
    L = oracle.ref()
    L.ref_set_min_log_level(2)
    return L
