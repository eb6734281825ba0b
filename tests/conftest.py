import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# hermetic tests: the run-time specialised kernels are compiled in every test process unless a test opts in to the on-disk
# cubin cache (tests/test_gpu_mvm.py::test_runtime_specialisation_disk_cache points it at a temporary directory)
os.environ.setdefault("COVFN_JIT_CACHE", "off")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def cf():
    import covfn_b200

    return covfn_b200


@pytest.fixture(scope="session")
def O():
    from oracle import oracle

    oracle.build()
    return oracle


def relerr(a, b):
    import numpy as np

    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    nb = np.linalg.norm(b)
    return np.linalg.norm(a - b) / (nb if nb > 0 else 1.0)
