import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracles_built():
    import oracles
    oracles.build_oracles()
    return oracles


@pytest.fixture(scope="session")
def dev():
    import newman_b200
    d = newman_b200.Device(0)
    # device-level tests exercise the level kernels (k3_fast / k3_level) on their small fixtures; the
    # run-to-completion kernel that production uses for such small frames has its own module (test_gpu_finish.py)
    d.set_option(newman_b200._lib.OPT_K3_FINISH_MAX, 0)
    yield d
    d.close()
