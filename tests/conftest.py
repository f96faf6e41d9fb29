import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_gpu():
    """True when libnewman_b200.so can create a context (nm_create != NM_ENODEV). Evaluated once, lazily."""
    global _GPU
    try:
        return _GPU
    except NameError:
        pass
    _GPU = True
    try:
        import newman_b200
        newman_b200.Device(0).close()
    except Exception as e:   # only "no CUDA device" skips; any other failure lets the GPU tests run and fail loudly
        _GPU = getattr(e, "code", None) != newman_b200._lib.NM_ENODEV if "newman_b200" in sys.modules else True
    return _GPU


def pytest_collection_modifyitems(config, items):
    """Tests marked `gpu` are SKIPPED (not failed) on a machine without a CUDA device, so a CPU-only run of the whole
    suite separates real regressions from a missing device. (`-m gpu` on the B200 box runs them.)"""
    gpu_items = [it for it in items if "gpu" in it.keywords]
    if gpu_items and not _have_gpu():
        skip = pytest.mark.skip(reason="no CUDA device (newman_b200 has no CPU fallback)")
        for it in gpu_items:
            it.add_marker(skip)


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
