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
def oracle():
    import cpu_libs
    return cpu_libs.load_oracle()


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference build (oracle/_ref); tests that need it skip when it is absent."""
    import cpu_libs
    lib = cpu_libs.load_ref()
    if lib is None:
        pytest.skip("oracle/_ref/libmufft_ref.so not built (needs /root/reference)")
    return lib


@pytest.fixture(scope="session")
def gpu():
    """libmufft_b200.so behind the same numpy-level wrapper as the CPU checkers."""
    import cpu_libs
    import mufft_b200
    if mufft_b200.mufft_cuda_device_count() < 1:
        pytest.fail("no CUDA device visible: the gpu tests must run on a GPU box (there is no CPU fallback)")
    return cpu_libs._CpuLib(mufft_b200.LIB_PATH, "mufft_")
