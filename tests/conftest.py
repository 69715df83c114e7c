import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def emu_lib():
    """Path of the TEST-ONLY CPU emulation build of the kernels (tests/emu)."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    from build_emu import build_emu
    return build_emu()


@pytest.fixture(scope="session")
def gpu_ctx():
    """Context on cuda:0 through the product library; fails loudly if the extension is missing."""
    import spl_slam_b200 as S
    lib = S.api.DEFAULT_LIB
    assert os.path.exists(lib), "libplf.so is not built: run __graft_entry__.build() first"
    return S.Context(0)
