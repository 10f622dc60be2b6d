import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import __graft_entry__ as ge  # noqa: E402

ge.load_package()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def built():
    """Library + checkers built (nvcc cross-compiles without a GPU)."""
    ge.build()
    return True


@pytest.fixture(scope="session")
def po(built):
    return ge.load_oracle()


@pytest.fixture(scope="session")
def oracle(po):
    return po.Oracle()


@pytest.fixture(scope="session")
def ref(po):
    if not po.have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    return po.Ref()


@pytest.fixture(scope="session")
def gpu(built):
    from h264_scroll_encoder_b200 import binding
    return binding.GpuComposer(0)
