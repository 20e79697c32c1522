import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def native_lib():
    """Build (if stale) and load libbfcnn_b200.so."""
    from blind_image_denoising_b200 import _native, build
    if build.needs_build():
        build.build_native()
    return _native.load_library()
