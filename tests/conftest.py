import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pgm-vae_b200"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _has_gpu():
    try:
        from pgmvae import _ffi
        return _ffi.device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def ctx():
    from pgmvae import _ffi
    if not _has_gpu():
        pytest.fail("gpu-marked test needs a CUDA device and the built libpgmvae.so (no CPU fallback)")
    return _ffi.get_context(0)
