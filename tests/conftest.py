import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _native_libs():
    """Build (if stale) the scene library, the oracle and -- when nvcc is present -- the CUDA library."""
    from surely_raytracing_b200 import build as b
    b.build_scenes_library()
    try:
        b.build_cuda_library()
    except Exception as e:  # no nvcc on this machine: the prebuilt .so must already be there
        if not (ROOT / "surely_raytracing_b200" / "librtb200.so").exists():
            raise
    from oracle import orc
    orc.build()
    yield


def has_gpu() -> bool:
    try:
        from surely_raytracing_b200 import capi
        return capi.load_library().rtb_device_count() > 0
    except Exception:
        return False
