import json
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = Path(__file__).resolve().parent / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _cuda_device_count():
    """devices the CUDA runtime sees, without importing torch (0 on a CPU-only box)"""
    import ctypes
    for name in ("libcudart.so", "libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so"):
        try:
            rt = ctypes.CDLL(name)
        except OSError:
            continue
        n = ctypes.c_int(0)
        return n.value if rt.cudaGetDeviceCount(ctypes.byref(n)) == 0 else 0
    return 0


def pytest_collection_modifyitems(config, items):
    """`gpu` tests need a CUDA device: without one they are skipped (instead of failing at the
    first C-ABI call).  WITH a device they always run -- a missing library then fails them
    loudly (stroemung_b200._capi.lib raises ImportError): there is no fallback to skip to."""
    gpu_items = [it for it in items if it.get_closest_marker("gpu")]
    if not gpu_items or _cuda_device_count() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in gpu_items:
        it.add_marker(skip)


@pytest.fixture(scope="session")
def kat():
    return json.loads((GOLDEN / "kat.json").read_text())


@pytest.fixture(scope="session")
def snapshots():
    return json.loads((GOLDEN / "snapshots.json").read_text())


@pytest.fixture(scope="session")
def fixtures():
    return json.loads((GOLDEN / "fixtures.json").read_text())
