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
    """`gpu` tests need a device AND the built library: skip them (instead of failing at the
    first C-ABI call) on a box that has neither."""
    gpu_items = [it for it in items if it.get_closest_marker("gpu")]
    if not gpu_items:
        return
    reason = None
    if not (ROOT / "stroemung_b200" / "libstroemung_b200.so").exists():
        reason = "stroemung_b200/libstroemung_b200.so is not built"
    elif _cuda_device_count() == 0:
        reason = "no CUDA device"
    if reason:
        skip = pytest.mark.skip(reason=reason)
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
