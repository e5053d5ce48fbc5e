"""CPU-side checks of the drop-in boundary: the shared library loads and exports every
symbol include/stroemung_b200.h declares (no compute calls -- there is no GPU here)."""
import re
from pathlib import Path

from stroemung_b200 import _capi

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "stroemung_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_bound_and_exported():
    decl = declared_symbols()
    assert len(decl) >= 40
    assert sorted(_capi.SYMBOLS) == decl
    L = _capi.lib()  # raises if the .so is missing or lacks a symbol
    for name in decl:
        assert hasattr(L, name), name
    assert b"sm_100a" in L.sb_version()


def test_struct_layouts_match_header():
    import ctypes as C
    # sizes follow from the header's field lists (no padding surprises)
    assert C.sizeof(_capi.Params) == 2 * 8 + 8 * 8 + 4 * 4 + 8 + 8 + 2 * 4 + 2 * 8 + 2 * 4 + 4 * 8
    assert C.sizeof(_capi.BoundaryVelocity) == 32
    assert C.sizeof(_capi.State) == 8 * 2 + 4 * 2 + 8 + 16 + 16 + 8 + 8 + 4 * 2 + 8


def test_no_cuda_device_fails_loudly():
    """Without a GPU the product path must refuse to run (no CPU fallback)."""
    import ctypes as C
    import numpy as np
    import torch
    if torch.cuda.is_available():
        return
    prm = _capi.Params(nx=4, ny=3, delx=0.1, dely=0.2, delt=0.005, gamma=0.9, reynolds=100.0,
                       sor_absolute_epsilon=1e-3, omega=1.7, max_iterations=10, device=-1)
    kind = np.zeros((4, 3), dtype=np.uint8)
    h = C.c_void_p()
    st = _capi.lib().sb_create(C.byref(prm), None, None, None,
                               kind.ctypes.data_as(C.POINTER(C.c_uint8)), None, 0, C.byref(h))
    assert st == _capi.SB_CUDA_ERROR
    assert b"no CPU fallback" in _capi.lib().sb_last_error_string() or \
        b"cuda" in _capi.lib().sb_last_error_string().lower()


def test_header_is_strict_c_and_field_offsets_match_the_bindings(tmp_path):
    """The header compiles as C99 with -pedantic -Werror (it is the C ABI, not a C++ header),
    and every field of the three structs sits at the offset the ctypes binding assumes --
    measured by the C compiler, not derived by hand."""
    import ctypes as C
    import subprocess
    structs = {"sb_params": _capi.Params, "sb_boundary_velocity": _capi.BoundaryVelocity,
               "sb_state": _capi.State}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "stroemung_b200.h"',
             'int main(void) {']
    for cname, cls in structs.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    r = subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror",
                        f"-I{ROOT / 'include'}", str(src), "-o", str(exe)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = dict(line.split() for line in subprocess.run([str(exe)], capture_output=True,
                                                       text=True).stdout.splitlines())
    for cname, cls in structs.items():
        assert int(got[cname]) == C.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(cls, fname).offset, (cname, fname)
