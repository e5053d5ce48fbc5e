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


def test_plain_c_host_links_and_fails_loudly_without_a_device(tmp_path):
    """A C99 host (no C++ runtime of its own, no Python) links against the library by name and
    drives the construction entry point: without a CUDA device it gets SB_CUDA_ERROR and the
    "no CPU fallback" message; with one it constructs and ticks the 4x3 case of the reference's
    `simulation_tick` test (src/simulation.rs:571-594) and must see 100 SOR iterations."""
    import subprocess
    src = tmp_path / "host.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "stroemung_b200.h"
int main(void) {
    sb_params p;
    unsigned char kind[12] = {1, 3, 1,  1, 0, 1,  1, 0, 1,  1, 2, 1};   /* simple_inflow([4, 3]) */
    sb_boundary_velocity inflow = {0, 1, 1.0, 0.0};
    sb_sim *sim = NULL;
    uint32_t it = 0;
    double norm = 0.0;
    sb_status st;
    memset(&p, 0, sizeof p);
    p.nx = 4; p.ny = 3; p.delx = 0.1; p.dely = 0.2; p.delt = 0.005; p.gamma = 0.9;
    p.reynolds = 100.0; p.sor_absolute_epsilon = 0.001; p.omega = 1.7; p.max_iterations = 100;
    p.sor_mode = SB_SOR_REFERENCE_ORDER; p.device = -1;
    st = sb_create(&p, NULL, NULL, NULL, kind, &inflow, 1, &sim);
    if (st != SB_OK) {
        printf("status %d: %s\n", (int)st, sb_last_error_string());
        return sim == NULL ? 10 + (int)st : 99;
    }
    st = sb_tick(sim, &it, &norm);
    printf("tick status %d iterations %u norm %.17g\n", (int)st, (unsigned)it, norm);
    sb_destroy(sim);
    return st == SB_OK && it == 100 ? 0 : 1;
}
''')
    exe = tmp_path / "host"
    lib_dir = ROOT / "stroemung_b200"
    r = subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror",
                        f"-I{ROOT / 'include'}", str(src), "-o", str(exe), f"-L{lib_dir}",
                        "-lstroemung_b200", f"-Wl,-rpath,{lib_dir}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    from tests.conftest import _cuda_device_count
    if _cuda_device_count() == 0:
        assert r.returncode == 10 + _capi.SB_CUDA_ERROR, (r.returncode, r.stdout)
        assert "no CPU fallback" in r.stdout or "cuda" in r.stdout.lower()
    else:
        assert r.returncode == 0, r.stdout
        assert "iterations 100 norm 562901.74471991" in r.stdout
