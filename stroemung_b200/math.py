"""The reference's per-cell operators (src/math.rs:19-186, src/simulation.rs:349-392),
evaluated ON THE DEVICE through the C ABI's known-answer entry points.

Arguments follow the Rust signatures: 3x3 views indexed `view[(a, b)]` with a the x
offset and b the y offset, i.e. `np.asarray(view)[a][b]`.
"""
import ctypes as C

import numpy as np

from . import _capi


def _blk(a):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(9))
    return a, a.ctypes.data_as(C.POINTER(C.c_double))


def _call(name, *args):
    out = C.c_double()
    _capi.check(getattr(_capi.lib(), name)(*args, C.byref(out)))
    return out.value


def du2dx(u_view, delx, gamma):
    a, pa = _blk(u_view)
    return _call("sb_du2dx", pa, delx, gamma)


def duvdx(u_view, v_view, delx, gamma):
    a, pa = _blk(u_view)
    b, pb = _blk(v_view)
    return _call("sb_duvdx", pa, pb, delx, gamma)


def duvdy(u_view, v_view, dely, gamma):
    a, pa = _blk(u_view)
    b, pb = _blk(v_view)
    return _call("sb_duvdy", pa, pb, dely, gamma)


def dv2dy(v_view, dely, gamma):
    a, pa = _blk(v_view)
    return _call("sb_dv2dy", pa, dely, gamma)


def laplacian(view, delx, dely):
    a, pa = _blk(view)
    return _call("sb_laplacian", pa, delx, dely)


def residual(p_view, delx, dely, rhs):
    a, pa = _blk(p_view)
    return _call("sb_residual", pa, delx, dely, rhs)


def calculate_f(u_view, v_view, delx, dely, delt, gamma, reynolds):
    a, pa = _blk(u_view)
    b, pb = _blk(v_view)
    return _call("sb_calculate_f", pa, pb, delx, dely, delt, gamma, reynolds)


def calculate_g(u_view, v_view, delx, dely, delt, gamma, reynolds):
    a, pa = _blk(u_view)
    b, pb = _blk(v_view)
    return _call("sb_calculate_g", pa, pb, delx, dely, delt, gamma, reynolds)
