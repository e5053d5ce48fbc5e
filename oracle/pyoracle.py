"""ctypes binding of the CPU oracle (oracle/stroemung_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(stroemung_b200/) never imports this module.
"""
import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_DIR = Path(__file__).resolve().parent
_LIB_PATH = _DIR / "_build" / "libstroemung_oracle.so"

KIND_FLUID, KIND_NOSLIP, KIND_OUTFLOW, KIND_INFLOW, KIND_MOVING_WALL = range(5)
EDGE_NAMES = ["None", "North", "NorthEast", "East", "SouthEast", "South", "SouthWest",
              "West", "NorthWest"]
SOR_REFERENCE_ORDER, SOR_RED_BLACK = 0, 1
OK, BOUNDARY_TOO_THIN, BOUNDARY_LIST_INCORRECT, INVALID = 0, 1, 2, 4


class Params(C.Structure):
    _fields_ = [
        ("nx", C.c_uint64), ("ny", C.c_uint64),
        ("delx", C.c_double), ("dely", C.c_double),
        ("delt", C.c_double), ("gamma", C.c_double), ("reynolds", C.c_double),
        ("sor_absolute_epsilon", C.c_double), ("omega", C.c_double), ("time", C.c_double),
        ("max_iterations", C.c_uint32), ("iterations", C.c_uint32),
        ("has_initial_norm", C.c_int32),
        ("initial_norm_squared", C.c_double),
        ("tau", C.c_double),
        ("sor_mode", C.c_int32), ("reserved", C.c_int32),
    ]


class State(C.Structure):
    _fields_ = [
        ("time", C.c_double), ("delt", C.c_double),
        ("iterations", C.c_uint32), ("has_initial_norm", C.c_int32),
        ("initial_norm_squared", C.c_double),
        ("pressure_range", C.c_double * 2), ("speed_range", C.c_double * 2),
        ("fluid_cells", C.c_double), ("n_boundary", C.c_uint64),
    ]


def build(force=False):
    """Compile the oracle if the shared object is missing (gcc is in the image)."""
    src = _DIR / "stroemung_oracle.c"
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_DIR)], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(str(_LIB_PATH))
    dp = C.POINTER(C.c_double)
    u8p = C.POINTER(C.c_uint8)
    u64p = C.POINTER(C.c_uint64)
    vp = C.c_void_p
    L.so_create.argtypes = [C.POINTER(Params), dp, dp, dp, u8p, dp, dp, C.POINTER(vp), u64p]
    L.so_create.restype = C.c_int
    L.so_destroy.argtypes = [vp]
    L.so_destroy.restype = None
    L.so_rebuild_boundary_list.argtypes = [vp, u64p]
    L.so_rebuild_boundary_list.restype = C.c_int
    for name in ("so_set_boundary_u_and_v", "so_copy_pressure_to_boundaries"):
        getattr(L, name).argtypes = [vp]
        getattr(L, name).restype = C.c_int
    for name in ("so_calculate_f_and_g", "so_calculate_rhs", "so_set_u_and_v",
                 "so_calculate_pressure_range", "so_calculate_speed_range", "so_sor_sweep"):
        getattr(L, name).argtypes = [vp]
        getattr(L, name).restype = None
    L.so_calculate_norm_squared.argtypes = [vp]
    L.so_calculate_norm_squared.restype = C.c_double
    for name in ("so_solve_sor", "so_tick"):
        getattr(L, name).argtypes = [vp, C.POINTER(C.c_uint32), dp]
        getattr(L, name).restype = C.c_int
    for name in ("so_p", "so_u", "so_v", "so_f", "so_g", "so_rhs", "so_bu", "so_bv"):
        getattr(L, name).argtypes = [vp]
        getattr(L, name).restype = dp
    L.so_kind.argtypes = [vp]
    L.so_kind.restype = u8p
    L.so_render_rgba.argtypes = [vp, C.c_int, u8p]
    L.so_render_rgba.restype = None
    L.so_get_state.argtypes = [vp, C.POINTER(State)]
    L.so_get_state.restype = None
    L.so_set_params.argtypes = [vp, C.POINTER(Params)]
    L.so_set_params.restype = None
    L.so_clear_initial_norm.argtypes = [vp]
    L.so_clear_initial_norm.restype = None
    L.so_boundary_list.argtypes = [vp, u64p, u8p, C.c_uint64]
    L.so_boundary_list.restype = C.c_uint64
    d = C.c_double
    L.so_du2dx.argtypes = [dp, d, d]
    L.so_dv2dy.argtypes = [dp, d, d]
    L.so_duvdx.argtypes = [dp, dp, d, d]
    L.so_duvdy.argtypes = [dp, dp, d, d]
    L.so_laplacian.argtypes = [dp, d, d]
    L.so_residual.argtypes = [dp, d, d, d]
    L.so_calculate_f.argtypes = [dp, dp, d, d, d, d, d]
    L.so_calculate_g.argtypes = [dp, dp, d, d, d, d, d]
    for name in ("so_du2dx", "so_dv2dy", "so_duvdx", "so_duvdy", "so_laplacian",
                 "so_residual", "so_calculate_f", "so_calculate_g"):
        getattr(L, name).restype = d
    for name in ("so_preset_empty", "so_preset_simple_inflow", "so_preset_obstacle"):
        getattr(L, name).argtypes = [C.c_uint64, C.c_uint64, u8p, dp, dp]
        getattr(L, name).restype = None
    L.so_draw_circle.argtypes = [C.c_uint64, C.c_uint64, u8p, C.c_uint64, C.c_uint64, d]
    L.so_draw_circle.restype = None
    _lib = L
    return L


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _u8p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def _blk(a):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(9))
    return a, _dp(a)


# cell-level operators ---------------------------------------------------
def du2dx(u, delx, gamma):
    a, p = _blk(u)
    return lib().so_du2dx(p, delx, gamma)


def dv2dy(v, dely, gamma):
    a, p = _blk(v)
    return lib().so_dv2dy(p, dely, gamma)


def duvdx(u, v, delx, gamma):
    a, pa = _blk(u)
    b, pb = _blk(v)
    return lib().so_duvdx(pa, pb, delx, gamma)


def duvdy(u, v, dely, gamma):
    a, pa = _blk(u)
    b, pb = _blk(v)
    return lib().so_duvdy(pa, pb, dely, gamma)


def laplacian(e, delx, dely):
    a, p = _blk(e)
    return lib().so_laplacian(p, delx, dely)


def residual(pv, delx, dely, rhs):
    a, p = _blk(pv)
    return lib().so_residual(p, delx, dely, rhs)


def calculate_f(u, v, delx, dely, delt, gamma, reynolds):
    a, pa = _blk(u)
    b, pb = _blk(v)
    return lib().so_calculate_f(pa, pb, delx, dely, delt, gamma, reynolds)


def calculate_g(u, v, delx, dely, delt, gamma, reynolds):
    a, pa = _blk(u)
    b, pb = _blk(v)
    return lib().so_calculate_g(pa, pb, delx, dely, delt, gamma, reynolds)


# presets ------------------------------------------------------------------
def preset(name, nx, ny):
    kind = np.zeros((nx, ny), dtype=np.uint8)
    bu = np.zeros((nx, ny))
    bv = np.zeros((nx, ny))
    getattr(lib(), f"so_preset_{name}")(nx, ny, _u8p(kind), _dp(bu), _dp(bv))
    return kind, bu, bv


def draw_circle(kind, cx, cy, radius):
    nx, ny = kind.shape
    lib().so_draw_circle(nx, ny, _u8p(kind), cx, cy, radius)


class BoundaryTooThin(Exception):
    def __init__(self, xy):
        super().__init__(f"BoundaryTooThinError at {xy}")
        self.xy = xy


class OracleSim:
    """Handle on one oracle simulation; mirrors Simulation (src/simulation.rs:49-69)."""

    def __init__(self, nx, ny, *, delx, dely, delt, gamma, reynolds, sor_absolute_epsilon,
                 max_iterations, omega, kind, p=None, u=None, v=None, bu=None, bv=None,
                 initial_norm_squared=None, iterations=0, time=0.0, tau=0.0,
                 sor_mode=SOR_REFERENCE_ORDER):
        self.nx, self.ny = int(nx), int(ny)
        prm = Params(nx=nx, ny=ny, delx=delx, dely=dely, delt=delt, gamma=gamma,
                     reynolds=reynolds, sor_absolute_epsilon=sor_absolute_epsilon,
                     omega=omega, time=time, max_iterations=max_iterations,
                     iterations=iterations,
                     has_initial_norm=0 if initial_norm_squared is None else 1,
                     initial_norm_squared=0.0 if initial_norm_squared is None
                     else initial_norm_squared,
                     tau=tau, sor_mode=sor_mode, reserved=0)
        self.prm = prm

        def arr(a):
            if a is None:
                return None
            a = np.ascontiguousarray(a, dtype=np.float64)
            assert a.shape == (self.nx, self.ny), a.shape
            return a
        p, u, v, bu, bv = map(arr, (p, u, v, bu, bv))
        kind = np.ascontiguousarray(kind, dtype=np.uint8)
        assert kind.shape == (self.nx, self.ny)
        h = C.c_void_p()
        err = (C.c_uint64 * 2)()
        rc = lib().so_create(C.byref(prm), _dp(p), _dp(u), _dp(v), _u8p(kind), _dp(bu),
                             _dp(bv), C.byref(h), err)
        if rc == BOUNDARY_TOO_THIN:
            raise BoundaryTooThin((err[0], err[1]))
        if rc != OK:
            raise RuntimeError(f"so_create failed: {rc}")
        self._h = h

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            lib().so_destroy(h)
            self._h = None

    def _view(self, fn, dtype=np.float64):
        ptr = getattr(lib(), fn)(self._h)
        return np.ctypeslib.as_array(ptr, shape=(self.nx, self.ny))

    p = property(lambda s: s._view("so_p"))
    u = property(lambda s: s._view("so_u"))
    v = property(lambda s: s._view("so_v"))
    f = property(lambda s: s._view("so_f"))
    g = property(lambda s: s._view("so_g"))
    rhs = property(lambda s: s._view("so_rhs"))
    kind = property(lambda s: s._view("so_kind"))
    bu = property(lambda s: s._view("so_bu"))
    bv = property(lambda s: s._view("so_bv"))

    def state(self):
        st = State()
        lib().so_get_state(self._h, C.byref(st))
        return st

    def set_params(self, **kw):
        for k, val in kw.items():
            setattr(self.prm, k, val)
        lib().so_set_params(self._h, C.byref(self.prm))

    def clear_initial_norm(self):
        """sim.initial_norm_squared = None (pub field, src/simulation.rs:62)"""
        lib().so_clear_initial_norm(self._h)

    def rebuild_boundary_list(self):
        err = (C.c_uint64 * 2)()
        rc = lib().so_rebuild_boundary_list(self._h, err)
        if rc == BOUNDARY_TOO_THIN:
            raise BoundaryTooThin((err[0], err[1]))

    def boundary_list(self):
        n = lib().so_boundary_list(self._h, None, None, 0)
        idx = np.zeros(n, dtype=np.uint64)
        edge = np.zeros(n, dtype=np.uint8)
        lib().so_boundary_list(self._h, idx.ctypes.data_as(C.POINTER(C.c_uint64)),
                               _u8p(edge), n)
        return idx, edge

    def set_boundary_u_and_v(self):
        rc = lib().so_set_boundary_u_and_v(self._h)
        assert rc == OK, rc

    def calculate_f_and_g(self):
        lib().so_calculate_f_and_g(self._h)

    def calculate_rhs(self):
        lib().so_calculate_rhs(self._h)

    def copy_pressure_to_boundaries(self):
        rc = lib().so_copy_pressure_to_boundaries(self._h)
        assert rc == OK, rc

    def calculate_norm_squared(self):
        return lib().so_calculate_norm_squared(self._h)

    def sor_sweep(self):
        lib().so_sor_sweep(self._h)

    def solve_sor(self):
        it = C.c_uint32()
        nrm = C.c_double()
        rc = lib().so_solve_sor(self._h, C.byref(it), C.byref(nrm))
        assert rc == OK, rc
        return it.value, nrm.value

    def set_u_and_v(self):
        lib().so_set_u_and_v(self._h)

    def calculate_pressure_range(self):
        lib().so_calculate_pressure_range(self._h)

    def calculate_speed_range(self):
        lib().so_calculate_speed_range(self._h)

    def render_simulation(self, color_type="pressure"):
        """render_simulation (src/visualization.rs:79-105): (ny, nx, 4) uint8"""
        img = np.empty((self.ny, self.nx, 4), dtype=np.uint8)
        lib().so_render_rgba(self._h, {"pressure": 0, "speed": 1}[color_type], _u8p(img))
        return img

    def run_simulation_tick(self):
        it = C.c_uint32()
        nrm = C.c_double()
        rc = lib().so_tick(self._h, C.byref(it), C.byref(nrm))
        assert rc == OK, rc
        return it.value, nrm.value
