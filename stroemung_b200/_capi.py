"""ctypes binding of the C ABI in include/stroemung_b200.h.

Loading fails loudly when stroemung_b200/libstroemung_b200.so is missing: there is
no CPU fallback in the product path.
"""
import ctypes as C
from pathlib import Path

import os

_DIR = Path(__file__).resolve().parent
# SB_LIB selects another build of the same library (kernel-variant experiments)
LIB_PATH = Path(os.environ.get("SB_LIB", _DIR / "libstroemung_b200.so"))

(SB_OK, SB_BOUNDARY_TOO_THIN, SB_BOUNDARY_LIST_INCORRECT, SB_CUDA_ERROR,
 SB_INVALID_ARGUMENT) = range(5)
KIND_FLUID, KIND_NOSLIP, KIND_OUTFLOW, KIND_INFLOW, KIND_MOVING_WALL = range(5)
SOR_REFERENCE_ORDER, SOR_RED_BLACK = 0, 1
COLOR_PRESSURE, COLOR_SPEED = 0, 1  # ColorType (src/visualization.rs:72-77)
SLAB_BLOB_BYTES = 1024
SLAB_HALO = 10
(FIELD_P, FIELD_U, FIELD_V, FIELD_F, FIELD_G, FIELD_RHS, FIELD_KIND, FIELD_EDGE) = range(8)
EDGE_NAMES = ["None", "North", "NorthEast", "East", "SouthEast", "South", "SouthWest", "West",
              "NorthWest"]
PRESETS = {"empty": 0, "simple_inflow": 1, "obstacle": 2, "channel_circle": 3,
           "backward_step": 4, "cavity": 5}


class Params(C.Structure):
    _fields_ = [
        ("nx", C.c_uint64), ("ny", C.c_uint64),
        ("delx", C.c_double), ("dely", C.c_double),
        ("delt", C.c_double), ("gamma", C.c_double), ("reynolds", C.c_double),
        ("sor_absolute_epsilon", C.c_double), ("omega", C.c_double), ("time", C.c_double),
        ("max_iterations", C.c_uint32), ("iterations", C.c_uint32),
        ("has_initial_norm", C.c_int32), ("sor_mode", C.c_int32),
        ("initial_norm_squared", C.c_double),
        ("tau", C.c_double),
        ("temporal_block", C.c_int32), ("device", C.c_int32),
        ("x_begin", C.c_uint64), ("x_end", C.c_uint64),
        ("rank", C.c_int32), ("world", C.c_int32),
        ("reserved", C.c_uint64 * 4),
    ]


class BoundaryVelocity(C.Structure):
    _fields_ = [("x", C.c_uint64), ("y", C.c_uint64), ("u", C.c_double), ("v", C.c_double)]


class State(C.Structure):
    _fields_ = [
        ("time", C.c_double), ("delt", C.c_double),
        ("iterations", C.c_uint32), ("has_initial_norm", C.c_int32),
        ("initial_norm_squared", C.c_double),
        ("pressure_range", C.c_double * 2), ("speed_range", C.c_double * 2),
        ("fluid_cells", C.c_double), ("n_boundary", C.c_uint64),
        ("last_sor_iterations", C.c_uint32), ("reserved", C.c_uint32),
        ("last_norm_squared", C.c_double),
    ]


# every symbol include/stroemung_b200.h declares (tests check the .so exports all of them)
SYMBOLS = [
    "sb_create", "sb_destroy", "sb_tick", "sb_tick_host", "sb_run_ticks", "sb_set_boundary_u_and_v",
    "sb_calculate_f_and_g", "sb_calculate_rhs", "sb_copy_pressure_to_boundaries",
    "sb_calculate_norm_squared", "sb_solve_sor", "sb_set_u_and_v",
    "sb_calculate_pressure_range", "sb_calculate_speed_range", "sb_sor_sweeps", "sb_download",
    "sb_upload", "sb_host_alloc", "sb_host_free", "sb_get_state", "sb_set_params",
    "sb_set_boundary_velocities", "sb_get_boundary_velocities", "sb_rebuild_boundary_list", "sb_boundary_list",
    "sb_edit_cells", "sb_render_rgba", "sb_create_preset", "sb_error_cell", "sb_last_error_string",
    "sb_slab_export", "sb_slab_connect", "sb_slab_sync_halos", "sb_du2dx", "sb_duvdx", "sb_duvdy",
    "sb_dv2dy", "sb_laplacian", "sb_residual", "sb_calculate_f", "sb_calculate_g",
    "sb_profile_enable", "sb_profile_read", "sb_timer_begin", "sb_timer_end", "sb_kernel_launches", "sb_last_sor_ms", "sb_last_stage_ms", "sb_stream",
    "sb_rb_plan", "sb_last_sor_path",
    "sb_version",
]

_lib = None


def lib():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `make -C stroemung_b200/csrc` "
            "(or __graft_entry__.build()).  stroemung_b200 has no CPU fallback.")
    L = C.CDLL(str(LIB_PATH))
    vp, dp, d = C.c_void_p, C.POINTER(C.c_double), C.c_double
    u8p, u32p, u64p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
    i32p = C.POINTER(C.c_int32)
    sig = {
        "sb_create": ([C.POINTER(Params), dp, dp, dp, u8p, C.POINTER(BoundaryVelocity),
                       C.c_size_t, C.POINTER(vp)], C.c_int),
        "sb_create_preset": ([C.POINTER(Params), C.c_int32, dp, C.c_size_t, C.POINTER(vp)],
                             C.c_int),
        "sb_destroy": ([vp], None),
        "sb_tick": ([vp, u32p, dp], C.c_int),
        "sb_tick_host": ([vp, vp, vp, vp, vp, vp, vp, u32p, dp], C.c_int),
        "sb_run_ticks": ([vp, C.c_uint32, u32p, dp], C.c_int),
        "sb_calculate_norm_squared": ([vp, dp], C.c_int),
        "sb_solve_sor": ([vp, u32p, dp], C.c_int),
        "sb_sor_sweeps": ([vp, C.c_uint32, dp], C.c_int),
        "sb_download": ([vp, C.c_int, vp], C.c_int),
        "sb_upload": ([vp, C.c_int, vp], C.c_int),
        "sb_host_alloc": ([C.c_size_t], vp),
        "sb_host_free": ([vp], None),
        "sb_get_state": ([vp, C.POINTER(State)], C.c_int),
        "sb_set_params": ([vp, C.POINTER(Params)], C.c_int),
        "sb_set_boundary_velocities": ([vp, C.POINTER(BoundaryVelocity), C.c_size_t], C.c_int),
        "sb_get_boundary_velocities": ([vp, C.POINTER(BoundaryVelocity), C.c_size_t,
                                        C.POINTER(C.c_size_t)], C.c_int),
        "sb_rebuild_boundary_list": ([vp], C.c_int),
        "sb_boundary_list": ([vp, u64p, u8p, C.c_uint64, u64p], C.c_int),
        "sb_edit_cells": ([vp, C.c_uint64, C.c_uint64, C.c_uint8, d, d, i32p], C.c_int),
        "sb_render_rgba": ([vp, C.c_int32, vp], C.c_int),
        "sb_error_cell": ([vp, u64p, u8p], C.c_int),
        "sb_last_error_string": ([], C.c_char_p),
        "sb_slab_export": ([vp, u8p], C.c_int),
        "sb_slab_connect": ([vp, u8p, C.c_size_t], C.c_int),
        "sb_slab_sync_halos": ([vp], C.c_int),
        "sb_du2dx": ([dp, d, d, dp], C.c_int),
        "sb_duvdx": ([dp, dp, d, d, dp], C.c_int),
        "sb_duvdy": ([dp, dp, d, d, dp], C.c_int),
        "sb_dv2dy": ([dp, d, d, dp], C.c_int),
        "sb_laplacian": ([dp, d, d, dp], C.c_int),
        "sb_residual": ([dp, d, d, d, dp], C.c_int),
        "sb_calculate_f": ([dp, dp, d, d, d, d, d, dp], C.c_int),
        "sb_calculate_g": ([dp, dp, d, d, d, d, d, dp], C.c_int),
        "sb_profile_enable": ([vp, C.c_int32], C.c_int),
        "sb_profile_read": ([vp, dp, C.c_size_t, C.POINTER(C.c_size_t)], C.c_int),
        "sb_timer_begin": ([vp], C.c_int),
        "sb_timer_end": ([vp, dp], C.c_int),
        "sb_kernel_launches": ([vp], C.c_uint64),
        "sb_rb_plan": ([vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32)], C.c_int),
        "sb_last_sor_path": ([vp, C.POINTER(C.c_int32)], C.c_int32),
        "sb_last_sor_ms": ([vp], C.c_double),
        "sb_last_stage_ms": ([vp, dp], C.c_int),
        "sb_stream": ([vp], vp),
        "sb_version": ([], C.c_char_p),
    }
    for name in ("sb_set_boundary_u_and_v", "sb_calculate_f_and_g", "sb_calculate_rhs",
                 "sb_copy_pressure_to_boundaries", "sb_set_u_and_v",
                 "sb_calculate_pressure_range", "sb_calculate_speed_range"):
        sig[name] = ([vp], C.c_int)
    for name in SYMBOLS:
        fn = getattr(L, name)  # AttributeError if the .so does not export it
        fn.argtypes, fn.restype = sig[name]
    _lib = L
    return L


class SbError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"sb_status {status}: {message}")
        self.status = status


class BoundaryTooThinError(SbError):
    """SimulationGridError::BoundaryTooThinError (src/grid/mod.rs:57-58)."""

    def __init__(self, xy, kind, message):
        super().__init__(SB_BOUNDARY_TOO_THIN, f"{message} at {xy}")
        self.xy = xy
        self.kind = kind


def check(status, handle=None):
    if status == SB_OK:
        return
    L = lib()
    msg = (L.sb_last_error_string() or b"").decode()
    if status == SB_BOUNDARY_TOO_THIN:
        xy = (C.c_uint64 * 2)()
        kind = C.c_uint8()
        L.sb_error_cell(handle, xy, C.byref(kind))
        raise BoundaryTooThinError((int(xy[0]), int(xy[1])), int(kind.value), msg)
    raise SbError(status, msg)
