"""Host-side mirror of the reference's `Simulation` / `SimulationGrid` API over the C ABI.

Same names, argument meaning and error behaviour as
/root/reference/src/simulation.rs:49-69, 71-120, 287-333 and
/root/reference/src/grid/mod.rs:112-153, 202-268, 343-651, so the parity tests read
like the reference's own tests.  State lives on the GPU; the reference's `pub` array
fields become properties that download on access and upload on assignment.
Nothing here computes: every method is one call into libstroemung_b200.so.
"""
import ctypes as C
import json

import numpy as np

from . import _capi, refjson
from ._capi import (BoundaryTooThinError, SbError, SOR_RED_BLACK,  # noqa: F401
                    SOR_REFERENCE_ORDER)


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _velocity_table(kind, bu, bv, x_offset=0):
    """Sparse (x, y, u, v) table of the Inflow / MovingWall cells."""
    xs, ys = np.nonzero((kind == _capi.KIND_INFLOW) | (kind == _capi.KIND_MOVING_WALL))
    tab = (_capi.BoundaryVelocity * max(len(xs), 1))()
    for i, (x, y) in enumerate(zip(xs, ys)):
        tab[i] = _capi.BoundaryVelocity(int(x) + x_offset, int(y), float(bu[x, y]),
                                        float(bv[x, y]))
    return tab, len(xs)


class BoundaryList:
    """BoundaryList (src/grid/mod.rs:61-69): x-major sorted list + fluid cell count."""

    def __init__(self, grid):
        self._grid = grid

    @property
    def sorted_boundary_list(self):
        """[((x, y), edge_name_or_None), ...] like Vec<(GridIndex, Option<EdgeType>)>."""
        idx, edge = self._grid._boundary_arrays()
        ny = self._grid.size[1]
        return [((int(i) // ny, int(i) % ny), None if e == 0 else _capi.EDGE_NAMES[e])
                for i, e in zip(idx, edge)]

    @property
    def fluid_cells(self):
        return self._grid._sim._state().fluid_cells


class SimulationGrid:
    """SimulationGrid (src/grid/mod.rs:112-125) backed by device memory."""

    def __init__(self, sim):
        self._sim = sim
        self.size = sim.size
        self.boundaries = BoundaryList(self)

    # pub pressure / u / v / cell_type ------------------------------------------
    def _get(self, field, dtype=np.float64):
        out = np.empty(self._sim._local_shape, dtype=dtype)
        self._sim._check(_capi.lib().sb_download(self._sim._h, field, out.ctypes.data))
        return out

    def _set(self, field, value, dtype=np.float64):
        a = np.ascontiguousarray(value, dtype=dtype)
        assert a.shape == self._sim._local_shape, (a.shape, self._sim._local_shape)
        self._sim._check(_capi.lib().sb_upload(self._sim._h, field, a.ctypes.data))

    pressure = property(lambda s: s._get(_capi.FIELD_P), lambda s, v: s._set(_capi.FIELD_P, v))
    u = property(lambda s: s._get(_capi.FIELD_U), lambda s, v: s._set(_capi.FIELD_U, v))
    v = property(lambda s: s._get(_capi.FIELD_V), lambda s, v: s._set(_capi.FIELD_V, v))
    cell_type = property(lambda s: s._get(_capi.FIELD_KIND, np.uint8),
                         lambda s, v: s._set(_capi.FIELD_KIND, v, np.uint8))
    edge_type = property(lambda s: s._get(_capi.FIELD_EDGE, np.uint8))

    @property
    def pressure_range(self):
        return list(self._sim._state().pressure_range)

    @property
    def speed_range(self):
        return list(self._sim._state().speed_range)

    def _boundary_arrays(self):
        L = _capi.lib()
        n = C.c_uint64()
        self._sim._check(L.sb_boundary_list(self._sim._h, None, None, 0, C.byref(n)))
        idx = np.zeros(n.value, dtype=np.uint64)
        edge = np.zeros(n.value, dtype=np.uint8)
        if n.value:
            self._sim._check(L.sb_boundary_list(
                self._sim._h, idx.ctypes.data_as(C.POINTER(C.c_uint64)),
                edge.ctypes.data_as(C.POINTER(C.c_uint8)), n.value, C.byref(n)))
        return idx, edge

    # pub fns ---------------------------------------------------------------------
    def rebuild_boundary_list(self):
        """src/grid/mod.rs:202-235; raises BoundaryTooThinError, old list stays active."""
        self._sim._check(_capi.lib().sb_rebuild_boundary_list(self._sim._h))

    def calculate_pressure_range(self):
        self._sim._check(_capi.lib().sb_calculate_pressure_range(self._sim._h))

    def calculate_speed_range(self):
        self._sim._check(_capi.lib().sb_calculate_speed_range(self._sim._h))

    def copy_pressure_to_boundaries(self):
        self._sim._check(_capi.lib().sb_copy_pressure_to_boundaries(self._sim._h))

    def set_boundary_u_and_v(self):
        self._sim._check(_capi.lib().sb_set_boundary_u_and_v(self._sim._h))

    def draw_cells(self, cell_kind, m_x, m_y, velocity=(0.0, 0.0)):
        """draw_cells of src/lib.rs:38-78; returns True if the edit was kept."""
        applied = C.c_int32()
        self._sim._check(_capi.lib().sb_edit_cells(self._sim._h, m_x, m_y, cell_kind,
                                                   velocity[0], velocity[1], C.byref(applied)))
        return bool(applied.value)


class Simulation:
    """Simulation (src/simulation.rs:49-69) on one B200 (or one row slab of a multi-GPU run)."""

    def __init__(self, handle, params):
        self._h = handle
        self._prm = params
        self.size = (int(params.nx), int(params.ny))
        xb, xe = int(params.x_begin), int(params.x_end)
        rows = (xe - xb) if params.world > 1 else self.size[0]
        self._local_shape = (rows, self.size[1])
        self.cell_size = (params.delx, params.dely)
        self.grid = SimulationGrid(self)

    # construction ----------------------------------------------------------------
    @staticmethod
    def _params(size, cell_size, delt, gamma, reynolds, sor_absolute_epsilon, max_iterations,
                omega, initial_norm_squared=None, iterations=0, time=0.0, tau=0.0,
                sor_mode=SOR_REFERENCE_ORDER, temporal_block=0, device=-1, x_begin=0, x_end=0,
                rank=0, world=0):
        return _capi.Params(
            nx=size[0], ny=size[1], delx=cell_size[0], dely=cell_size[1], delt=delt,
            gamma=gamma, reynolds=reynolds, sor_absolute_epsilon=sor_absolute_epsilon,
            omega=omega, time=time, max_iterations=max_iterations, iterations=iterations,
            has_initial_norm=0 if initial_norm_squared is None else 1, sor_mode=sor_mode,
            initial_norm_squared=0.0 if initial_norm_squared is None else initial_norm_squared,
            tau=tau, temporal_block=temporal_block, device=device, x_begin=x_begin,
            x_end=x_end, rank=rank, world=world)

    @classmethod
    def try_from(cls, unfinalized, velocity_table=None, **ext):
        """Simulation::try_from(UnfinalizedSimulation) (src/simulation.rs:71-99).

        `unfinalized`: dict with the fields of UnfinalizedSimulation; its "grid" is a dict
        with p/u/v f64 arrays [nx, ny] (or None = zeros), kind u8 and bu/bv f64 arrays
        (refjson.grid_from_json / presets.* produce it).  `ext`: tau, sor_mode,
        temporal_block, device, and the slab fields.
        """
        grid = unfinalized["grid"]
        prm = cls._params(unfinalized["size"], unfinalized["cell_size"], unfinalized["delt"],
                          unfinalized["gamma"], unfinalized["reynolds"],
                          unfinalized["sor_absolute_epsilon"], unfinalized["max_iterations"],
                          unfinalized["omega"], unfinalized.get("initial_norm_squared"),
                          unfinalized.get("iterations", 0), unfinalized.get("time", 0.0), **ext)
        kind = np.ascontiguousarray(grid["kind"], dtype=np.uint8)

        def arr(a):
            return None if a is None else np.ascontiguousarray(a, dtype=np.float64)
        p, u, v = arr(grid.get("p")), arr(grid.get("u")), arr(grid.get("v"))
        bu = grid.get("bu") if grid.get("bu") is not None else np.zeros(kind.shape)
        bv = grid.get("bv") if grid.get("bv") is not None else np.zeros(kind.shape)
        if velocity_table is not None:
            # slab mode: the merged table of stroemung_b200.multi (own + halo rows, global x)
            ntab = len(velocity_table)
            tab = (_capi.BoundaryVelocity * max(ntab, 1))()
            for i, (x, y, tu, tv) in enumerate(velocity_table):
                tab[i] = _capi.BoundaryVelocity(int(x), int(y), float(tu), float(tv))
        else:
            tab, ntab = _velocity_table(kind, bu, bv,
                                        x_offset=int(prm.x_begin) if prm.world > 1 else 0)
        h = C.c_void_p()
        st = _capi.lib().sb_create(C.byref(prm), _dp(p), _dp(u), _dp(v),
                                   kind.ctypes.data_as(C.POINTER(C.c_uint8)), tab, ntab,
                                   C.byref(h))
        _capi.check(st, None)
        return cls(h, prm)

    @classmethod
    def from_reader(cls, reader, **ext):
        """Simulation::from_reader (src/simulation.rs:117-120): the reference's JSON."""
        prm, grid = refjson.simulation_from_json(json.load(reader))
        prm["grid"] = grid
        return cls.try_from(prm, **ext)

    @classmethod
    def load(cls, path, quirk_serde_json=False, **ext):
        """A document of the reference (`serde_json::to_writer(.., &simulation)`) or of
        `save` (binary sidecar for the arrays of large grids, refjson.save_simulation)."""
        prm, grid = refjson.load_simulation(path, quirk_serde_json=quirk_serde_json)
        prm["grid"] = grid
        return cls.try_from(prm, **ext)

    # `#[derive(Serialize)] Simulation` (src/simulation.rs:49-69): f, g, rhs and the boundary
    # list are skipped, the grid writes pressure / u / v / cell_type (src/grid/mod.rs:112-125)
    def _host_state(self):
        assert self._prm.world <= 1, "serialise a slab run through stroemung_b200.multi.gather_field"
        L = _capi.lib()
        st = self._state()
        n = C.c_size_t()
        self._check(L.sb_get_boundary_velocities(self._h, None, 0, C.byref(n)))
        tab = (_capi.BoundaryVelocity * max(n.value, 1))()
        self._check(L.sb_get_boundary_velocities(self._h, tab, n.value, C.byref(n)))
        kind = self.grid.cell_type
        bu, bv = np.zeros(kind.shape), np.zeros(kind.shape)
        for i in range(n.value):
            x, y = int(tab[i].x), int(tab[i].y)
            if kind[x, y] in (_capi.KIND_INFLOW, _capi.KIND_MOVING_WALL):
                bu[x, y], bv[x, y] = tab[i].u, tab[i].v
        prm = {"size": self.size, "cell_size": self.cell_size, "delt": st.delt,
               "gamma": self._prm.gamma, "reynolds": self._prm.reynolds,
               "initial_norm_squared": st.initial_norm_squared if st.has_initial_norm else None,
               "sor_absolute_epsilon": self._prm.sor_absolute_epsilon,
               "max_iterations": int(self._prm.max_iterations), "iterations": int(st.iterations),
               "time": st.time, "omega": self._prm.omega}
        grid = {"p": self.grid.pressure, "u": self.grid.u, "v": self.grid.v, "kind": kind,
                "bu": bu, "bv": bv}
        return prm, grid

    def to_json(self):
        """The document `serde_json::to_value(&simulation)` gives (a dict)."""
        return refjson.simulation_to_json(*self._host_state())

    def save(self, path, sidecar=None):
        """Write `to_json()` to `path`; large grids keep their arrays in `path + ".bin"`
        (refjson.save_simulation).  Returns the document."""
        return refjson.save_simulation(path, *self._host_state(), sidecar=sidecar)

    @classmethod
    def from_preset(cls, preset, size, cell_size, delt, gamma, reynolds, sor_absolute_epsilon,
                    max_iterations, omega, preset_args=(), **ext):
        """Device-side mask generation (no host arrays): sb_create_preset."""
        prm = cls._params(size, cell_size, delt, gamma, reynolds, sor_absolute_epsilon,
                          max_iterations, omega, **ext)
        args = (C.c_double * max(len(preset_args), 1))(*preset_args)
        h = C.c_void_p()
        st = _capi.lib().sb_create_preset(C.byref(prm), _capi.PRESETS[preset], args,
                                          len(preset_args), C.byref(h))
        _capi.check(st, None)
        return cls(h, prm)

    # row slabs (include/stroemung_b200.h "multi-GPU"; host plumbing in multi.py) -------
    def slab_export(self):
        blob = (C.c_uint8 * _capi.SLAB_BLOB_BYTES)()
        self._check(_capi.lib().sb_slab_export(self._h, blob))
        return bytes(blob)

    def slab_connect(self, blobs):
        """collective: every rank passes all ranks' blobs in rank order"""
        from .multi import blob_buffer
        self._check(_capi.lib().sb_slab_connect(self._h, blob_buffer(blobs), len(blobs)))

    def slab_sync_halos(self):
        """collective: after assigning grid.pressure / u / v on any rank"""
        self._check(_capi.lib().sb_slab_sync_halos(self._h))

    def close(self):
        if getattr(self, "_h", None):
            _capi.lib().sb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, status):
        _capi.check(status, self._h)

    def _state(self):
        st = _capi.State()
        self._check(_capi.lib().sb_get_state(self._h, C.byref(st)))
        return st

    # pub scalar fields ---------------------------------------------------------------
    def _scalar(name):  # noqa: N805
        def get(self):
            return getattr(self._prm, name)

        def set_(self, value):
            setattr(self._prm, name, value)
            self._push_params()
        return property(get, set_)

    gamma = _scalar("gamma")
    reynolds = _scalar("reynolds")
    sor_absolute_epsilon = _scalar("sor_absolute_epsilon")
    max_iterations = _scalar("max_iterations")
    omega = _scalar("omega")
    tau = _scalar("tau")
    sor_mode = _scalar("sor_mode")
    temporal_block = _scalar("temporal_block")
    del _scalar

    def _push_params(self):
        st = self._state()
        self._prm.time, self._prm.iterations = st.time, st.iterations
        self._prm.has_initial_norm = st.has_initial_norm
        self._prm.initial_norm_squared = st.initial_norm_squared
        if self._prm.tau > 0:
            self._prm.delt = st.delt
        self._check(_capi.lib().sb_set_params(self._h, C.byref(self._prm)))

    @property
    def delt(self):
        return self._state().delt

    @delt.setter
    def delt(self, value):
        self._prm.delt = value
        self._push_params()

    time = property(lambda s: s._state().time)
    iterations = property(lambda s: s._state().iterations)

    @property
    def initial_norm_squared(self):
        st = self._state()
        return st.initial_norm_squared if st.has_initial_norm else None

    @initial_norm_squared.setter
    def initial_norm_squared(self, value):
        """pub initial_norm_squared: Option<Real> (src/simulation.rs:62); None makes the next
        solve_sor latch the norm after its first sweep (:229-237, :276)"""
        st = self._state()
        self._prm.time, self._prm.iterations = st.time, st.iterations
        if self._prm.tau > 0:
            self._prm.delt = st.delt
        self._prm.has_initial_norm = 0 if value is None else 1
        self._prm.initial_norm_squared = 0.0 if value is None else value
        self._check(_capi.lib().sb_set_params(self._h, C.byref(self._prm)))

    f = property(lambda s: s.grid._get(_capi.FIELD_F), lambda s, v: s.grid._set(_capi.FIELD_F, v))
    g = property(lambda s: s.grid._get(_capi.FIELD_G), lambda s, v: s.grid._set(_capi.FIELD_G, v))
    rhs = property(lambda s: s.grid._get(_capi.FIELD_RHS),
                   lambda s, v: s.grid._set(_capi.FIELD_RHS, v))

    # the hot path --------------------------------------------------------------------
    def run_simulation_tick(self):
        """src/simulation.rs:324-333 -> (sor_iterations, norm_squared)."""
        it, nrm = C.c_uint32(), C.c_double()
        self._check(_capi.lib().sb_tick(self._h, C.byref(it), C.byref(nrm)))
        return it.value, nrm.value

    def tick_host(self, p_in, u_in, v_in, p_out=None, u_out=None, v_out=None):
        """One tick on HOST buffers (sb_tick_host): upload p, u, v, tick, download them, one
        synchronisation.  Buffers are addresses (int / c_void_p, e.g. from sb_host_alloc) or
        C-contiguous float64 arrays of this handle's shape; outputs default to the inputs
        (in place, like the reference's tick on its own arrays)."""
        def addr(b):
            if isinstance(b, np.ndarray):
                assert b.dtype == np.float64 and b.flags.c_contiguous and b.shape == self._local_shape
                return C.c_void_p(b.ctypes.data)
            return b if isinstance(b, C.c_void_p) else C.c_void_p(b)
        outs = [i if o is None else o for i, o in zip((p_in, u_in, v_in), (p_out, u_out, v_out))]
        it, nrm = C.c_uint32(), C.c_double()
        self._check(_capi.lib().sb_tick_host(self._h, addr(p_in), addr(u_in), addr(v_in),
                                             addr(outs[0]), addr(outs[1]), addr(outs[2]),
                                             C.byref(it), C.byref(nrm)))
        return it.value, nrm.value

    def run_ticks(self, n):
        it, nrm = C.c_uint32(), C.c_double()
        self._check(_capi.lib().sb_run_ticks(self._h, n, C.byref(it), C.byref(nrm)))
        return it.value, nrm.value

    def calculate_f_and_g(self):
        self._check(_capi.lib().sb_calculate_f_and_g(self._h))

    def calculate_rhs(self):
        self._check(_capi.lib().sb_calculate_rhs(self._h))

    def calculate_norm_squared(self):
        nrm = C.c_double()
        self._check(_capi.lib().sb_calculate_norm_squared(self._h, C.byref(nrm)))
        return nrm.value

    def solve_sor(self):
        it, nrm = C.c_uint32(), C.c_double()
        self._check(_capi.lib().sb_solve_sor(self._h, C.byref(it), C.byref(nrm)))
        return it.value, nrm.value

    def sor_sweeps(self, n):
        """exactly n SOR iterations, no exit test; returns the n residual norms."""
        norms = np.zeros(max(n, 1))
        self._check(_capi.lib().sb_sor_sweeps(self._h, n, _dp(norms)))
        return norms[:n]

    def set_u_and_v(self):
        """pub fn set_u_and_v (src/simulation.rs:287-322)."""
        self._check(_capi.lib().sb_set_u_and_v(self._h))

    # instrumentation -------------------------------------------------------------------
    @property
    def kernel_launches(self):
        return int(_capi.lib().sb_kernel_launches(self._h))

    def render_simulation(self, color_type="pressure"):
        """render_simulation (src/visualization.rs:79-105) on the device: the RGBA8 frame in
        macroquad's Image layout, shape (ny, owned rows, 4) -- image[y, x] is the pixel of
        cell (x, y).  color_type: "pressure" | "speed" (ColorType, :72-77)."""
        ct = {"pressure": _capi.COLOR_PRESSURE, "speed": _capi.COLOR_SPEED}[color_type]
        img = np.empty((self.size[1], self._local_shape[0], 4), dtype=np.uint8)
        self._check(_capi.lib().sb_render_rgba(self._h, ct, img.ctypes.data))
        return img

    @property
    def rb_plan(self):
        """(tiles on the tile kernel, work items of the streaming kernel) of the last pass"""
        a, b = C.c_int32(), C.c_int32()
        self._check(_capi.lib().sb_rb_plan(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    @property
    def sor_path(self):
        """(path, ctas) of the last red-black solve: 0 pass by pass, 1 one launch on one SM
        (sor_small.cu), 2 / 3 one cooperative launch over `ctas` SMs (sor_mid.cu: generic /
        register-window kernel)"""
        n = C.c_int32()
        path = _capi.lib().sb_last_sor_path(self._h, C.byref(n))
        return int(path), int(n.value)

    def timer_begin(self):
        self._check(_capi.lib().sb_timer_begin(self._h))

    def timer_end(self):
        """device time (ms, CUDA events on the handle's stream) since timer_begin"""
        ms = C.c_double()
        self._check(_capi.lib().sb_timer_end(self._h, C.byref(ms)))
        return ms.value

    def profile_enable(self, enable=True):
        self._check(_capi.lib().sb_profile_enable(self._h, 1 if enable else 0))

    def profile_read(self):
        """durations (ms) of the SOR sweep-kernel launches since the last read"""
        buf = np.zeros(4096)
        n = C.c_size_t()
        self._check(_capi.lib().sb_profile_read(self._h, _dp(buf), buf.size, C.byref(n)))
        return buf[:min(n.value, buf.size)].copy()

    @property
    def last_stage_ms(self):
        """device ms of the last tick's stages: velocity BC, F/G + RHS, SOR, velocity update"""
        ms = np.zeros(4)
        self._check(_capi.lib().sb_last_stage_ms(self._h, _dp(ms)))
        return ms

    @property
    def last_sor_ms(self):
        return float(_capi.lib().sb_last_sor_ms(self._h))
