//! `src/gpu.rs` -- the reference-side binding of libstroemung_b200.so (include/stroemung_b200.h).
//!
//! Besides adding this file, ONE change to existing code is needed: the five fields of
//! `UnfinalizedSimulationGrid` (src/grid/mod.rs:86-91: `size`, `pressure`, `u`, `v`, `cell_type`)
//! are private to `grid`; make them `pub(crate)` so that `try_from_with` below can hand the
//! arrays to the GPU without going through the CPU `SimulationGrid::try_from` first.
//!
//! A maintainer of wickedchicken/stroemung adds this file (plus `mod gpu;` in `src/lib.rs` and
//! the `build.rs` next to it) to run `Simulation::run_simulation_tick` and everything it calls
//! (src/simulation.rs:324-333) on a B200.  **Source only**: the image this library is built in has
//! no rustc / cargo, so this file has never been compiled.  What keeps it honest instead:
//! tests/test_rust_shim_abi.py parses the `extern "C"` block and the `#[repr(C)]` structs below
//! and checks them, name by name and argument by argument, against the C header; the same ABI is
//! driven end to end by the C++ mirror (include/stroemung_b200.hpp, tests/cpp/) and the Python
//! one (stroemung_b200/simulation.py), which replay the reference's own tests on the GPU.
use crate::cell::{BoundaryCell, Cell};
use crate::grid::{SimulationGridError, UnfinalizedSimulationGrid};
use crate::math::Real;
use crate::simulation::{SimulationError, UnfinalizedSimulation};
use crate::types::{CellPhysicalSize, GridArray, GridIndex, GridSize};
use crate::visualization::ColorType;
use std::os::raw::{c_char, c_void};

// ---- include/stroemung_b200.h, type for type -------------------------------------------------

/// `sb_params`: UnfinalizedSimulation (src/simulation.rs:30-44) minus the arrays, plus extensions
#[repr(C)]
#[derive(Default, Clone, Copy)]
pub struct SbParams {
    pub nx: u64,
    pub ny: u64,
    pub delx: f64,
    pub dely: f64,
    pub delt: f64,
    pub gamma: f64,
    pub reynolds: f64,
    pub sor_absolute_epsilon: f64,
    pub omega: f64,
    pub time: f64,
    pub max_iterations: u32,
    pub iterations: u32,
    pub has_initial_norm: i32,
    pub sor_mode: i32,
    pub initial_norm_squared: f64,
    pub tau: f64,
    pub temporal_block: i32,
    pub device: i32,
    pub x_begin: u64,
    pub x_end: u64,
    pub rank: i32,
    pub world: i32,
    pub reserved: [u64; 4],
}

/// `sb_boundary_velocity`: BoundaryCell::Inflow { velocity } (src/cell.rs:8) as a sparse table
#[repr(C)]
#[derive(Default, Clone, Copy)]
pub struct SbBoundaryVelocity {
    pub x: u64,
    pub y: u64,
    pub u: f64,
    pub v: f64,
}

/// `sb_state`: calculated / bookkeeping fields of Simulation and SimulationGrid
#[repr(C)]
#[derive(Default, Clone, Copy)]
pub struct SbState {
    pub time: f64,
    pub delt: f64,
    pub iterations: u32,
    pub has_initial_norm: i32,
    pub initial_norm_squared: f64,
    pub pressure_range: [f64; 2],
    pub speed_range: [f64; 2],
    pub fluid_cells: f64,
    pub n_boundary: u64,
    pub last_sor_iterations: u32,
    pub reserved: u32,
    pub last_norm_squared: f64,
}

/// opaque `sb_sim`
#[repr(C)]
pub struct SbSim {
    _private: [u8; 0],
}

pub const SB_OK: i32 = 0;
pub const SB_BOUNDARY_TOO_THIN: i32 = 1;
pub const SB_BOUNDARY_LIST_INCORRECT: i32 = 2;
pub const SB_CUDA_ERROR: i32 = 3;
pub const SB_INVALID_ARGUMENT: i32 = 4;

pub const SB_KIND_FLUID: u8 = 0;
pub const SB_KIND_NOSLIP: u8 = 1;
pub const SB_KIND_OUTFLOW: u8 = 2;
pub const SB_KIND_INFLOW: u8 = 3;

pub const SB_SOR_REFERENCE_ORDER: i32 = 0; // bit-for-bit the reference's lexicographic SOR
pub const SB_SOR_RED_BLACK: i32 = 1; // performance mode

pub const SB_FIELD_P: i32 = 0;
pub const SB_FIELD_U: i32 = 1;
pub const SB_FIELD_V: i32 = 2;
pub const SB_FIELD_F: i32 = 3;
pub const SB_FIELD_G: i32 = 4;
pub const SB_FIELD_RHS: i32 = 5;
pub const SB_FIELD_KIND: i32 = 6;
pub const SB_FIELD_EDGE: i32 = 7;

pub const SB_SLAB_BLOB_BYTES: usize = 1024;

#[link(name = "stroemung_b200")]
extern "C" {
    // construction / destruction
    fn sb_create(params: *const SbParams, p: *const f64, u: *const f64, v: *const f64, kind: *const u8, velocities: *const SbBoundaryVelocity, n_velocities: usize, out: *mut *mut SbSim) -> i32;
    fn sb_destroy(sim: *mut SbSim);
    fn sb_create_preset(params: *const SbParams, preset: i32, args: *const f64, n_args: usize, out: *mut *mut SbSim) -> i32;
    // the hot path
    fn sb_tick(sim: *mut SbSim, sor_iterations: *mut u32, norm_squared: *mut f64) -> i32;
    fn sb_run_ticks(sim: *mut SbSim, n: u32, sor_iterations: *mut u32, norm_squared: *mut f64) -> i32;
    fn sb_tick_host(sim: *mut SbSim, p_in: *const f64, u_in: *const f64, v_in: *const f64, p_out: *mut f64, u_out: *mut f64, v_out: *mut f64, sor_iterations: *mut u32, norm_squared: *mut f64) -> i32;
    // stages, one per reference function
    fn sb_set_boundary_u_and_v(sim: *mut SbSim) -> i32;
    fn sb_calculate_f_and_g(sim: *mut SbSim) -> i32;
    fn sb_calculate_rhs(sim: *mut SbSim) -> i32;
    fn sb_copy_pressure_to_boundaries(sim: *mut SbSim) -> i32;
    fn sb_calculate_norm_squared(sim: *mut SbSim, norm_squared: *mut f64) -> i32;
    fn sb_solve_sor(sim: *mut SbSim, sor_iterations: *mut u32, norm_squared: *mut f64) -> i32;
    fn sb_set_u_and_v(sim: *mut SbSim) -> i32;
    fn sb_calculate_pressure_range(sim: *mut SbSim) -> i32;
    fn sb_calculate_speed_range(sim: *mut SbSim) -> i32;
    fn sb_sor_sweeps(sim: *mut SbSim, n: u32, norms: *mut f64) -> i32;
    // state access
    fn sb_download(sim: *mut SbSim, field: i32, dst: *mut c_void) -> i32;
    fn sb_upload(sim: *mut SbSim, field: i32, src: *const c_void) -> i32;
    fn sb_host_alloc(bytes: usize) -> *mut c_void;
    fn sb_host_free(ptr: *mut c_void);
    fn sb_get_state(sim: *mut SbSim, state: *mut SbState) -> i32;
    fn sb_set_params(sim: *mut SbSim, params: *const SbParams) -> i32;
    fn sb_set_boundary_velocities(sim: *mut SbSim, v: *const SbBoundaryVelocity, n: usize) -> i32;
    fn sb_get_boundary_velocities(sim: *mut SbSim, v: *mut SbBoundaryVelocity, capacity: usize, n: *mut usize) -> i32;
    fn sb_rebuild_boundary_list(sim: *mut SbSim) -> i32;
    fn sb_boundary_list(sim: *mut SbSim, index: *mut u64, edge: *mut u8, capacity: u64, n: *mut u64) -> i32;
    fn sb_edit_cells(sim: *mut SbSim, x: u64, y: u64, kind: u8, bu: f64, bv: f64, applied: *mut i32) -> i32;
    fn sb_render_rgba(sim: *mut SbSim, color_type: i32, dst: *mut u8) -> i32;
    // errors
    fn sb_error_cell(sim: *const SbSim, xy: *mut u64, kind: *mut u8) -> i32;
    fn sb_last_error_string() -> *const c_char;
    // multi-GPU row slabs (extension)
    fn sb_slab_export(sim: *mut SbSim, blob: *mut u8) -> i32;
    fn sb_slab_connect(sim: *mut SbSim, blobs: *const u8, n_blobs: usize) -> i32;
    fn sb_slab_sync_halos(sim: *mut SbSim) -> i32;
    // cell-level operators (src/math.rs, src/simulation.rs:349-392), evaluated on the device
    fn sb_du2dx(u: *const f64, delx: f64, gamma: f64, out: *mut f64) -> i32;
    fn sb_duvdx(u: *const f64, v: *const f64, delx: f64, gamma: f64, out: *mut f64) -> i32;
    fn sb_duvdy(u: *const f64, v: *const f64, dely: f64, gamma: f64, out: *mut f64) -> i32;
    fn sb_dv2dy(v: *const f64, dely: f64, gamma: f64, out: *mut f64) -> i32;
    fn sb_laplacian(e: *const f64, delx: f64, dely: f64, out: *mut f64) -> i32;
    fn sb_residual(p: *const f64, delx: f64, dely: f64, rhs: f64, out: *mut f64) -> i32;
    fn sb_calculate_f(u: *const f64, v: *const f64, delx: f64, dely: f64, delt: f64, gamma: f64, reynolds: f64, out: *mut f64) -> i32;
    fn sb_calculate_g(u: *const f64, v: *const f64, delx: f64, dely: f64, delt: f64, gamma: f64, reynolds: f64, out: *mut f64) -> i32;
    // instrumentation
    fn sb_kernel_launches(sim: *const SbSim) -> u64;
    fn sb_last_sor_ms(sim: *const SbSim) -> f64;
    fn sb_last_stage_ms(sim: *mut SbSim, ms: *mut f64) -> i32;
    fn sb_rb_plan(sim: *const SbSim, tile_kernel_tiles: *mut i32, stream_items: *mut i32) -> i32;
    fn sb_last_sor_path(sim: *const SbSim, ctas: *mut i32) -> i32;
    fn sb_profile_enable(sim: *mut SbSim, enable: i32) -> i32;
    fn sb_profile_read(sim: *mut SbSim, ms: *mut f64, capacity: usize, n: *mut usize) -> i32;
    fn sb_timer_begin(sim: *mut SbSim) -> i32;
    fn sb_timer_end(sim: *mut SbSim, elapsed_ms: *mut f64) -> i32;
    fn sb_stream(sim: *const SbSim) -> *mut c_void;
    fn sb_version() -> *const c_char;
}

// ---- the reference's `Simulation`, device-resident ------------------------------------------------

fn last_error() -> String {
    unsafe { std::ffi::CStr::from_ptr(sb_last_error_string()) }
        .to_string_lossy()
        .into_owned()
}

/// `Cell` grid -> u8 kinds + the sparse table of Inflow velocities (x outer, y inner: standard layout)
fn flatten_cells(cells: &GridArray<Cell>) -> (Vec<u8>, Vec<SbBoundaryVelocity>) {
    let mut kind = Vec::with_capacity(cells.len());
    let mut vel = Vec::new();
    for ((x, y), c) in cells.indexed_iter() {
        kind.push(match c {
            Cell::Fluid => SB_KIND_FLUID,
            Cell::Boundary(BoundaryCell::NoSlip) => SB_KIND_NOSLIP,
            Cell::Boundary(BoundaryCell::Outflow) => SB_KIND_OUTFLOW,
            Cell::Boundary(BoundaryCell::Inflow { velocity }) => {
                vel.push(SbBoundaryVelocity { x: x as u64, y: y as u64, u: velocity[0], v: velocity[1] });
                SB_KIND_INFLOW
            }
        });
    }
    (kind, vel)
}

/// Same public face as `simulation::Simulation` (src/simulation.rs:49-69).  State lives on the
/// GPU; the host arrays the renderer reads (`visualization.rs:86-103`) are refreshed by
/// `sync_to_host()`, or never, when `render_into` draws the frame on the device.
pub struct Simulation {
    pub size: GridSize,
    pub cell_size: CellPhysicalSize,
    pub delt: Real,
    pub gamma: Real,
    pub reynolds: Real,
    pub initial_norm_squared: Option<Real>,
    pub sor_absolute_epsilon: Real,
    pub max_iterations: u32,
    pub iterations: u32,
    pub time: Real,
    pub omega: Real,
    // SimulationGrid's pub fields (src/grid/mod.rs:112-125), flattened into this struct
    pub pressure: GridArray<Real>,
    pub u: GridArray<Real>,
    pub v: GridArray<Real>,
    pub cell_type: GridArray<Cell>,
    pub pressure_range: [Real; 2],
    pub speed_range: [Real; 2],
    pub fluid_cells: Real,
    handle: *mut SbSim,
}

impl Simulation {
    fn thin_boundary(&self, handle: *const SbSim) -> SimulationGridError {
        let (mut xy, mut kind) = ([0u64; 2], 0u8);
        unsafe { sb_error_cell(handle, xy.as_mut_ptr(), &mut kind) };
        let idx: GridIndex = (xy[0] as usize, xy[1] as usize);
        SimulationGridError::BoundaryTooThinError(self.cell_type[idx].to_string(), format!("{:?}", idx))
    }

    /// `Simulation::try_from` with the extension knobs (`sor_mode`, `temporal_block`)
    pub fn try_from_with(item: UnfinalizedSimulation, sor_mode: i32, temporal_block: i32) -> Result<Self, SimulationError> {
        let g: UnfinalizedSimulationGrid = item.grid;
        let prm = SbParams {
            nx: item.size[0] as u64,
            ny: item.size[1] as u64,
            delx: item.cell_size[0],
            dely: item.cell_size[1],
            delt: item.delt,
            gamma: item.gamma,
            reynolds: item.reynolds,
            sor_absolute_epsilon: item.sor_absolute_epsilon,
            omega: item.omega,
            time: item.time,
            max_iterations: item.max_iterations,
            iterations: item.iterations,
            has_initial_norm: item.initial_norm_squared.is_some() as i32,
            initial_norm_squared: item.initial_norm_squared.unwrap_or(0.0),
            sor_mode,
            temporal_block,
            device: -1,
            ..Default::default()
        };
        let (kind, vel) = flatten_cells(&g.cell_type);
        let (p, u, v) = (g.pressure.as_standard_layout(), g.u.as_standard_layout(), g.v.as_standard_layout());
        let mut handle: *mut SbSim = std::ptr::null_mut();
        let st = unsafe { sb_create(&prm, p.as_ptr(), u.as_ptr(), v.as_ptr(), kind.as_ptr(), vel.as_ptr(), vel.len(), &mut handle) };
        let mut sim = Simulation {
            size: item.size,
            cell_size: item.cell_size,
            delt: item.delt,
            gamma: item.gamma,
            reynolds: item.reynolds,
            initial_norm_squared: item.initial_norm_squared,
            sor_absolute_epsilon: item.sor_absolute_epsilon,
            max_iterations: item.max_iterations,
            iterations: item.iterations,
            time: item.time,
            omega: item.omega,
            pressure: p.to_owned(),
            u: u.to_owned(),
            v: v.to_owned(),
            cell_type: g.cell_type,
            pressure_range: [0.0; 2],
            speed_range: [0.0; 2],
            fluid_cells: 0.0,
            handle,
        };
        match st {
            SB_OK => {
                sim.sync_to_host();
                Ok(sim)
            }
            SB_BOUNDARY_TOO_THIN => Err(sim.thin_boundary(std::ptr::null()).into()),
            // no CPU fallback: without a B200 the construction fails
            _ => panic!("stroemung_b200: sb_create failed with status {}: {}", st, last_error()),
        }
    }

    /// `Simulation::run_simulation_tick` (src/simulation.rs:324-333)
    pub fn run_simulation_tick(&mut self) -> Result<(u32, Real), SimulationError> {
        let (mut it, mut norm) = (0u32, 0.0f64);
        let st = unsafe { sb_tick(self.handle, &mut it, &mut norm) };
        assert_eq!(st, SB_OK, "stroemung_b200: {}", last_error());
        Ok((it, norm))
    }

    /// the GUI's `for _ in 0..20 { sim.run_simulation_tick() }` (src/lib.rs:214-219) in one call
    pub fn run_ticks(&mut self, n: u32) -> Result<(u32, Real), SimulationError> {
        let (mut it, mut norm) = (0u32, 0.0f64);
        let st = unsafe { sb_run_ticks(self.handle, n, &mut it, &mut norm) };
        assert_eq!(st, SB_OK, "stroemung_b200: {}", last_error());
        Ok((it, norm))
    }

    /// one tick with the HOST arrays authoritative: uploads `pressure`, `u`, `v`, ticks, downloads
    /// them again (`ndarray::Array2<f64>` in standard layout is exactly the [nx][ny] block copied)
    pub fn run_simulation_tick_on_host_arrays(&mut self) -> Result<(u32, Real), SimulationError> {
        let (mut it, mut norm) = (0u32, 0.0f64);
        let st = unsafe {
            sb_tick_host(self.handle, self.pressure.as_ptr(), self.u.as_ptr(), self.v.as_ptr(), self.pressure.as_mut_ptr(), self.u.as_mut_ptr(), self.v.as_mut_ptr(), &mut it, &mut norm)
        };
        assert_eq!(st, SB_OK, "stroemung_b200: {}", last_error());
        Ok((it, norm))
    }

    /// before rendering / inspecting on the host: what `lib.rs:221-241` reads every frame
    pub fn sync_to_host(&mut self) {
        let mut s = SbState::default();
        unsafe {
            sb_download(self.handle, SB_FIELD_P, self.pressure.as_mut_ptr() as *mut c_void);
            sb_download(self.handle, SB_FIELD_U, self.u.as_mut_ptr() as *mut c_void);
            sb_download(self.handle, SB_FIELD_V, self.v.as_mut_ptr() as *mut c_void);
            sb_get_state(self.handle, &mut s);
        }
        self.time = s.time;
        self.delt = s.delt;
        self.iterations = s.iterations;
        self.initial_norm_squared = if s.has_initial_norm != 0 { Some(s.initial_norm_squared) } else { None };
        self.pressure_range = s.pressure_range;
        self.speed_range = s.speed_range;
        self.fluid_cells = s.fluid_cells;
    }

    /// `render_simulation` (src/visualization.rs:79-105) without the field download: the frame is
    /// colour-mapped on the device straight into `image.bytes` (macroquad `Image`, RGBA8)
    pub fn render_into(&mut self, image: &mut macroquad::prelude::Image, color_type: ColorType) {
        let ct = match color_type {
            ColorType::Pressure => 0,
            ColorType::Speed => 1,
        };
        assert_eq!(image.bytes.len(), self.size[0] * self.size[1] * 4);
        let st = unsafe { sb_render_rgba(self.handle, ct, image.bytes.as_mut_ptr()) };
        assert_eq!(st, SB_OK, "stroemung_b200: {}", last_error());
    }

    /// `draw_cells` (src/lib.rs:38-78) on the device: paint the 2x2 block, re-classify, roll back
    /// when the wall would be too thin.  Returns whether the edit was kept.
    pub fn draw_cells(&mut self, cell_type: Cell, m_x: usize, m_y: usize) -> bool {
        let (kind, bu, bv) = match cell_type {
            Cell::Fluid => (SB_KIND_FLUID, 0.0, 0.0),
            Cell::Boundary(BoundaryCell::NoSlip) => (SB_KIND_NOSLIP, 0.0, 0.0),
            Cell::Boundary(BoundaryCell::Outflow) => (SB_KIND_OUTFLOW, 0.0, 0.0),
            Cell::Boundary(BoundaryCell::Inflow { velocity }) => (SB_KIND_INFLOW, velocity[0], velocity[1]),
        };
        let mut applied = 0i32;
        let st = unsafe { sb_edit_cells(self.handle, m_x as u64, m_y as u64, kind, bu, bv, &mut applied) };
        assert_eq!(st, SB_OK, "stroemung_b200: {}", last_error());
        if applied != 0 {
            for (x, y) in [(m_x, m_y), (m_x + 1, m_y), (m_x, m_y + 1), (m_x + 1, m_y + 1)] {
                if x > 0 && x < self.size[0] - 1 && y > 0 && y < self.size[1] - 1 {
                    self.cell_type[(x, y)] = cell_type;
                }
            }
        }
        applied != 0
    }

    /// `SimulationGrid::rebuild_boundary_list` (src/grid/mod.rs:202-235) after the HOST copy of
    /// `cell_type` (and the fields) was edited, as `lib.rs:56-70` does
    pub fn rebuild_boundary_list(&mut self) -> Result<(), SimulationGridError> {
        let (kind, vel) = flatten_cells(&self.cell_type);
        let st = unsafe {
            sb_upload(self.handle, SB_FIELD_KIND, kind.as_ptr() as *const c_void);
            sb_upload(self.handle, SB_FIELD_P, self.pressure.as_ptr() as *const c_void);
            sb_upload(self.handle, SB_FIELD_U, self.u.as_ptr() as *const c_void);
            sb_upload(self.handle, SB_FIELD_V, self.v.as_ptr() as *const c_void);
            sb_set_boundary_velocities(self.handle, vel.as_ptr(), vel.len());
            sb_rebuild_boundary_list(self.handle)
        };
        match st {
            SB_OK => Ok(()),
            SB_BOUNDARY_TOO_THIN => Err(self.thin_boundary(self.handle)),
            _ => panic!("stroemung_b200: {}", last_error()),
        }
    }

    /// `pub fn set_u_and_v` (src/simulation.rs:287-322)
    pub fn set_u_and_v(&mut self) {
        let st = unsafe { sb_set_u_and_v(self.handle) };
        assert_eq!(st, SB_OK, "stroemung_b200: {}", last_error());
    }
}

impl TryFrom<UnfinalizedSimulation> for Simulation {
    type Error = SimulationError;
    /// drop-in default: reference order, bit for bit (`SB_SOR_RED_BLACK` is the fast mode)
    fn try_from(item: UnfinalizedSimulation) -> Result<Self, Self::Error> {
        Simulation::try_from_with(item, SB_SOR_REFERENCE_ORDER, 0)
    }
}

impl Drop for Simulation {
    fn drop(&mut self) {
        if !self.handle.is_null() {
            unsafe { sb_destroy(self.handle) }
        }
    }
}
