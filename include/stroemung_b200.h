/*
 * stroemung_b200.h -- C ABI of the B200-native (sm_100a) per-timestep solver.
 *
 * Drop-in boundary for ONE path of wickedchicken/stroemung 0.1.2:
 * `Simulation::run_simulation_tick` (src/simulation.rs:324-333) and everything
 * it calls.  The reference has no FFI of its own; the seam is the `pub` surface
 * of src/simulation.rs, src/grid/mod.rs and src/math.rs.  Each entry point below
 * names the reference item it replaces (file:line, relative to the reference
 * repo).  INTEGRATION.md shows the Rust `extern "C"` binding a maintainer adds.
 *
 * Conventions
 *   - Host arrays are caller-owned, row-major [nx][ny] with index (x, y) and y
 *     contiguous (src/types.rs:8-17); f64 fields, u8 cell kinds.
 *   - Device buffers are library-owned.  One handle is used from one host thread
 *     at a time (matches `&mut self`).
 *   - Every function returns an sb_status; there is NO CPU fallback -- without a
 *     CUDA device sb_create fails with SB_CUDA_ERROR.
 */
#ifndef STROEMUNG_B200_H
#define STROEMUNG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sb_sim sb_sim;

/* SimulationError / SimulationGridError (src/simulation.rs:22-28, src/grid/mod.rs:51-59) */
typedef enum {
    SB_OK = 0,
    SB_BOUNDARY_TOO_THIN = 1,       /* BoundaryTooThinError: see sb_error_cell      */
    SB_BOUNDARY_LIST_INCORRECT = 2, /* BoundaryListIncorrectError                   */
    SB_CUDA_ERROR = 3,              /* CUDA / NCCL failure: see sb_last_error_string */
    SB_INVALID_ARGUMENT = 4
} sb_status;

/* Cell / BoundaryCell (src/cell.rs:6-23) as a u8; 4 is an extension */
typedef enum {
    SB_KIND_FLUID = 0,
    SB_KIND_NOSLIP = 1,
    SB_KIND_OUTFLOW = 2,
    SB_KIND_INFLOW = 3,
    SB_KIND_MOVING_WALL = 4 /* extension: wall moving with its (u, v) (lid-driven cavity) */
} sb_kind;

/* Option<EdgeType> (src/grid/mod.rs:19-49); the neighbour indices are implied */
typedef enum {
    SB_EDGE_NONE = 0, SB_EDGE_N = 1, SB_EDGE_NE = 2, SB_EDGE_E = 3, SB_EDGE_SE = 4,
    SB_EDGE_S = 5, SB_EDGE_SW = 6, SB_EDGE_W = 7, SB_EDGE_NW = 8
} sb_edge;

typedef enum {
    /* lexicographic in-place SOR exactly as src/simulation.rs:250-274 (wavefront kernel),
     * strict IEEE arithmetic everywhere: results equal the reference's bit for bit
     * (the residual norm to summation order) */
    SB_SOR_REFERENCE_ORDER = 0,
    /* performance mode: red-black ordering, fused residual norm, `temporal_block`
     * sweeps per pass through shared memory, FMA arithmetic */
    SB_SOR_RED_BLACK = 1
} sb_sor_mode;

typedef enum {
    SB_FIELD_P = 0, SB_FIELD_U = 1, SB_FIELD_V = 2,     /* grid.pressure / u / v (f64) */
    SB_FIELD_F = 3, SB_FIELD_G = 4, SB_FIELD_RHS = 5,   /* sim.f / g / rhs (f64)       */
    SB_FIELD_KIND = 6,                                  /* grid.cell_type as sb_kind (u8) */
    SB_FIELD_EDGE = 7                                   /* boundary edge class (u8, read-only) */
} sb_field;

/* ColorType (src/visualization.rs:72-77) */
typedef enum { SB_COLOR_PRESSURE = 0, SB_COLOR_SPEED = 1 } sb_color_type;

/* UnfinalizedSimulation (src/simulation.rs:30-44) minus the arrays, plus extensions.
 * Zero-initialise, then fill. */
typedef struct {
    uint64_t nx, ny;              /* size                                        */
    double delx, dely;            /* cell_size                                   */
    double delt, gamma, reynolds;
    double sor_absolute_epsilon, omega, time;
    uint32_t max_iterations, iterations;
    int32_t has_initial_norm;     /* initial_norm_squared: Option<Real>          */
    int32_t sor_mode;             /* sb_sor_mode                                 */
    double initial_norm_squared;
    /* ---- extensions (0 = reference behaviour / defaults) ---- */
    double tau;                   /* > 0: adaptive delt (NaSt2D COMP_delt)       */
    int32_t temporal_block;       /* red-black sweeps fused per pass, 1..4; 0 = default (4) */
    int32_t device;               /* CUDA device ordinal; -1 = current device    */
    /* row-slab decomposition along x: this handle owns global rows
     * [x_begin, x_end) of the nx rows; 0,0 = whole grid.  Host arrays passed to
     * sb_create / sb_upload / sb_download then cover only the owned rows. */
    uint64_t x_begin, x_end;
    int32_t rank, world;          /* position in the slab chain (0,0|1 = single GPU) */
    uint64_t reserved[4];
} sb_params;

/* Sparse table of boundary velocities: BoundaryCell::Inflow { velocity }
 * (src/cell.rs:8) and the MovingWall extension.  (x, y) are GLOBAL indices. */
typedef struct {
    uint64_t x, y;
    double u, v;
} sb_boundary_velocity;

/* Calculated / bookkeeping fields of Simulation and SimulationGrid */
typedef struct {
    double time;                  /* sim.time        (src/simulation.rs:66) */
    double delt;                  /* sim.delt (changes only when tau > 0)   */
    uint32_t iterations;          /* sim.iterations  (src/simulation.rs:65) */
    int32_t has_initial_norm;
    double initial_norm_squared;  /* sim.initial_norm_squared (:62)         */
    double pressure_range[2];     /* grid.pressure_range (src/grid/mod.rs:122) */
    double speed_range[2];        /* grid.speed_range    (src/grid/mod.rs:124) */
    double fluid_cells;           /* boundaries.fluid_cells (src/grid/mod.rs:65) */
    uint64_t n_boundary;          /* boundaries.sorted_boundary_list.len()  */
    uint32_t last_sor_iterations; /* result of the last solve               */
    uint32_t reserved;
    double last_norm_squared;
} sb_state;

/* ---- construction / destruction ------------------------------------------- */

/* Simulation::try_from(UnfinalizedSimulation) (src/simulation.rs:71-99) incl.
 * SimulationGrid::try_from (src/grid/mod.rs:127-153): classify boundaries, ranges,
 * F/G, RHS, initial residual norm (unless provided).  p/u/v may be NULL (zeros).
 * On SB_BOUNDARY_TOO_THIN no handle is returned; sb_error_cell(NULL, ..) gives
 * the offending cell. */
sb_status sb_create(const sb_params *params, const double *p, const double *u,
                    const double *v, const uint8_t *kind,
                    const sb_boundary_velocity *velocities, size_t n_velocities,
                    sb_sim **out);
void sb_destroy(sb_sim *sim);

/* ---- the hot path ----------------------------------------------------------- */

/* Simulation::run_simulation_tick (src/simulation.rs:324-333) */
sb_status sb_tick(sb_sim *sim, uint32_t *sor_iterations, double *norm_squared);
/* n ticks back to back without reading results back in between (the GUI calls
 * the tick 20x per frame, src/lib.rs:214-219); returns the last tick's pair */
sb_status sb_run_ticks(sb_sim *sim, uint32_t n, uint32_t *sor_iterations, double *norm_squared);

/* One tick on HOST buffers: the call a host-side owner of the fields makes (the reference
 * keeps `pressure`, `u`, `v` in host arrays, src/grid/mod.rs:112-125, and ticks them in
 * place, src/simulation.rs:324-333).  Uploads p, u, v ([rows][ny] f64, this slab's own rows),
 * runs sb_tick, downloads the three fields into p_out, u_out, v_out (may alias the inputs);
 * all copies are asynchronous on the handle's stream with ONE synchronisation at the end, so
 * several handles driven from several host threads overlap their copies (PCIe is full
 * duplex) and their kernels (stroemung_b200/pipeline.py).  Page-locked buffers
 * (sb_host_alloc) make the copies truly asynchronous. */
sb_status sb_tick_host(sb_sim *sim, const double *p_in, const double *u_in, const double *v_in,
                       double *p_out, double *u_out, double *v_out, uint32_t *sor_iterations,
                       double *norm_squared);

/* stage entry points, one per reference function (stage-level parity) */
sb_status sb_set_boundary_u_and_v(sb_sim *sim);        /* src/grid/mod.rs:414-651   */
sb_status sb_calculate_f_and_g(sb_sim *sim);           /* src/simulation.rs:122-202 */
sb_status sb_calculate_rhs(sb_sim *sim);               /* src/simulation.rs:204-214 */
sb_status sb_copy_pressure_to_boundaries(sb_sim *sim); /* src/grid/mod.rs:343-412   */
sb_status sb_calculate_norm_squared(sb_sim *sim, double *norm_squared); /* src/simulation.rs:216-227 */
sb_status sb_solve_sor(sb_sim *sim, uint32_t *sor_iterations, double *norm_squared); /* :239-285 */
sb_status sb_set_u_and_v(sb_sim *sim);                 /* src/simulation.rs:287-322 */
sb_status sb_calculate_pressure_range(sb_sim *sim);    /* src/grid/mod.rs:237-251   */
sb_status sb_calculate_speed_range(sb_sim *sim);       /* src/grid/mod.rs:253-268   */
/* exactly `n` SOR iterations (BC copy + sweep each) with no exit test; norms (may be
 * NULL) receives the n residual norms.  Benchmark / parity helper. */
sb_status sb_sor_sweeps(sb_sim *sim, uint32_t n, double *norms);

/* ---- state access ------------------------------------------------------------ */

/* direct `pub` field access of the reference becomes explicit copies */
sb_status sb_download(sb_sim *sim, sb_field field, void *dst);
sb_status sb_upload(sb_sim *sim, sb_field field, const void *src);
/* pinned-host staging for the end-to-end path: allocate / free page-locked memory */
void *sb_host_alloc(size_t bytes);
void sb_host_free(void *ptr);
sb_status sb_get_state(sb_sim *sim, sb_state *state);
/* scalar parameters only (delt, gamma, reynolds, eps, omega, max_iterations, tau,
 * time, iterations, initial norm, sor_mode, temporal_block); geometry is fixed */
sb_status sb_set_params(sb_sim *sim, const sb_params *params);
/* replace the sparse velocity table (after sb_upload(SB_FIELD_KIND)) */
sb_status sb_set_boundary_velocities(sb_sim *sim, const sb_boundary_velocity *v, size_t n);
/* read it back (the `velocity` payload of the Inflow cells, src/cell.rs:12-16, that
 * `#[derive(Serialize)]` writes with the cell types): up to `capacity` entries into v (may
 * be NULL), the table's length into *n */
sb_status sb_get_boundary_velocities(sb_sim *sim, sb_boundary_velocity *v, size_t capacity,
                                     size_t *n);

/* SimulationGrid::rebuild_boundary_list (src/grid/mod.rs:202-235), called after
 * sb_upload(SB_FIELD_KIND, ..) as src/lib.rs:70 does.  On SB_BOUNDARY_TOO_THIN the
 * previous list stays active (src/grid/mod.rs:232-233). */
sb_status sb_rebuild_boundary_list(sb_sim *sim);
/* sorted_boundary_list (src/grid/mod.rs:64): linear global indices x*ny+y in
 * x-major order and their edge class; returns the list length through *n. */
sb_status sb_boundary_list(sb_sim *sim, uint64_t *index, uint8_t *edge, uint64_t capacity,
                           uint64_t *n);
/* draw_cells (src/lib.rs:38-78): paint the 2x2 block at (x, y) with `kind`,
 * zeroing u, v, p; re-classify; roll back on a too-thin wall.  *applied = 1 if kept. */
sb_status sb_edit_cells(sb_sim *sim, uint64_t x, uint64_t y, uint8_t kind, double bu,
                        double bv, int32_t *applied);

/* device-side presets (src/grid/presets.rs:8-87); fields are zero.  preset:
 * 0 empty, 1 simple_inflow, 2 obstacle (circle at (20, ny/2) r=5), 3 channel with a
 * circle (args: cx, cy, r), 4 backward-facing step (args: step_len, step_top),
 * 5 lid-driven cavity (args: lid_u) */
sb_status sb_create_preset(const sb_params *params, int32_t preset, const double *args,
                           size_t n_args, sb_sim **out);

/* render_simulation (src/visualization.rs:79-105) with color_pressure / color_speed /
 * hsl_to_rgb (:7-70) evaluated on the device from the resident fields and the current
 * pressure_range / speed_range: writes the RGBA8 frame in macroquad's Image layout (row-major,
 * width = nx, pixel (x, y) at byte 4 (y nx + x)) to host memory -- 4 bytes per cell over
 * PCIe instead of downloading p / u / v and the cell types.  A slab handle renders its
 * owned rows: an image of width x_end - x_begin. */
sb_status sb_render_rgba(sb_sim *sim, int32_t color_type, uint8_t *dst);

/* ---- errors -------------------------------------------------------------------- */
/* cell named by the last SB_BOUNDARY_TOO_THIN (global x, y) and its kind;
 * sim may be NULL for a failed sb_create */
sb_status sb_error_cell(const sb_sim *sim, uint64_t xy[2], uint8_t *kind);
const char *sb_last_error_string(void);

/* ---- multi-GPU: row slabs along x, one handle per GPU (one process per GPU) ---------
 * The reference is single-process; this is the build's extension for grids beyond one
 * GPU.  Protocol, every rank:
 *   1. sb_create / sb_create_preset with world > 1 and this rank's [x_begin, x_end)
 *      (>= 10 rows).  Host arrays cover the owned rows only; the velocity table must hold
 *      every Inflow / MovingWall cell of the owned rows AND of 10 rows beyond either end.
 *      The handle is not usable yet.
 *   2. sb_slab_export -> a blob; the host all-gathers the blobs of all ranks (any channel).
 *   3. sb_slab_connect(all blobs, rank order): maps the neighbours' halo rows and every
 *      rank's mailbox (CUDA IPC over NVLink; handles of one process use peer access) and
 *      finishes construction collectively (classification, ranges, F/G, RHS, initial norm).
 * After that sb_tick & co. are collective calls: every rank makes the same calls in the
 * same order.  p halos move inside the red-black pass (P2P stores), u/v halos after the
 * velocity update, norms / ranges / counts through an all-gather in rank order, so all
 * ranks return bit-identical scalars.  A peer that never shows up ends in SB_CUDA_ERROR
 * after ~20 s, not in a hang.  Red-black mode only; sb_edit_cells is single-GPU. */
#define SB_SLAB_BLOB_BYTES 1024
sb_status sb_slab_export(sb_sim *sim, uint8_t blob[SB_SLAB_BLOB_BYTES]);
sb_status sb_slab_connect(sb_sim *sim, const uint8_t *blobs, size_t n_blobs);
/* after sb_upload of p / u / v on any rank (collective): refresh all halo rows */
sb_status sb_slab_sync_halos(sb_sim *sim);

/* ---- cell-level operators (src/math.rs, src/simulation.rs:349-392) --------------
 * Evaluated ON THE DEVICE with the same __device__ functions the kernels use, so
 * the reference's exact known-answer tests can be run against the CUDA path.
 * 3x3 blocks are in the reference's [x][y] order: view[(a, b)] == blk[3*a + b]. */
sb_status sb_du2dx(const double u[9], double delx, double gamma, double *out);    /* math.rs:19  */
sb_status sb_duvdx(const double u[9], const double v[9], double delx, double gamma,
                   double *out);                                                  /* math.rs:53  */
sb_status sb_duvdy(const double u[9], const double v[9], double dely, double gamma,
                   double *out);                                                  /* math.rs:97  */
sb_status sb_dv2dy(const double v[9], double dely, double gamma, double *out);    /* math.rs:136 */
sb_status sb_laplacian(const double e[9], double delx, double dely, double *out); /* math.rs:162 */
sb_status sb_residual(const double p[9], double delx, double dely, double rhs,
                      double *out);                                               /* math.rs:176 */
sb_status sb_calculate_f(const double u[9], const double v[9], double delx, double dely,
                         double delt, double gamma, double reynolds, double *out); /* simulation.rs:349 */
sb_status sb_calculate_g(const double u[9], const double v[9], double delx, double dely,
                         double delt, double gamma, double reynolds, double *out); /* simulation.rs:378 */

/* ---- instrumentation ------------------------------------------------------------ */
/* number of kernels this handle has launched so far (bench.py's gpu_launches) */
uint64_t sb_kernel_launches(const sb_sim *sim);
/* device time (ms) of the SOR solve inside the last sb_tick / sb_solve_sor / sb_sor_sweeps,
 * measured with CUDA events on the handle's stream */
double sb_last_sor_ms(const sb_sim *sim);
/* device time (ms, CUDA events on the handle's stream) of the four stages of the last sb_tick,
 * in the order of run_simulation_tick (/root/reference/src/simulation.rs:324-333):
 * ms[0] set_boundary_u_and_v, ms[1] calculate_f_and_g + calculate_rhs, ms[2] solve_sor,
 * ms[3] set_u_and_v (+ ranges) */
sb_status sb_last_stage_ms(sb_sim *sim, double ms[4]);
/* how the last red-black pass split the grid: tiles on the tile kernel (walls, obstacles,
 * grid ring, slab edges) and work items of the streaming kernel (all-fluid regions) */
sb_status sb_rb_plan(const sb_sim *sim, int32_t *tile_kernel_tiles, int32_t *stream_items);
/* which kernels ran the last red-black solve_sor (simulation.rs:239-285): 0 pass by pass (tile +
 * streaming kernels), 1 one launch on one SM (small grids), 2 / 3 one cooperative launch with
 * the grid resident in the shared memory (2) or the registers + shared memory (3) of `*ctas`
 * SMs (mid-size grids); ctas may be NULL */
int32_t sb_last_sor_path(const sb_sim *sim, int32_t *ctas);
/* per-pass profiling of the dominant kernel: when enabled, every SOR sweep-kernel launch
 * (red-black pass or wavefront sweep) is bracketed by CUDA events on the handle's stream.
 * sb_profile_read returns the durations (ms) recorded since the last read, oldest first;
 * launches that found the solve already finished (no-op passes) are included. */
sb_status sb_profile_enable(sb_sim *sim, int32_t enable);
sb_status sb_profile_read(sb_sim *sim, double *ms, size_t capacity, size_t *n);
/* device-side stopwatch: CUDA events recorded on the handle's stream */
sb_status sb_timer_begin(sb_sim *sim);
sb_status sb_timer_end(sb_sim *sim, double *elapsed_ms); /* synchronises the stream */
/* raw CUDA stream of the handle (cudaStream_t) so a host can bracket it with events */
void *sb_stream(const sb_sim *sim);
const char *sb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* STROEMUNG_B200_H */
