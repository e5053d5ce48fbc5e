/*
 * stroemung_oracle.h -- CPU restatement of the stroemung per-timestep solver.
 *
 * TEST INFRASTRUCTURE ONLY.  This is the parity oracle and the reported CPU
 * baseline for stroemung_b200.  Nothing in the product path (stroemung_b200/)
 * includes, links or calls it; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs do.
 *
 * It restates, in plain C and in the reference's exact f64 evaluation order
 * (no FMA contraction, IEEE division where the source divides), the algorithm
 * of wickedchicken/stroemung 0.1.2:
 *     src/math.rs:19-186        stencil operators
 *     src/simulation.rs:71-333  driver (try_from, F/G, RHS, SOR, norm, tick)
 *     src/grid/mod.rs:167-651   classification, pressure/velocity BCs, ranges
 *     src/grid/presets.rs:8-87  presets
 * The reference itself (Rust) cannot be compiled in this environment, so this
 * port is pinned against the reference's own golden vectors instead
 * (tests/golden/, see tests/test_oracle_golden.py): the 32 exact KATs of
 * math.rs / simulation.rs, all nine `simulation_tick` snapshots, the two
 * (iterations, norm) asserts, `initial_norm_squared` 899.9547140394143 and the
 * boundary-classification cases.
 *
 * Extensions that are NOT in the reference (parity unpinned by reference
 * tests; defined here so the CUDA path has something to be compared with):
 *   - SO_SOR_RED_BLACK: red-black ordering with the fused-multiply-add
 *     arithmetic the performance-mode CUDA kernel uses (bit-identical p).
 *   - tau > 0: adaptive time step (NaSt2D COMP_delt formula).
 *   - SO_KIND_MOVING_WALL: tangential-velocity wall (lid-driven cavity).
 */
#ifndef STROEMUNG_ORACLE_H
#define STROEMUNG_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* cell kinds: src/cell.rs:6-23 (+ one extension) */
enum {
    SO_KIND_FLUID = 0,
    SO_KIND_NOSLIP = 1,
    SO_KIND_OUTFLOW = 2,
    SO_KIND_INFLOW = 3,
    SO_KIND_MOVING_WALL = 4 /* extension: NoSlip wall moving with (bu,bv) */
};

/* edge types: src/grid/mod.rs:19-49 (0 = Option::None) */
enum {
    SO_EDGE_NONE = 0,
    SO_EDGE_N = 1,
    SO_EDGE_NE = 2,
    SO_EDGE_E = 3,
    SO_EDGE_SE = 4,
    SO_EDGE_S = 5,
    SO_EDGE_SW = 6,
    SO_EDGE_W = 7,
    SO_EDGE_NW = 8
};

enum { SO_OK = 0, SO_BOUNDARY_TOO_THIN = 1, SO_BOUNDARY_LIST_INCORRECT = 2, SO_INVALID = 4 };

enum { SO_SOR_REFERENCE_ORDER = 0, SO_SOR_RED_BLACK = 1 };

/* mirrors UnfinalizedSimulation, src/simulation.rs:30-44 (+ extensions) */
typedef struct {
    uint64_t nx, ny;              /* size            */
    double delx, dely;            /* cell_size       */
    double delt, gamma, reynolds;
    double sor_absolute_epsilon, omega, time;
    uint32_t max_iterations, iterations;
    int32_t has_initial_norm;     /* Option<Real>::is_some */
    double initial_norm_squared;
    /* extensions */
    double tau;                   /* <= 0: fixed delt (reference behaviour) */
    int32_t sor_mode;             /* SO_SOR_*         */
    int32_t reserved;
} so_params;

typedef struct so_sim so_sim;

/* Simulation::try_from (src/simulation.rs:71-99).  Arrays are row-major
 * [nx][ny] (y contiguous), caller-owned, copied.  bu/bv: per-cell inflow /
 * moving-wall velocity (ignored for other kinds; may be NULL = zeros).
 * On SO_BOUNDARY_TOO_THIN *out is NULL and err_xy receives the offending cell. */
int so_create(const so_params *prm, const double *p, const double *u, const double *v,
              const uint8_t *kind, const double *bu, const double *bv, so_sim **out,
              uint64_t err_xy[2]);
void so_destroy(so_sim *s);

/* SimulationGrid::rebuild_boundary_list (src/grid/mod.rs:202-235).  On error the
 * previously active list stays in force, as in the reference. */
int so_rebuild_boundary_list(so_sim *s, uint64_t err_xy[2]);

/* stages, one per reference function */
int so_set_boundary_u_and_v(so_sim *s);       /* src/grid/mod.rs:414-651   */
void so_calculate_f_and_g(so_sim *s);         /* src/simulation.rs:122-202 */
void so_calculate_rhs(so_sim *s);             /* src/simulation.rs:204-214 */
int so_copy_pressure_to_boundaries(so_sim *s);/* src/grid/mod.rs:343-412   */
double so_calculate_norm_squared(const so_sim *s); /* src/simulation.rs:216-227 */
int so_solve_sor(so_sim *s, uint32_t *iters, double *norm_squared); /* :239-285 */
void so_set_u_and_v(so_sim *s);               /* src/simulation.rs:287-322 */
void so_calculate_pressure_range(so_sim *s);  /* src/grid/mod.rs:237-251   */
void so_calculate_speed_range(so_sim *s);     /* src/grid/mod.rs:253-268   */
int so_tick(so_sim *s, uint32_t *iters, double *norm_squared); /* src/simulation.rs:324-333 */
/* render_simulation (src/visualization.rs:79-105): RGBA8, ny rows of nx pixels;
 * color_type 0 = Pressure, 1 = Speed */
void so_render_rgba(const so_sim *s, int color_type, uint8_t *rgba);

/* exactly one SOR iteration in the mode of the handle (BC copy + sweep), no norm */
void so_sor_sweep(so_sim *s);

/* field access (pointers into the handle, row-major [nx][ny]) */
double *so_p(so_sim *s);
double *so_u(so_sim *s);
double *so_v(so_sim *s);
double *so_f(so_sim *s);
double *so_g(so_sim *s);
double *so_rhs(so_sim *s);
uint8_t *so_kind(so_sim *s);
double *so_bu(so_sim *s);
double *so_bv(so_sim *s);

typedef struct {
    double time, delt;
    uint32_t iterations;
    int32_t has_initial_norm;
    double initial_norm_squared;
    double pressure_range[2], speed_range[2];
    double fluid_cells;
    uint64_t n_boundary;
} so_state;
void so_get_state(const so_sim *s, so_state *st);
void so_set_params(so_sim *s, const so_params *prm); /* scalar fields only */
void so_clear_initial_norm(so_sim *s);  /* initial_norm_squared = None */

/* boundary list read-back: idx = x*ny+y, edge = SO_EDGE_*; returns count */
uint64_t so_boundary_list(const so_sim *s, uint64_t *idx, uint8_t *edge, uint64_t cap);

/* cell-level operators on 3x3 blocks in the reference's [x][y] order
 * (view[(a,b)] == blk[3*a+b]), src/math.rs and src/simulation.rs:349-392 */
double so_du2dx(const double u[9], double delx, double gamma);
double so_duvdx(const double u[9], const double v[9], double delx, double gamma);
double so_duvdy(const double u[9], const double v[9], double dely, double gamma);
double so_dv2dy(const double v[9], double dely, double gamma);
double so_laplacian(const double e[9], double delx, double dely);
double so_residual(const double p[9], double delx, double dely, double rhs);
double so_calculate_f(const double u[9], const double v[9], double delx, double dely,
                      double delt, double gamma, double reynolds);
double so_calculate_g(const double u[9], const double v[9], double delx, double dely,
                      double delt, double gamma, double reynolds);

/* presets (src/grid/presets.rs): fill kind/bu/bv for an [nx][ny] grid */
void so_preset_empty(uint64_t nx, uint64_t ny, uint8_t *kind, double *bu, double *bv);
void so_preset_simple_inflow(uint64_t nx, uint64_t ny, uint8_t *kind, double *bu, double *bv);
void so_preset_obstacle(uint64_t nx, uint64_t ny, uint8_t *kind, double *bu, double *bv);
/* generalised draw_circle (same integer rasteriser, src/grid/presets.rs:42-62) */
void so_draw_circle(uint64_t nx, uint64_t ny, uint8_t *kind, uint64_t cx, uint64_t cy,
                    double radius);

#ifdef __cplusplus
}
#endif
#endif
