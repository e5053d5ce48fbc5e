/*
 * stroemung_oracle.c -- CPU restatement of the stroemung per-timestep solver.
 * TEST INFRASTRUCTURE ONLY (see stroemung_oracle.h).  Build with
 *     gcc -O3 -ffp-contract=off -fno-fast-math
 * so that every expression below rounds exactly as the Rust source does
 * (Rust never contracts a*b+c; f64::powi(2) is x*x).
 *
 * Index convention (src/types.rs:8-19): arrays are [nx][ny] row-major, index
 * (x, y), y contiguous; (0,0) is the upper-left, "north" is y-1
 * (src/grid/mod.rs:168-176).
 */
#include "stroemung_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    uint64_t idx; /* x*ny + y */
    uint8_t edge;
} so_bentry;

typedef struct {
    uint64_t idx;
    uint8_t has_u, has_v;
    double u, v;
} so_restore;

struct so_sim {
    so_params prm;
    uint64_t nx, ny, n;
    double *p, *u, *v, *f, *g, *rhs, *bu, *bv;
    uint8_t *kind;
    /* BoundaryList (src/grid/mod.rs:61-69) */
    so_bentry *blist;
    uint64_t n_boundary;
    double fluid_cells;
    so_restore *restore;
    uint64_t n_restore, cap_restore;
    double pressure_range[2], speed_range[2];
    int has_initial_norm;
    double initial_norm_squared;
    double time;
    uint32_t iterations;
};

#define IDX(s, x, y) ((uint64_t)(x) * (s)->ny + (uint64_t)(y))

/* ------------------------------------------------------------------ */
/* src/math.rs                                                         */
/* ------------------------------------------------------------------ */
#define B(blk, a, b) ((blk)[3 * (a) + (b)])

/* src/math.rs:19-33 */
double so_du2dx(const double u[9], double delx, double gamma) {
    double u_i_m1 = B(u, 0, 1), u_i = B(u, 1, 1), u_i_p1 = B(u, 2, 1);
    double a = u_i + u_i_p1, b = u_i_m1 + u_i;
    double inner_left1 = a * a;
    double inner_right1 = b * b;
    double left_side = inner_left1 - inner_right1;
    double inner_left2 = fabs(u_i + u_i_p1) * (u_i - u_i_p1);
    double inner_right2 = fabs(u_i_m1 + u_i) * (u_i_m1 - u_i);
    return (left_side + (gamma * (inner_left2 - inner_right2))) / (4.0 * delx);
}

/* src/math.rs:53-77 */
double so_duvdx(const double u[9], const double v[9], double delx, double gamma) {
    double u_i_j = B(u, 1, 1), u_i_j_p1 = B(u, 1, 2), u_i_m1_j = B(u, 0, 1),
           u_i_m1_j_p1 = B(u, 0, 2);
    double v_i_j = B(v, 1, 1), v_i_p1_j = B(v, 2, 1), v_i_m1_j = B(v, 0, 1);
    double inner_left1 = (u_i_j + u_i_j_p1) * (v_i_j + v_i_p1_j);
    double inner_right1 = (u_i_m1_j + u_i_m1_j_p1) * (v_i_m1_j + v_i_j);
    double left_side = inner_left1 - inner_right1;
    double inner_left2 = fabs(u_i_j + u_i_j_p1) * (v_i_j - v_i_p1_j);
    double inner_right2 = fabs(u_i_m1_j + u_i_m1_j_p1) * (v_i_m1_j - v_i_j);
    return (left_side + (gamma * (inner_left2 - inner_right2))) / (4.0 * delx);
}

/* src/math.rs:97-120 */
double so_duvdy(const double u[9], const double v[9], double dely, double gamma) {
    double u_i_j = B(u, 1, 1), u_i_j_m1 = B(u, 1, 0), u_i_j_p1 = B(u, 1, 2);
    double v_i_j = B(v, 1, 1), v_i_j_m1 = B(v, 1, 0), v_i_p1_j = B(v, 2, 1),
           v_i_p1_j_m1 = B(v, 2, 0);
    double inner_left1 = (v_i_j + v_i_p1_j) * (u_i_j + u_i_j_p1);
    double inner_right1 = (v_i_j_m1 + v_i_p1_j_m1) * (u_i_j_m1 + u_i_j);
    double left_side = inner_left1 - inner_right1;
    double inner_left2 = fabs(v_i_j + v_i_p1_j) * (u_i_j - u_i_j_p1);
    double inner_right2 = fabs(v_i_j_m1 + v_i_p1_j_m1) * (u_i_j_m1 - u_i_j);
    return (left_side + (gamma * (inner_left2 - inner_right2))) / (4.0 * dely);
}

/* src/math.rs:136-150 */
double so_dv2dy(const double v[9], double dely, double gamma) {
    double v_i_j = B(v, 1, 1), v_i_j_p1 = B(v, 1, 2), v_i_j_m1 = B(v, 1, 0);
    double a = v_i_j + v_i_j_p1, b = v_i_j_m1 + v_i_j;
    double inner_left1 = a * a;
    double inner_right1 = b * b;
    double left_side = inner_left1 - inner_right1;
    double inner_left2 = fabs(v_i_j + v_i_j_p1) * (v_i_j - v_i_j_p1);
    double inner_right2 = fabs(v_i_j_m1 + v_i_j) * (v_i_j_m1 - v_i_j);
    return (left_side + (gamma * (inner_left2 - inner_right2))) / (4.0 * dely);
}

/* src/math.rs:162-174 */
double so_laplacian(const double e[9], double delx, double dely) {
    double e_i_j = B(e, 1, 1), e_i_j_m1 = B(e, 1, 0), e_i_j_p1 = B(e, 1, 2),
           e_i_m1_j = B(e, 0, 1), e_i_p1_j = B(e, 2, 1);
    double d2edx2 = ((e_i_p1_j - (2. * e_i_j)) + e_i_m1_j) / (delx * delx);
    double d2edy2 = ((e_i_j_p1 - (2. * e_i_j)) + e_i_j_m1) / (dely * dely);
    return d2edx2 + d2edy2;
}

/* src/math.rs:176-186 */
double so_residual(const double p[9], double delx, double dely, double rhs) {
    double p_i_p1_j = B(p, 2, 1), p_i_j = B(p, 1, 1), p_i_m1_j = B(p, 0, 1),
           p_i_j_p1 = B(p, 1, 2), p_i_j_m1 = B(p, 1, 0);
    double part1 = ((p_i_p1_j - p_i_j) - (p_i_j - p_i_m1_j)) / (delx * delx);
    double part2 = ((p_i_j_p1 - p_i_j) - (p_i_j - p_i_j_m1)) / (dely * dely);
    return (part1 + part2) - rhs;
}

/* src/simulation.rs:349-363 */
double so_calculate_f(const double u[9], const double v[9], double delx, double dely,
                      double delt, double gamma, double reynolds) {
    return B(u, 1, 1) + (delt * (((so_laplacian(u, delx, dely) / reynolds) -
                                  so_du2dx(u, delx, gamma)) -
                                 so_duvdy(u, v, dely, gamma)));
}

/* src/simulation.rs:378-392 */
double so_calculate_g(const double u[9], const double v[9], double delx, double dely,
                      double delt, double gamma, double reynolds) {
    return B(v, 1, 1) + (delt * (((so_laplacian(v, delx, dely) / reynolds) -
                                  so_duvdx(u, v, delx, gamma)) -
                                 so_dv2dy(v, dely, gamma)));
}

static inline void gather3x3(const double *a, uint64_t ny, uint64_t x, uint64_t y,
                             double out[9]) {
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) out[3 * i + j] = a[(x - 1 + i) * ny + (y - 1 + j)];
}

/* ------------------------------------------------------------------ */
/* classification: src/grid/mod.rs:167-332                             */
/* ------------------------------------------------------------------ */
static int is_fluid(const so_sim *s, int64_t x, int64_t y) {
    if (x < 0 || y < 0 || x >= (int64_t)s->nx || y >= (int64_t)s->ny) return 0;
    return s->kind[IDX(s, x, y)] == SO_KIND_FLUID;
}

/* calculate_edges (src/grid/mod.rs:270-332): returns -1 for BoundaryTooThin */
static int calculate_edge(const so_sim *s, uint64_t x, uint64_t y) {
    int left = is_fluid(s, (int64_t)x - 1, (int64_t)y);
    int right = is_fluid(s, (int64_t)x + 1, (int64_t)y);
    int up = is_fluid(s, (int64_t)x, (int64_t)y - 1);
    int down = is_fluid(s, (int64_t)x, (int64_t)y + 1);
    int m = (left << 3) | (right << 2) | (up << 1) | down;
    switch (m) {
    case 0x0: return SO_EDGE_NONE;
    case 0x8: return SO_EDGE_W;
    case 0xA: return SO_EDGE_NW;
    case 0x2: return SO_EDGE_N;
    case 0x6: return SO_EDGE_NE;
    case 0x4: return SO_EDGE_E;
    case 0x5: return SO_EDGE_SE;
    case 0x1: return SO_EDGE_S;
    case 0x9: return SO_EDGE_SW;
    default: return -1;
    }
}

/* rebuild_boundary_list (src/grid/mod.rs:202-235).  The BTreeSet of
 * BoundaryIndex(x, y) iterates x-major == increasing linear index. */
int so_rebuild_boundary_list(so_sim *s, uint64_t err_xy[2]) {
    uint64_t fluid = 0, nb = 0;
    for (uint64_t i = 0; i < s->n; i++) {
        if (s->kind[i] == SO_KIND_FLUID) fluid++;
        else nb++;
    }
    so_bentry *list = (so_bentry *)malloc((nb ? nb : 1) * sizeof(so_bentry));
    uint64_t k = 0;
    for (uint64_t x = 0; x < s->nx; x++) {
        for (uint64_t y = 0; y < s->ny; y++) {
            if (s->kind[IDX(s, x, y)] == SO_KIND_FLUID) continue;
            int e = calculate_edge(s, x, y);
            if (e < 0) {
                /* `result?` returns before sorted_boundary_list / fluid_cells
                 * are assigned (src/grid/mod.rs:232-233) */
                if (err_xy) { err_xy[0] = x; err_xy[1] = y; }
                free(list);
                /* u_v_restore was reset before the scan (src/grid/mod.rs:205) */
                s->n_restore = 0;
                return SO_BOUNDARY_TOO_THIN;
            }
            list[k].idx = IDX(s, x, y);
            list[k].edge = (uint8_t)e;
            k++;
        }
    }
    free(s->blist);
    s->blist = list;
    s->n_boundary = nb;
    s->fluid_cells = (double)fluid;
    s->n_restore = 0;
    return SO_OK;
}

/* ------------------------------------------------------------------ */
/* ranges: src/grid/mod.rs:237-268                                     */
/* ------------------------------------------------------------------ */
void so_calculate_pressure_range(so_sim *s) {
    double mn = DBL_MAX, mx = 0.0;
    for (uint64_t i = 0; i < s->n; i++) {
        if (s->kind[i] == SO_KIND_FLUID) {
            mn = fmin(mn, s->p[i]);
            mx = fmax(mx, s->p[i]);
        }
    }
    s->pressure_range[0] = mn;
    s->pressure_range[1] = mx;
}

void so_calculate_speed_range(so_sim *s) {
    double mn = DBL_MAX, mx = 0.0;
    for (uint64_t i = 0; i < s->n; i++) {
        if (s->kind[i] == SO_KIND_FLUID) {
            double sq = (s->u[i] * s->u[i]) + (s->v[i] * s->v[i]);
            mn = fmin(mn, sq);
            mx = fmax(mx, sq);
        }
    }
    s->speed_range[0] = sqrt(mn);
    s->speed_range[1] = sqrt(mx);
}

/* ------------------------------------------------------------------ */
/* colour mapping: src/visualization.rs (N4)                           */
/* ------------------------------------------------------------------ */
/* Rust `as u8` on an f32: saturating, NaN -> 0 */
static uint8_t sat_u8(float v) {
    if (!(v > 0.0f)) return 0;
    if (v >= 255.0f) return 255;
    return (uint8_t)v;
}

/* hsl_to_rgb (src/visualization.rs:7-27), all f32; `%` on floats is fmodf.  The result goes
 * through macroquad 0.4.13's `From<Color> for [u8; 4]` = `(c * 255.) as u8` per channel
 * (crate not vendored: Cargo.lock:284-285; Image::set_pixel stores that at y*width + x). */
static void hsl_to_rgba8(float hue, float saturation, float lightness, uint8_t out[4]) {
    const float c = (1.0f - fabsf(2.0f * lightness - 1.0f)) * saturation;
    const float q = hue / 60.0f;
    const float x = c * (1.0f - fabsf(fmodf(q, 2.0f) - 1.0f));
    const float m = lightness - c / 2.0f;
    float r, g, b;
    if (hue < 60.0f) { r = c; g = x; b = 0.0f; }
    else if (hue < 120.0f) { r = x; g = c; b = 0.0f; }
    else if (hue < 180.0f) { r = 0.0f; g = c; b = x; }
    else if (hue < 240.0f) { r = 0.0f; g = x; b = c; }
    else if (hue < 300.0f) { r = x; g = 0.0f; b = c; }
    else { r = c; g = 0.0f; b = x; }
    const float rr = r + m, gg = g + m, bb = b + m;
    out[0] = sat_u8(rr * 255.0f);
    out[1] = sat_u8(gg * 255.0f);
    out[2] = sat_u8(bb * 255.0f);
    out[3] = sat_u8(1.0f * 255.0f);
}

/* render_simulation (src/visualization.rs:79-105) with color_pressure (:49-70) and
 * color_speed (:29-47); color_type 0 = Pressure, 1 = Speed (ColorType, :72-77).
 * rgba: ny rows of nx pixels (macroquad Image layout). */
void so_render_rgba(const so_sim *s, int color_type, uint8_t *rgba) {
    for (uint64_t x = 0; x < s->nx; x++) {
        for (uint64_t y = 0; y < s->ny; y++) {
            uint64_t i = IDX(s, x, y);
            uint8_t *px = rgba + 4 * (y * s->nx + x);
            if (s->kind[i] != SO_KIND_FLUID) {
                /* Color::new(0.5, 0.0, 0.0, 1.0) / (0.5, 0.5, 0.5, 1.0) */
                px[0] = sat_u8(0.5f * 255.0f);
                px[1] = px[2] = color_type ? sat_u8(0.5f * 255.0f) : 0;
                px[3] = 255;
                continue;
            }
            double q, lo, hi;
            if (color_type) {
                q = sqrt((s->u[i] * s->u[i]) + (s->v[i] * s->v[i]));
                lo = s->speed_range[0]; hi = s->speed_range[1];
            } else {
                q = s->p[i];
                lo = s->pressure_range[0]; hi = s->pressure_range[1];
            }
            float hue = (float)(240.0 - (((q - lo) * 240.0) / (hi - lo)));
            hsl_to_rgba8(hue, 1.0f, 0.5f, px);
        }
    }
}

/* ------------------------------------------------------------------ */
/* pressure BC: src/grid/mod.rs:343-412                                */
/* ------------------------------------------------------------------ */
int so_copy_pressure_to_boundaries(so_sim *s) {
    const uint64_t ny = s->ny;
    double *p = s->p;
    for (uint64_t k = 0; k < s->n_boundary; k++) {
        uint64_t b = s->blist[k].idx;
        int e = s->blist[k].edge;
        if (e == SO_EDGE_NONE) continue;
        if (s->kind[b] == SO_KIND_FLUID) return SO_BOUNDARY_LIST_INCORRECT;
        uint64_t n = b - 1, so = b + 1, ea = b + ny, w = b - ny;
        switch (e) {
        case SO_EDGE_N: p[b] = p[n]; break;
        case SO_EDGE_NE: p[b] = (p[n] + p[ea]) / 2.0; break;
        case SO_EDGE_E: p[b] = p[ea]; break;
        case SO_EDGE_SE: p[b] = (p[so] + p[ea]) / 2.0; break;
        case SO_EDGE_S: p[b] = p[so]; break;
        case SO_EDGE_SW: p[b] = (p[so] + p[w]) / 2.0; break;
        case SO_EDGE_W: p[b] = p[w]; break;
        case SO_EDGE_NW: p[b] = (p[n] + p[w]) / 2.0; break;
        }
    }
    return SO_OK;
}

/* ------------------------------------------------------------------ */
/* velocity BC: src/grid/mod.rs:414-651 (sequential, in place)         */
/* ------------------------------------------------------------------ */
static void push_restore(so_sim *s, uint64_t idx, int has_u, double u, int has_v, double v) {
    if (s->n_restore == s->cap_restore) {
        s->cap_restore = s->cap_restore ? 2 * s->cap_restore : 64;
        s->restore = (so_restore *)realloc(s->restore, s->cap_restore * sizeof(so_restore));
    }
    so_restore *r = &s->restore[s->n_restore++];
    r->idx = idx;
    r->has_u = (uint8_t)has_u;
    r->has_v = (uint8_t)has_v;
    r->u = u;
    r->v = v;
}

int so_set_boundary_u_and_v(so_sim *s) {
    const uint64_t ny = s->ny;
    double *u = s->u, *v = s->v;
    s->n_restore = 0;
    for (uint64_t k = 0; k < s->n_boundary; k++) {
        uint64_t b = s->blist[k].idx;
        int e = s->blist[k].edge;
        if (e == SO_EDGE_NONE) {
            push_restore(s, b, 1, u[b], 1, v[b]);
            continue;
        }
        uint64_t n = b - 1, so = b + 1, ea = b + ny, w = b - ny;
        int kind = s->kind[b];
        if (kind == SO_KIND_NOSLIP || kind == SO_KIND_INFLOW) {
            /* src/grid/mod.rs:437-488 (NoSlip: 0,0) and :537-586 (Inflow) */
            double boundary_u = kind == SO_KIND_INFLOW ? s->bu[b] : 0.0;
            double boundary_v = kind == SO_KIND_INFLOW ? s->bv[b] : 0.0;
            switch (e) {
            case SO_EDGE_N: u[b] = -u[n]; v[n] = boundary_v; break;
            case SO_EDGE_NE: u[b] = boundary_u; v[n] = boundary_v; v[b] = -v[ea]; break;
            case SO_EDGE_E: u[b] = boundary_u; v[b] = -v[ea]; break;
            case SO_EDGE_SE: u[b] = boundary_u; v[b] = boundary_v; break;
            case SO_EDGE_S: u[b] = -u[so]; v[b] = boundary_v; break;
            case SO_EDGE_SW: u[w] = boundary_u; u[b] = -u[so]; v[b] = boundary_v; break;
            case SO_EDGE_W: u[w] = boundary_u; v[b] = -v[w]; break;
            case SO_EDGE_NW:
                u[w] = boundary_u; u[b] = -u[n]; v[n] = boundary_v; v[b] = -v[w];
                break;
            }
        } else if (kind == SO_KIND_MOVING_WALL) {
            /* extension (not in the reference): wall moving with (bu, bv);
             * tangential ghost value is the reflection about the wall velocity */
            double boundary_u = s->bu[b], boundary_v = s->bv[b];
            switch (e) {
            case SO_EDGE_N: u[b] = (2.0 * boundary_u) - u[n]; v[n] = boundary_v; break;
            case SO_EDGE_NE:
                u[b] = boundary_u; v[n] = boundary_v; v[b] = (2.0 * boundary_v) - v[ea];
                break;
            case SO_EDGE_E: u[b] = boundary_u; v[b] = (2.0 * boundary_v) - v[ea]; break;
            case SO_EDGE_SE: u[b] = boundary_u; v[b] = boundary_v; break;
            case SO_EDGE_S: u[b] = (2.0 * boundary_u) - u[so]; v[b] = boundary_v; break;
            case SO_EDGE_SW:
                u[w] = boundary_u; u[b] = (2.0 * boundary_u) - u[so]; v[b] = boundary_v;
                break;
            case SO_EDGE_W: u[w] = boundary_u; v[b] = (2.0 * boundary_v) - v[w]; break;
            case SO_EDGE_NW:
                u[w] = boundary_u; u[b] = (2.0 * boundary_u) - u[n]; v[n] = boundary_v;
                v[b] = (2.0 * boundary_v) - v[w];
                break;
            }
        } else if (kind == SO_KIND_OUTFLOW) {
            /* src/grid/mod.rs:489-536 */
            switch (e) {
            case SO_EDGE_N: u[b] = u[n]; v[b] = v[n]; break;
            case SO_EDGE_NE: u[b] = u[n]; v[b] = v[ea]; break;
            case SO_EDGE_E: u[b] = u[ea]; v[b] = v[ea]; break;
            case SO_EDGE_SE: u[b] = u[ea]; v[b] = v[so]; break;
            case SO_EDGE_S: u[b] = u[so]; v[b] = v[so]; break;
            case SO_EDGE_SW: u[b] = u[w]; v[b] = v[so]; break;
            case SO_EDGE_W: u[b] = u[w]; v[b] = v[w]; break;
            case SO_EDGE_NW: u[b] = u[n]; v[b] = v[w]; break;
            }
        } else {
            return SO_BOUNDARY_LIST_INCORRECT; /* src/grid/mod.rs:587-592 */
        }
        push_restore(s, b, 1, u[b], 1, v[b]); /* :595-599 */
        /* second record, keyed by the BOUNDARY cell's index (:602-648) */
        switch (e) {
        case SO_EDGE_N:
        case SO_EDGE_NE: push_restore(s, b, 0, 0.0, 1, v[n]); break;
        case SO_EDGE_SW:
        case SO_EDGE_W: push_restore(s, b, 1, u[w], 0, 0.0); break;
        case SO_EDGE_NW: push_restore(s, b, 1, u[w], 1, v[n]); break;
        default: break;
        }
    }
    return SO_OK;
}

/* ------------------------------------------------------------------ */
/* F, G, RHS: src/simulation.rs:122-214                                */
/* ------------------------------------------------------------------ */
/* EXTENSION (performance mode, SO_SOR_RED_BLACK only; not in the reference): F, G and RHS
 * with the kernels' arithmetic -- the formulas of src/math.rs:19-174 and
 * src/simulation.rs:349-392, 204-214 with every division replaced by a multiplication with a
 * reciprocal computed once per run and a*b+c taken fused where the kernel does (explicit
 * fma(); this file is built with -ffp-contract=off, so nothing else contracts).  The GPU
 * kernel (stages.cu, fgr_cell_fast) evaluates the same tree: bit-identical.  Against the
 * strict formulas the values differ in the last bits only; like the red-black ordering this
 * is covered by the performance-mode tolerance (converged fields within the SOR eps). */
typedef struct {
    double rdx2, rdy2, r4dx, r4dy, rre, rdx, rdy, rdt, gamma, delt;
} fg_fast_consts;

static fg_fast_consts fg_fast_constants(const so_sim *s) {
    fg_fast_consts k;
    const double dx = s->prm.delx, dy = s->prm.dely;
    k.rdx2 = 1.0 / (dx * dx);
    k.rdy2 = 1.0 / (dy * dy);
    k.r4dx = 1.0 / (4.0 * dx);
    k.r4dy = 1.0 / (4.0 * dy);
    k.rre = 1.0 / s->prm.reynolds;
    k.rdx = 1.0 / dx;
    k.rdy = 1.0 / dy;
    k.rdt = 1.0 / s->prm.delt;
    k.gamma = s->prm.gamma;
    k.delt = s->prm.delt;
    return k;
}

/* laplacian: rdx2*((e - 2c) + w) + rdy2*((s - 2c) + n) */
static inline double lap_fast(const fg_fast_consts *k, double c, double n, double s_, double w,
                              double e) {
    const double tc = 2.0 * c;
    return fma(k->rdx2, (e - tc) + w, k->rdy2 * ((s_ - tc) + n));
}

/* numerator of a donor-cell term: (a*pa - b*pb) + gamma*(|a|*da - |b|*db) */
static inline double donor_fast(const fg_fast_consts *k, double a, double pa, double b,
                                double pb, double da, double db) {
    const double left = fma(a, pa, -(b * pb));
    const double d = fma(fabs(a), da, -(fabs(b) * db));
    return fma(k->gamma, d, left);
}

static inline void fg_cell_fast(const fg_fast_consts *k, const double u[9], const double v[9],
                                double *f, double *g) {
    const double u_c = B(u, 1, 1), u_n = B(u, 1, 0), u_s = B(u, 1, 2), u_w = B(u, 0, 1),
                 u_e = B(u, 2, 1), u_sw = B(u, 0, 2);
    const double v_c = B(v, 1, 1), v_n = B(v, 1, 0), v_s = B(v, 1, 2), v_w = B(v, 0, 1),
                 v_e = B(v, 2, 1), v_ne = B(v, 2, 0);
    /* du2dx: a = u_c + u_e, b = u_w + u_c */
    const double a1 = u_c + u_e, b1 = u_w + u_c;
    const double x_f = donor_fast(k, a1, a1, b1, b1, u_c - u_e, u_w - u_c);
    /* duvdy: a = v_c + v_e, b = v_n + v_ne */
    const double a2 = v_c + v_e, b2 = v_n + v_ne;
    const double y_f = donor_fast(k, a2, u_c + u_s, b2, u_n + u_c, u_c - u_s, u_n - u_c);
    /* duvdx: a = u_c + u_s, b = u_w + u_sw */
    const double a3 = u_c + u_s, b3 = u_w + u_sw;
    const double x_g = donor_fast(k, a3, v_c + v_e, b3, v_w + v_c, v_c - v_e, v_w - v_c);
    /* dv2dy: a = v_c + v_s, b = v_n + v_c */
    const double a4 = v_c + v_s, b4 = v_n + v_c;
    const double y_g = donor_fast(k, a4, a4, b4, b4, v_c - v_s, v_n - v_c);
    const double lu = lap_fast(k, u_c, u_n, u_s, u_w, u_e);
    const double lv = lap_fast(k, v_c, v_n, v_s, v_w, v_e);
    /* F = u + dt*(lap(u)/Re - du2dx - duvdy), the divisions by 4dx / 4dy folded in */
    *f = fma(k->delt, fma(-k->r4dy, y_f, fma(-k->r4dx, x_f, k->rre * lu)), u_c);
    *g = fma(k->delt, fma(-k->r4dy, y_g, fma(-k->r4dx, x_g, k->rre * lv)), v_c);
}

void so_calculate_f_and_g(so_sim *s) {
    const uint64_t nx = s->nx, ny = s->ny;
    const double delx = s->prm.delx, dely = s->prm.dely, delt = s->prm.delt,
                 gamma = s->prm.gamma, re = s->prm.reynolds;
    const int fast = s->prm.sor_mode == SO_SOR_RED_BLACK;
    const fg_fast_consts kf = fg_fast_constants(s);
    if (nx >= 3 && ny >= 3) {
        for (uint64_t x = 1; x + 1 < nx; x++) {
            for (uint64_t y = 1; y + 1 < ny; y++) {
                double ub[9], vb[9];
                gather3x3(s->u, ny, x, y, ub);
                gather3x3(s->v, ny, x, y, vb);
                if (fast) {
                    fg_cell_fast(&kf, ub, vb, &s->f[IDX(s, x, y)], &s->g[IDX(s, x, y)]);
                    continue;
                }
                s->f[IDX(s, x, y)] = so_calculate_f(ub, vb, delx, dely, delt, gamma, re);
                s->g[IDX(s, x, y)] = so_calculate_g(ub, vb, delx, dely, delt, gamma, re);
            }
        }
    }
    /* :167-201 */
    for (uint64_t k = 0; k < s->n_boundary; k++) {
        uint64_t b = s->blist[k].idx;
        s->f[b] = s->u[b];
        s->g[b] = s->v[b];
        uint64_t n = b - 1, w = b - ny;
        switch (s->blist[k].edge) {
        case SO_EDGE_N: s->g[n] = s->v[n]; break;
        case SO_EDGE_NW: s->f[w] = s->u[w]; s->g[n] = s->v[n]; break;
        case SO_EDGE_W: s->f[w] = s->u[w]; break;
        case SO_EDGE_SW: s->f[w] = s->u[w]; break;
        case SO_EDGE_NE: s->g[n] = s->v[n]; break;
        default: break;
        }
    }
}

void so_calculate_rhs(so_sim *s) {
    const uint64_t nx = s->nx, ny = s->ny;
    const double delx = s->prm.delx, dely = s->prm.dely, delt = s->prm.delt;
    if (s->prm.sor_mode == SO_SOR_RED_BLACK) { /* extension: performance-mode arithmetic */
        const fg_fast_consts k = fg_fast_constants(s);
        for (uint64_t x = 1; x < nx; x++) {
            for (uint64_t y = 1; y < ny; y++) {
                uint64_t c = IDX(s, x, y);
                s->rhs[c] = k.rdt * fma(k.rdx, s->f[c] - s->f[c - ny],
                                        k.rdy * (s->g[c] - s->g[c - 1]));
            }
        }
        return;
    }
    for (uint64_t x = 1; x < nx; x++) {
        for (uint64_t y = 1; y < ny; y++) {
            uint64_t c = IDX(s, x, y);
            s->rhs[c] = (((s->f[c] - s->f[c - ny]) / delx) + ((s->g[c] - s->g[c - 1]) / dely)) /
                        delt;
        }
    }
}

/* ------------------------------------------------------------------ */
/* residual norm: src/simulation.rs:216-237                            */
/* ------------------------------------------------------------------ */
static double norm_reference(const so_sim *s) {
    const uint64_t nx = s->nx, ny = s->ny;
    const double delx = s->prm.delx, dely = s->prm.dely;
    double acc = 0.0;
    if (nx >= 3 && ny >= 3) {
        for (uint64_t x = 1; x + 1 < nx; x++) {
            for (uint64_t y = 1; y + 1 < ny; y++) {
                double pb[9];
                gather3x3(s->p, ny, x, y, pb);
                double r = so_residual(pb, delx, dely, s->rhs[IDX(s, x, y)]);
                acc = acc + (r * r);
            }
        }
    }
    return acc / s->fluid_cells;
}

/* performance-mode constants shared by the red-black sweep and its norm */
typedef struct {
    double rdx2, rdy2, diag, mid, omw;
} rb_consts;

static rb_consts rb_constants(const so_sim *s) {
    rb_consts c;
    double dx2 = s->prm.delx * s->prm.delx, dy2 = s->prm.dely * s->prm.dely;
    c.rdx2 = 1.0 / dx2;
    c.rdy2 = 1.0 / dy2;
    c.diag = (2.0 * c.rdx2) + (2.0 * c.rdy2);
    c.mid = s->prm.omega / ((2.0 / dx2) + (2.0 / dy2));
    c.omw = 1.0 - s->prm.omega;
    return c;
}

/* t = rdx2*(pE+pW) + rdy2*(pS+pN) - rhs with the kernel's FMA association */
static inline double rb_t(const so_sim *s, const rb_consts *c, uint64_t i) {
    const uint64_t ny = s->ny;
    const double *p = s->p;
    return fma(c->rdx2, p[i + ny] + p[i - ny], fma(c->rdy2, p[i + 1] + p[i - 1], -s->rhs[i]));
}

static double norm_red_black(const so_sim *s) {
    const uint64_t nx = s->nx, ny = s->ny;
    rb_consts c = rb_constants(s);
    double acc = 0.0;
    if (nx >= 3 && ny >= 3) {
        for (uint64_t x = 1; x + 1 < nx; x++) {
            for (uint64_t y = 1; y + 1 < ny; y++) {
                uint64_t i = IDX(s, x, y);
                double r = fma(-c.diag, s->p[i], rb_t(s, &c, i));
                acc = acc + (r * r);
            }
        }
    }
    return acc / s->fluid_cells;
}

double so_calculate_norm_squared(const so_sim *s) {
    return s->prm.sor_mode == SO_SOR_RED_BLACK ? norm_red_black(s) : norm_reference(s);
}

static double get_initial_norm_squared(so_sim *s) {
    if (s->has_initial_norm) return s->initial_norm_squared;
    double norm = so_calculate_norm_squared(s);
    s->has_initial_norm = 1;
    s->initial_norm_squared = norm;
    return norm;
}

/* ------------------------------------------------------------------ */
/* SOR: src/simulation.rs:239-285                                      */
/* ------------------------------------------------------------------ */
static void sweep_reference(so_sim *s) {
    const uint64_t nx = s->nx, ny = s->ny;
    double delx2 = s->prm.delx * s->prm.delx;
    double dely2 = s->prm.dely * s->prm.dely;
    double one_minus_w = 1.0 - s->prm.omega;
    double middle = s->prm.omega / ((2.0 / delx2) + (2.0 / dely2));
    double *p = s->p;
    if (nx < 3 || ny < 3) return;
    for (uint64_t x = 1; x + 1 < nx; x++) {
        for (uint64_t y = 1; y + 1 < ny; y++) {
            uint64_t c = IDX(s, x, y);
            if (s->kind[c] != SO_KIND_FLUID) continue;
            double p_i_j = p[c], p_i_m1_j = p[c - ny], p_i_p1_j = p[c + ny],
                   p_i_j_m1 = p[c - 1], p_i_j_p1 = p[c + 1];
            double rhs = s->rhs[c];
            p[c] = (one_minus_w * p_i_j) +
                   middle * ((((p_i_p1_j + p_i_m1_j) / delx2) +
                              ((p_i_j_p1 + p_i_j_m1) / dely2)) -
                             rhs);
        }
    }
}

/* extension: red ((x+y) even) then black half-sweep, kernel arithmetic */
static void sweep_red_black(so_sim *s) {
    const uint64_t nx = s->nx, ny = s->ny;
    rb_consts c = rb_constants(s);
    double *p = s->p;
    if (nx < 3 || ny < 3) return;
    for (int colour = 0; colour < 2; colour++) {
        for (uint64_t x = 1; x + 1 < nx; x++) {
            for (uint64_t y = 1; y + 1 < ny; y++) {
                if (((x + y) & 1) != (uint64_t)colour) continue;
                uint64_t i = IDX(s, x, y);
                if (s->kind[i] != SO_KIND_FLUID) continue;
                double t = rb_t(s, &c, i);
                p[i] = fma(c.mid, t, c.omw * p[i]);
            }
        }
    }
}

void so_sor_sweep(so_sim *s) {
    so_copy_pressure_to_boundaries(s);
    if (s->prm.sor_mode == SO_SOR_RED_BLACK) sweep_red_black(s);
    else sweep_reference(s);
}

int so_solve_sor(so_sim *s, uint32_t *iters, double *norm_out) {
    double epsilon_squared = s->prm.sor_absolute_epsilon * s->prm.sor_absolute_epsilon;
    double norm_squared = 0.0;
    for (uint32_t i = 0; i < s->prm.max_iterations; i++) {
        int rc = so_copy_pressure_to_boundaries(s);
        if (rc) return rc;
        if (s->prm.sor_mode == SO_SOR_RED_BLACK) sweep_red_black(s);
        else sweep_reference(s);
        double initial_norm_squared = get_initial_norm_squared(s);
        norm_squared = so_calculate_norm_squared(s);
        if ((norm_squared < initial_norm_squared) || (norm_squared < epsilon_squared)) {
            *iters = i + 1;
            *norm_out = norm_squared;
            return SO_OK;
        }
    }
    so_calculate_pressure_range(s);
    *iters = s->prm.max_iterations;
    *norm_out = norm_squared;
    return SO_OK;
}

/* ------------------------------------------------------------------ */
/* velocity update: src/simulation.rs:287-322                          */
/* ------------------------------------------------------------------ */
void so_set_u_and_v(so_sim *s) {
    const uint64_t nx = s->nx, ny = s->ny;
    const double delx = s->prm.delx, dely = s->prm.dely, delt = s->prm.delt;
    if (nx >= 2 && ny >= 2) {
        for (uint64_t x = 0; x + 1 < nx; x++) {
            for (uint64_t y = 0; y + 1 < ny; y++) {
                uint64_t c = IDX(s, x, y);
                double p_i_j = s->p[c], p_i_p1_j = s->p[c + ny], p_i_j_p1 = s->p[c + 1];
                s->u[c] = s->f[c] - (delt / delx) * (p_i_p1_j - p_i_j);
                s->v[c] = s->g[c] - (delt / dely) * (p_i_j_p1 - p_i_j);
            }
        }
    }
    for (uint64_t k = 0; k < s->n_restore; k++) {
        const so_restore *r = &s->restore[k];
        if (r->has_u) s->u[r->idx] = r->u;
        if (r->has_v) s->v[r->idx] = r->v;
    }
    so_calculate_speed_range(s);
}

/* extension A9: NaSt2D COMP_delt; maxima over Fluid cells */
static void adapt_delt(so_sim *s) {
    double umax = 0.0, vmax = 0.0;
    for (uint64_t i = 0; i < s->n; i++) {
        if (s->kind[i] == SO_KIND_FLUID) {
            umax = fmax(umax, fabs(s->u[i]));
            vmax = fmax(vmax, fabs(s->v[i]));
        }
    }
    double dx = s->prm.delx, dy = s->prm.dely;
    double d = (s->prm.reynolds / 2.0) / ((1.0 / (dx * dx)) + (1.0 / (dy * dy)));
    if (umax > 0.0) d = fmin(d, dx / umax);
    if (vmax > 0.0) d = fmin(d, dy / vmax);
    s->prm.delt = s->prm.tau * d;
}

/* run_simulation_tick: src/simulation.rs:324-333 */
int so_tick(so_sim *s, uint32_t *iters, double *norm_squared) {
    if (s->prm.tau > 0.0) adapt_delt(s);
    int rc = so_set_boundary_u_and_v(s);
    if (rc) return rc;
    so_calculate_f_and_g(s);
    so_calculate_rhs(s);
    rc = so_solve_sor(s, iters, norm_squared);
    if (rc) return rc;
    so_set_u_and_v(s);
    s->time += s->prm.delt;
    s->iterations += 1;
    return SO_OK;
}

/* ------------------------------------------------------------------ */
/* construction: src/simulation.rs:71-99, src/grid/mod.rs:127-153      */
/* ------------------------------------------------------------------ */
static double *dup_or_zero(const double *src, uint64_t n) {
    double *d = (double *)calloc(n ? n : 1, sizeof(double));
    if (src) memcpy(d, src, n * sizeof(double));
    return d;
}

int so_create(const so_params *prm, const double *p, const double *u, const double *v,
              const uint8_t *kind, const double *bu, const double *bv, so_sim **out,
              uint64_t err_xy[2]) {
    *out = NULL;
    if (!prm || !kind || prm->nx == 0 || prm->ny == 0) return SO_INVALID;
    so_sim *s = (so_sim *)calloc(1, sizeof(so_sim));
    s->prm = *prm;
    s->nx = prm->nx;
    s->ny = prm->ny;
    s->n = s->nx * s->ny;
    s->p = dup_or_zero(p, s->n);
    s->u = dup_or_zero(u, s->n);
    s->v = dup_or_zero(v, s->n);
    s->bu = dup_or_zero(bu, s->n);
    s->bv = dup_or_zero(bv, s->n);
    s->f = dup_or_zero(NULL, s->n);
    s->g = dup_or_zero(NULL, s->n);
    s->rhs = dup_or_zero(NULL, s->n);
    s->kind = (uint8_t *)malloc(s->n);
    memcpy(s->kind, kind, s->n);
    s->has_initial_norm = prm->has_initial_norm;
    s->initial_norm_squared = prm->initial_norm_squared;
    s->time = prm->time;
    s->iterations = prm->iterations;
    int rc = so_rebuild_boundary_list(s, err_xy);
    if (rc) {
        so_destroy(s);
        return rc;
    }
    so_calculate_pressure_range(s);
    so_calculate_speed_range(s);
    so_calculate_f_and_g(s);
    so_calculate_rhs(s);
    get_initial_norm_squared(s);
    *out = s;
    return SO_OK;
}

void so_destroy(so_sim *s) {
    if (!s) return;
    free(s->p); free(s->u); free(s->v); free(s->f); free(s->g); free(s->rhs);
    free(s->bu); free(s->bv); free(s->kind); free(s->blist); free(s->restore);
    free(s);
}

double *so_p(so_sim *s) { return s->p; }
double *so_u(so_sim *s) { return s->u; }
double *so_v(so_sim *s) { return s->v; }
double *so_f(so_sim *s) { return s->f; }
double *so_g(so_sim *s) { return s->g; }
double *so_rhs(so_sim *s) { return s->rhs; }
uint8_t *so_kind(so_sim *s) { return s->kind; }
double *so_bu(so_sim *s) { return s->bu; }
double *so_bv(so_sim *s) { return s->bv; }

void so_get_state(const so_sim *s, so_state *st) {
    st->time = s->time;
    st->delt = s->prm.delt;
    st->iterations = s->iterations;
    st->has_initial_norm = s->has_initial_norm;
    st->initial_norm_squared = s->initial_norm_squared;
    st->pressure_range[0] = s->pressure_range[0];
    st->pressure_range[1] = s->pressure_range[1];
    st->speed_range[0] = s->speed_range[0];
    st->speed_range[1] = s->speed_range[1];
    st->fluid_cells = s->fluid_cells;
    st->n_boundary = s->n_boundary;
}

void so_set_params(so_sim *s, const so_params *prm) {
    uint64_t nx = s->prm.nx, ny = s->prm.ny;
    s->prm = *prm;
    s->prm.nx = nx;
    s->prm.ny = ny;
}

/* sim.initial_norm_squared = None (the field is pub, src/simulation.rs:62): the next
 * solve_sor latches the norm after its first sweep (src/simulation.rs:229-237, 276) */
void so_clear_initial_norm(so_sim *s) {
    s->has_initial_norm = 0;
    s->initial_norm_squared = 0.0;
}

uint64_t so_boundary_list(const so_sim *s, uint64_t *idx, uint8_t *edge, uint64_t cap) {
    uint64_t n = s->n_boundary < cap ? s->n_boundary : cap;
    for (uint64_t k = 0; k < n; k++) {
        if (idx) idx[k] = s->blist[k].idx;
        if (edge) edge[k] = s->blist[k].edge;
    }
    return s->n_boundary;
}

/* ------------------------------------------------------------------ */
/* presets: src/grid/presets.rs                                        */
/* ------------------------------------------------------------------ */
void so_preset_empty(uint64_t nx, uint64_t ny, uint8_t *kind, double *bu, double *bv) {
    memset(kind, SO_KIND_FLUID, nx * ny);
    if (bu) memset(bu, 0, nx * ny * sizeof(double));
    if (bv) memset(bv, 0, nx * ny * sizeof(double));
}

/* src/grid/presets.rs:19-40 */
void so_preset_simple_inflow(uint64_t nx, uint64_t ny, uint8_t *kind, double *bu, double *bv) {
    so_preset_empty(nx, ny, kind, bu, bv);
    for (uint64_t x = 0; x < nx; x++) {
        kind[x * ny + 0] = SO_KIND_NOSLIP;
        kind[x * ny + (ny - 1)] = SO_KIND_NOSLIP;
    }
    for (uint64_t y = 1; y + 1 < ny; y++) {
        kind[0 * ny + y] = SO_KIND_INFLOW;
        if (bu) bu[0 * ny + y] = 1.0;
        if (bv) bv[0 * ny + y] = 0.0;
        kind[(nx - 1) * ny + y] = SO_KIND_OUTFLOW;
    }
}

/* src/grid/presets.rs:42-62; `radius as usize` truncates, saturating ops */
void so_draw_circle(uint64_t nx, uint64_t ny, uint8_t *kind, uint64_t cx, uint64_t cy,
                    double radius) {
    uint64_t r = (uint64_t)radius;
    uint64_t x_lo = cx >= r ? cx - r : 0, x_hi = cx + r;
    uint64_t y_lo = cy >= r ? cy - r : 0, y_hi = cy + r;
    for (uint64_t xi = x_lo; xi < x_hi; xi++) {
        if (xi >= nx) continue;
        int64_t x_dist = (int64_t)xi - (int64_t)cx;
        for (uint64_t yi = y_lo; yi < y_hi; yi++) {
            if (yi >= ny) continue;
            int64_t y_dist = (int64_t)yi - (int64_t)cy;
            double distance = sqrt((double)(x_dist * x_dist + y_dist * y_dist));
            if (distance < radius) kind[xi * ny + yi] = SO_KIND_NOSLIP;
        }
    }
}

/* src/grid/presets.rs:64-87 */
void so_preset_obstacle(uint64_t nx, uint64_t ny, uint8_t *kind, double *bu, double *bv) {
    so_preset_simple_inflow(nx, ny, kind, bu, bv);
    so_draw_circle(nx, ny, kind, 20, ny / 2, 5.0);
}
