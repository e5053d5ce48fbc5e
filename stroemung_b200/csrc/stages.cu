// stages.cu -- every per-tick stage except the SOR sweeps, strict IEEE f64.
//
//   K1 velocity BC      SimulationGrid::set_boundary_u_and_v   src/grid/mod.rs:414-651
//   K2 F, G             Simulation::calculate_f_and_g          src/simulation.rs:122-202
//      RHS              Simulation::calculate_rhs              src/simulation.rs:204-214
//   K3 pressure BC      copy_pressure_to_boundaries            src/grid/mod.rs:343-412
//   K5 residual norm    calculate_norm_squared / residual      src/simulation.rs:216-227
//   K6 velocity update  Simulation::set_u_and_v                src/simulation.rs:287-322
//   K7 ranges           calculate_{pressure,speed}_range       src/grid/mod.rs:237-268
// (paths relative to /root/reference).  Compiled with -fmad=false so that no a*b+c
// is contracted; see cellops.cuh for the operator restatements.
#include "cellops.cuh"
#include "sor_rb.cuh"
#include "slab_dev.cuh"

#include <float.h>

namespace sb {

namespace {

constexpr int TPB = 256;

// ---- block reductions ---------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
    for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
    for (int o = 16; o; o >>= 1) v = fmin(v, __shfl_down_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
    for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
    return v;
}

// ---- K1: velocity boundary conditions ------------------------------------------------
// The reference pass is sequential and in place.  Restated as a gather:
// every read sees the pre-pass value, except v[west_neighbor] (edges W, NW), which the
// boundary cell (bx-1, by+1) -- earlier in x-major order -- may already have overwritten
// through its own north neighbour: NoSlip -> 0.0, Inflow/MovingWall -> its v, else unchanged.
// Phase 1 (this kernel) gathers into list scratch, phase 2 scatters; two launches make the
// "pre-pass" reads race-free.
__global__ void velocity_bc_gather(Geom g, const double *__restrict__ u,
                                   const double *__restrict__ v,
                                   const uint8_t *__restrict__ cflag, BList bl, int64_t row0,
                                   int64_t row1) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= bl.n) return;
    int64_t b = bl.lin[k];
    int64_t lx = b / g.pitch;
    int kind = bl.ke[k] & 7, edge = bl.ke[k] >> 3;
    double ub = u[b], vb = v[b];
    double nu = ub, nv = vb;         // new u[b], v[b]
    double wu = 0.0, wv = 0.0;       // values for u[w], v[n] (when written)
    double ru = ub, rv = vb;         // what set_u_and_v's restore leaves in u[b], v[b]
    if (edge != SB_EDGE_NONE && lx >= row0 && lx < row1) {
        int64_t n = b - 1, so = b + 1, ea = b + g.pitch, w = b - g.pitch;
        bool has_n = edge == SB_EDGE_N || edge == SB_EDGE_NE || edge == SB_EDGE_NW;
        bool has_e = edge == SB_EDGE_NE || edge == SB_EDGE_E || edge == SB_EDGE_SE;
        bool has_s = edge == SB_EDGE_SE || edge == SB_EDGE_S || edge == SB_EDGE_SW;
        bool has_w = edge == SB_EDGE_SW || edge == SB_EDGE_W || edge == SB_EDGE_NW;
        double u_n = has_n ? u[n] : 0.0, v_n = has_n ? v[n] : 0.0;
        double u_e = has_e ? u[ea] : 0.0, v_e = has_e ? v[ea] : 0.0;
        double u_s = has_s ? u[so] : 0.0, v_s = has_s ? v[so] : 0.0;
        double u_w = has_w ? u[w] : 0.0, v_w = has_w ? v[w] : 0.0;
        if (has_w && (edge == SB_EDGE_W || edge == SB_EDGE_NW)) {
            // the one read-after-write of the sequential pass: cell (bx-1, by+1)
            int64_t by = b - lx * g.pitch;
            uint8_t fl = by + 1 < g.NY ? cflag[w + 1] : (uint8_t)0;
            int k2 = cf_kind(fl);
            if ((fl & CF_VALID) && k2 != SB_KIND_FLUID && k2 != SB_KIND_OUTFLOW) {
                // its edge contains North (w is fluid), so it writes v[w] = its boundary_v
                if (k2 == SB_KIND_NOSLIP) v_w = 0.0;
                else {
                    // Inflow / MovingWall: look its velocity up in the list
                    int64_t key = w + 1;
                    uint64_t lo = 0, hi = bl.n;
                    while (lo < hi) {
                        uint64_t mid = (lo + hi) >> 1;
                        if (bl.lin[mid] < key) lo = mid + 1;
                        else hi = mid;
                    }
                    v_w = bl.bv[lo];
                }
            }
        }
        if (kind == SB_KIND_OUTFLOW) {
            // src/grid/mod.rs:489-536
            switch (edge) {
            case SB_EDGE_N: nu = u_n; nv = v_n; break;
            case SB_EDGE_NE: nu = u_n; nv = v_e; break;
            case SB_EDGE_E: nu = u_e; nv = v_e; break;
            case SB_EDGE_SE: nu = u_e; nv = v_s; break;
            case SB_EDGE_S: nu = u_s; nv = v_s; break;
            case SB_EDGE_SW: nu = u_w; nv = v_s; break;
            case SB_EDGE_W: nu = u_w; nv = v_w; break;
            case SB_EDGE_NW: nu = u_n; nv = v_w; break;
            }
            // second restore record holds the (unchanged) neighbour values (:602-648)
            ru = has_w ? u_w : nu;
            rv = has_n ? v_n : nv;
            // mark "no neighbour writes" with NaN-free sentinel: handled in scatter by kind
        } else {
            double bu = 0.0, bv = 0.0;
            if (kind != SB_KIND_NOSLIP) { bu = bl.bu[k]; bv = bl.bv[k]; }
            if (kind == SB_KIND_MOVING_WALL) {
                // extension: tangential ghost value reflects about the wall velocity
                switch (edge) {
                case SB_EDGE_N: nu = (2.0 * bu) - u_n; break;
                case SB_EDGE_NE: nu = bu; nv = (2.0 * bv) - v_e; break;
                case SB_EDGE_E: nu = bu; nv = (2.0 * bv) - v_e; break;
                case SB_EDGE_SE: nu = bu; nv = bv; break;
                case SB_EDGE_S: nu = (2.0 * bu) - u_s; nv = bv; break;
                case SB_EDGE_SW: nu = (2.0 * bu) - u_s; nv = bv; break;
                case SB_EDGE_W: nv = (2.0 * bv) - v_w; break;
                case SB_EDGE_NW: nu = (2.0 * bu) - u_n; nv = (2.0 * bv) - v_w; break;
                }
            } else {
                // NoSlip (:437-488) and Inflow (:537-586) share one table
                switch (edge) {
                case SB_EDGE_N: nu = -u_n; break;
                case SB_EDGE_NE: nu = bu; nv = -v_e; break;
                case SB_EDGE_E: nu = bu; nv = -v_e; break;
                case SB_EDGE_SE: nu = bu; nv = bv; break;
                case SB_EDGE_S: nu = -u_s; nv = bv; break;
                case SB_EDGE_SW: nu = -u_s; nv = bv; break;
                case SB_EDGE_W: nv = -v_w; break;
                case SB_EDGE_NW: nu = -u_n; nv = -v_w; break;
                }
            }
            wu = bu;  // u[west_neighbor] = boundary_u  (edges SW, W, NW)
            wv = bv;  // v[north_neighbor] = boundary_v (edges N, NE, NW)
            ru = has_w ? bu : nu;
            rv = has_n ? bv : nv;
        }
    }
    bl.nu[k] = nu; bl.nv[k] = nv; bl.wu[k] = wu; bl.wv[k] = wv;
    bl.ru[k] = ru; bl.rv[k] = rv;
}

__global__ void velocity_bc_scatter(Geom g, double *__restrict__ u, double *__restrict__ v,
                                    BList bl, int64_t row0, int64_t row1) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= bl.n) return;
    int64_t b = bl.lin[k];
    int64_t lx = b / g.pitch;
    int kind = bl.ke[k] & 7, edge = bl.ke[k] >> 3;
    if (edge == SB_EDGE_NONE || lx < row0 || lx >= row1) return;
    u[b] = bl.nu[k];
    v[b] = bl.nv[k];
    if (kind != SB_KIND_OUTFLOW) {
        if (edge == SB_EDGE_SW || edge == SB_EDGE_W || edge == SB_EDGE_NW) u[b - g.pitch] = bl.wu[k];
        if (edge == SB_EDGE_N || edge == SB_EDGE_NE || edge == SB_EDGE_NW) v[b - 1] = bl.wv[k];
    }
}

// ---- K2: F, G and RHS in one pass -------------------------------------------------------------
// F, G: all interior cells get the stencil value (fluid or not); afterwards the reference
// overwrites f[b]=u[b], g[b]=v[b] on every boundary cell and g[n]=v[n] / f[w]=u[w] on the
// fluid cell north / west of one (src/simulation.rs:167-201).  Per cell that is:
//   non-fluid c                       -> f = u[c], g = v[c]
//   fluid c with (x+1, y) non-fluid   -> f = u[c]   (c is that cell's west neighbour)
//   fluid c with (x, y+1) non-fluid   -> g = v[c]   (c is that cell's north neighbour)
// Ring cells are outside the stencil loop: they only receive the overwrites (a fluid ring cell
// keeps whatever f, g it had).
// RHS (src/simulation.rs:204-214), all cells with x >= 1 and y >= 1, fluid or not:
//   rhs = (((f[x][y] - f[x-1][y]) / dx) + ((g[x][y] - g[x][y-1]) / dy)) / dt
//
// One warp owns 31 output columns (lane 0 is a helper that only produces G of the column to
// the left, which lane 1 needs for its rhs) and marches FGR_ROWS rows along x with a 3-row
// register window of u and v: 6 loads per cell instead of 18, F of the previous row stays in
// a register for the rhs, G of the north neighbour comes by one shuffle.  The row in front
// of the strip is evaluated for F only.  All divisions are by run constants (DivC).
constexpr int FGR_ROWS = 32;
constexpr int FGR_WARPS = 8;
constexpr int FGR_COLS = 31 * FGR_WARPS;  // output columns per block

struct FgrConsts {
    // divisors and their reciprocals: dx*dx, dy*dy, 4dx, 4dy, Re, dx, dy, dt
    double d[8], r[8];
    int fast;  // every divisor qualifies for the correction-step path (make_divc)
    double delt, gamma;
};

template <class D>
struct FgrDiv {
    FgDiv<D> k;
    D dx, dy, dt;
};

// F, G (final values incl. the boundary overwrites) of one fluid interior cell and, from
// them, its rhs; `D` decides how the 13 divisions are carried out
template <class D>
__device__ __forceinline__ void fgr_cell(const Stencil9 &su, const Stencil9 &sv,
                                         const FgrDiv<D> &dv, double delt, double gamma,
                                         bool need_f, bool need_g, double &fv, double &gv) {
    if (need_f) fv = calculate_f(su, sv, dv.k, delt, gamma);
    if (need_g) gv = calculate_g(su, sv, dv.k, delt, gamma);
}

// Three CTAs per SM: 80 registers and 36 bytes of spills -- the row load_row() prefetches is
// spilled right after the load, and the STL waiting for that load shows up with 23 % of the
// kernel's stall samples (profiles/r1_fg_rhs_8192_stalls.txt).  Two CTAs per SM (121 registers,
// no spills, -DFGR_MINB=2) were measured and are not faster: non-SOR part of a tick 2.42 vs 2.32 ms.
#ifndef FGR_MINB
#define FGR_MINB 3
#endif
__global__ void __launch_bounds__(32 * FGR_WARPS, FGR_MINB)
fg_rhs_kernel(Geom g, const double *__restrict__ u, const double *__restrict__ v,
              const uint8_t *__restrict__ cflag, double *__restrict__ f, double *__restrict__ gq,
              double *__restrict__ rhs, int64_t fg_row0, int64_t row1, int64_t rhs_row0,
              FgrConsts kc, int write_fg, int write_rhs) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t y = (int64_t)blockIdx.y * FGR_COLS - 1 + (int64_t)warp * 31 + lane;
    const int64_t xs = fg_row0 + (int64_t)blockIdx.x * FGR_ROWS;  // first output row
    const int64_t xe = min(xs + (int64_t)FGR_ROWS, row1);
    if (y - lane >= g.NY) return;  // warp-uniform: nothing of this warp is inside the grid
    const bool y_ok = y >= 0 && y < g.NY;
    const bool out_lane = lane >= 1 && y_ok;
    // clamped column indices: loads stay inside the row, clamped values are never used by a
    // cell that takes the stencil (those have all eight neighbours)
    const int64_t yc = min(max(y, (int64_t)0), g.NY - 1);
    const int64_t yn = max(yc - 1, (int64_t)0), ys = min(yc + 1, g.NY - 1);
    const bool has_s = y + 1 < g.NY;
    const int64_t last = g.nxl - 1;
    auto rowp = [&](int64_t lx) { return min(max(lx, (int64_t)0), last) * g.pitch; };

    unsigned bad = 0;
    FgrDiv<DivF> df;
    {
        DivF *slots[8] = {&df.k.dx2, &df.k.dy2, &df.k.four_dx, &df.k.four_dy, &df.k.re,
                          &df.dx,    &df.dy,    &df.dt};
#pragma unroll
        for (int i = 0; i < 8; i++) {
            slots[i]->d = kc.d[i];
            slots[i]->r = kc.r[i];
            slots[i]->bad = &bad;
        }
    }
    FgrDiv<DivT> dt_;
    dt_.k.dx2.d = kc.d[0]; dt_.k.dy2.d = kc.d[1]; dt_.k.four_dx.d = kc.d[2];
    dt_.k.four_dy.d = kc.d[3]; dt_.k.re.d = kc.d[4];
    dt_.dx.d = kc.d[5]; dt_.dy.d = kc.d[6]; dt_.dt.d = kc.d[7];
    const bool fast = kc.fast != 0;

    // the pre-row (xs - 1) is needed when this strip produces rhs for row xs
    const bool pre = write_rhs && xs >= rhs_row0 && xs - 1 >= 0;
    int64_t x = pre ? xs - 1 : xs;
    // window rows: 0 = x-1, 1 = x, 2 = x+1, 3 = x+2 (in flight);  columns n, c, s
    double un[4], uc[4], us[4], vn[4], vc[4], vs[4];
    uint8_t fl_c[4], fl_s[4];  // flags of (row, y) and (row, y+1); index 0 unused
    auto load_row = [&](int k, int64_t lx) {
        const int64_t r = rowp(lx);
        un[k] = u[r + yn]; uc[k] = u[r + yc]; us[k] = u[r + ys];
        vn[k] = v[r + yn]; vc[k] = v[r + yc]; vs[k] = v[r + ys];
        const bool in = lx >= 0 && lx < g.nxl;
        fl_c[k] = in ? cflag[r + yc] : (uint8_t)0;
        fl_s[k] = (in && has_s) ? cflag[r + ys] : (uint8_t)0;
    };
    load_row(0, x - 1);
    load_row(1, x);
    load_row(2, x + 1);
    double f_west = 0.0;  // final f of (x-1, y)
    for (; x < xe; x++) {
        load_row(3, x + 2);  // consumed in the next iteration: a full row of compute hides it

        const int64_t gx = g.gx0 + x;
        const int64_t c = x * g.pitch + yc;
        const uint8_t fl = fl_c[1];
        const bool valid = y_ok && (fl & CF_VALID);
        const bool interior = gx >= 1 && gx <= g.NX - 2 && y >= 1 && y <= g.NY - 2;
        const bool want_rhs = write_rhs && x >= xs && x >= rhs_row0 && gx >= 1 && y >= 1;
        double fv = 0.0, gv = 0.0;
        bool st_f = false, st_g = false, need_f = false, need_g = false;
        if (valid) {
            if (!cf_is_fluid(fl)) {
                fv = uc[1];
                gv = vc[1];
                st_f = st_g = true;
            } else {
                const bool east_solid = gx + 1 < g.NX && x + 1 < g.nxl && !cf_is_fluid(fl_c[2]);
                const bool south_solid = has_s && !cf_is_fluid(fl_s[1]);
                if (east_solid) { fv = uc[1]; st_f = true; }
                else if (interior) { need_f = st_f = true; }
                else fv = f[c];   // fluid ring cell: untouched by the reference
                if (south_solid) { gv = vc[1]; st_g = true; }
                else if (interior) { need_g = st_g = true; }
                else gv = gq[c];
            }
        }
        Stencil9 su, sv;
        su.nw = un[0]; su.w = uc[0]; su.sw = us[0];
        su.n = un[1];  su.c = uc[1]; su.s = us[1];
        su.ne = un[2]; su.e = uc[2]; su.se = us[2];
        sv.nw = vn[0]; sv.w = vc[0]; sv.sw = vs[0];
        sv.n = vn[1];  sv.c = vc[1]; sv.s = vs[1];
        sv.ne = vn[2]; sv.e = vc[2]; sv.se = vs[2];
        double rv = 0.0;
        bad = fast ? 0u : 1u;
        if (fast) {
            fgr_cell(su, sv, df, kc.delt, kc.gamma, need_f, need_g, fv, gv);
            const double g_north = __shfl_up_sync(0xffffffffu, gv, 1);
            rv = df.dt(df.dx(fv - f_west) + df.dy(gv - g_north));
        }
        // rare: an operand outside the safe exponent window somewhere in this warp's row --
        // every lane redoes its cell with plain IEEE divisions (the shuffle needs all lanes)
        if (__any_sync(0xffffffffu, bad != 0u)) {
            fgr_cell(su, sv, dt_, kc.delt, kc.gamma, need_f, need_g, fv, gv);
            const double g_north = __shfl_up_sync(0xffffffffu, gv, 1);
            rv = dt_.dt(dt_.dx(fv - f_west) + dt_.dy(gv - g_north));
        }
        if (x >= xs && out_lane && valid) {  // not the pre-row
            if (write_fg && st_f) f[c] = fv;
            if (write_fg && st_g) gq[c] = gv;
            if (want_rhs) rhs[c] = rv;
        }
        f_west = fv;
        // slide the window
#pragma unroll
        for (int k = 0; k < 3; k++) {
            un[k] = un[k + 1]; uc[k] = uc[k + 1]; us[k] = us[k + 1];
            vn[k] = vn[k + 1]; vc[k] = vc[k + 1]; vs[k] = vs[k + 1];
            fl_c[k] = fl_c[k + 1]; fl_s[k] = fl_s[k + 1];
        }
    }
}

// ---- K2, performance mode: the same stage with the FgFast arithmetic (cellops.cuh) -----------
// Same decomposition as fg_rhs_kernel (one warp = 31 output columns + a helper lane, marching
// along x), rebuilt around the ~70 FP64 instructions per cell that are left once the exact
// divisions are gone -- at that size the bookkeeping of the strict kernel (window slide,
// per-cell role tests: 310 of its 377 instructions in this mode) would dominate:
//   * the 3-row window rotates by NAME (the row loop is unrolled 4x), nothing is copied;
//   * a row whose 32 cells are all fluid cells without a non-fluid neighbour (flag bit
//     CF_NEAR clear) inside the stencil range takes a straight-line path: no role tests, no
//     flag of any neighbour is read; everything else (walls, obstacles, the ring, slab
//     edges) takes the general path with the reference's overwrite rules
//     (src/simulation.rs:167-201);
//   * the row two ahead is requested before the current one is computed.
constexpr int FGF_ROWS = 64;
constexpr int FGF_WARPS = 8;
constexpr int FGF_COLS = 31 * FGF_WARPS;
#ifndef FGF_MINB
#define FGF_MINB 2
#endif
#ifndef FGF_PF
#define FGF_PF 1   // rows requested ahead of the stencil window
#endif

struct FgRow {
    double un, uc, us, vn, vc, vs;
    unsigned fl;   // flag byte of (row, y); 0 outside the local rows
};

__global__ void __launch_bounds__(32 * FGF_WARPS, FGF_MINB)
fg_rhs_fast_kernel(Geom g, const double *__restrict__ u, const double *__restrict__ v,
                   const uint8_t *__restrict__ cflag, double *__restrict__ f,
                   double *__restrict__ gq, double *__restrict__ rhs, int fg_row0, int row1,
                   FgFast kf, int write_fg, int write_rhs) {
    // 32-bit row / column arithmetic (rows and columns are < 2^31; the launcher checks);
    // 64-bit only in the row base pointers, which advance by one pitch per step
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int NY = (int)g.NY, NX = (int)g.NX, nxl = (int)g.nxl, gx0 = (int)g.gx0;
    const int y = (int)blockIdx.y * FGF_COLS - 1 + warp * 31 + lane;
    const int xs = fg_row0 + (int)blockIdx.x * FGF_ROWS;  // first output row
    const int xe = min(xs + FGF_ROWS, row1);
    if (y - lane >= NY) return;  // warp-uniform: nothing of this warp is inside the grid
    const bool y_ok = y >= 0 && y < NY;
    const bool out_lane = lane >= 1 && y_ok;
    const bool y_int = y >= 1 && y <= NY - 2;
    // clamped column indices: loads stay inside the row; a clamped value is never used by a
    // cell that takes the stencil (those have all eight neighbours)
    const int yc = min(max(y, 0), NY - 1);
    const int yn = max(yc - 1, 0), ys = min(yc + 1, NY - 1);
    const bool has_s = y + 1 < NY;
    const int last = nxl - 1;
    const int64_t pitch = g.pitch;

    // the pre-row (xs - 1) is needed when this strip produces rhs for row xs
    const bool pre = write_rhs && xs - 1 >= 0;
    const int x_first = pre ? xs - 1 : xs;
    // row `lr` = the next row to load, clamped to the local rows; its base offset
    int lr = x_first - 1;
    int64_t lro = (int64_t)min(max(lr, 0), last) * pitch;
    auto load_row = [&](FgRow &r) {
        const double *ur = u + lro, *vr = v + lro;
        r.un = ur[yn]; r.uc = ur[yc]; r.us = ur[ys];
        r.vn = vr[yn]; r.vc = vr[yc]; r.vs = vr[ys];
        r.fl = (unsigned)lr <= (unsigned)last ? (unsigned)cflag[lro + yc] : 0u;
        lro += (lr >= 0 && lr < last) ? pitch : 0;   // the clamp, kept incrementally
        lr++;
    };
    // FGF_PF rows are in flight ahead of the three the stencil uses; the window rotates by
    // NAME (the row loop is unrolled over its NS slots), nothing is copied
    constexpr int NS = 3 + FGF_PF;
    FgRow W[NS];   // slot k of step i holds row x - 1 + ((k - i) mod NS)
#pragma unroll
    for (int k = 0; k < NS - 1; k++) load_row(W[k]);
    double f_west = 0.0;  // final f of (x-1, y)
    int64_t c = (int64_t)x_first * pitch + yc;   // (x, y) of the row being computed
    for (int x0 = x_first; x0 < xe; x0 += NS) {
#pragma unroll
        for (int i = 0; i < NS; i++) {
            // rows past xe (the tail of the last group) are computed from clamped loads and
            // never stored
            const int x = x0 + i;
            FgRow &A = W[i % NS], &B = W[(i + 1) % NS], &C = W[(i + 2) % NS];
            load_row(W[(i + NS - 1) % NS]);   // row x + 1 + FGF_PF
            const int gx = gx0 + x;
            const bool row_int = gx >= 1 && gx <= NX - 2;
            const unsigned fl = B.fl;
            Stencil9 su, sv;
            su.nw = A.un; su.w = A.uc; su.sw = A.us;
            su.n = B.un;  su.c = B.uc; su.s = B.us;
            su.ne = C.un; su.e = C.uc; su.se = C.us;
            sv.nw = A.vn; sv.w = A.vc; sv.sw = A.vs;
            sv.n = B.vn;  sv.c = B.vc; sv.s = B.vs;
            sv.ne = C.vn; sv.e = C.vc; sv.se = C.vs;
            // the stencil values, needed or not (one copy of the arithmetic)
            double fv = calculate_f_fast(su, sv, kf);
            double gv = calculate_g_fast(su, sv, kf);
            bool valid = true, st_f = true, st_g = true;
            // a fluid cell with four fluid neighbours: valid, kind 0, CF_NEAR clear
            const bool plain = row_int && y_int && (fl & 0x8fu) == (unsigned)CF_FLUID;
            if (!__all_sync(0xffffffffu, plain)) {   // rare: walls, obstacles, ring, slab edges
                valid = y_ok && (fl & CF_VALID);
                if (!valid) {
                    fv = gv = 0.0;
                    st_f = st_g = false;
                } else if (!cf_is_fluid((uint8_t)fl)) {
                    fv = B.uc;
                    gv = B.vc;
                } else {
                    const bool interior = row_int && y_int;
                    const bool east_solid = gx + 1 < NX && x + 1 < nxl &&
                                            !cf_is_fluid((uint8_t)C.fl);
                    const bool in_rows = (unsigned)x <= (unsigned)last;
                    const uint8_t fl_s = (has_s && in_rows) ? cflag[c + 1] : (uint8_t)0;
                    const bool south_solid = has_s && !cf_is_fluid(fl_s);
                    if (east_solid) fv = B.uc;
                    else if (!interior) { fv = in_rows ? f[c] : 0.0; st_f = false; }  // fluid ring cell
                    if (south_solid) gv = B.vc;
                    else if (!interior) { gv = in_rows ? gq[c] : 0.0; st_g = false; }
                }
            }
            const double g_north = __shfl_up_sync(0xffffffffu, gv, 1);
            if (x >= xs && x < xe && out_lane && valid) {  // not the pre-row, not the tail
                if (write_fg && st_f) f[c] = fv;
                if (write_fg && st_g) gq[c] = gv;
                if (write_rhs && gx >= 1 && y >= 1)
                    rhs[c] = rhs_fast(kf, fv, f_west, gv, g_north);
            }
            f_west = fv;
            c += pitch;
        }
    }
}

// RHS on its own (sb_calculate_rhs): reads f, g from memory
__global__ void rhs_kernel(Geom g, const double *__restrict__ f, const double *__restrict__ gq,
                           double *__restrict__ rhs, int64_t row0, int64_t row1, DivC dx,
                           DivC dy, DivC dt, int fast, FgFast kf) {
    int64_t y = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
    int64_t lx = row0 + blockIdx.x;
    if (y < 1 || y >= g.NY || lx >= row1) return;
    int64_t gx = g.gx0 + lx;
    if (gx < 1 || gx >= g.NX) return;
    int64_t c = lx * g.pitch + y;
    if (fast) rhs[c] = rhs_fast(kf, f[c], f[c - g.pitch], gq[c], gq[c - 1]);
    else rhs[c] = dt(dx(f[c] - f[c - g.pitch]) + dy(gq[c] - gq[c - 1]));
}

// ---- K3: pressure BC over the boundary list (reads fluid cells, writes boundary cells) ---
__global__ void pressure_bc_kernel(Geom g, double *const *__restrict__ pbuf,
                                   const SorCtl *__restrict__ ctl, int guarded, BList bl) {
    if (guarded && ctl->active_T == 0) return;
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= bl.n) return;
    int edge = bl.ke[k] >> 3;
    if (edge == SB_EDGE_NONE) return;
    double *p = pbuf[ctl->src];
    int64_t b = bl.lin[k];
    int64_t n = b - 1, so = b + 1, ea = b + g.pitch, w = b - g.pitch;
    switch (edge) {
    case SB_EDGE_N: p[b] = p[n]; break;
    case SB_EDGE_NE: p[b] = (p[n] + p[ea]) / 2.0; break;
    case SB_EDGE_E: p[b] = p[ea]; break;
    case SB_EDGE_SE: p[b] = (p[so] + p[ea]) / 2.0; break;
    case SB_EDGE_S: p[b] = p[so]; break;
    case SB_EDGE_SW: p[b] = (p[so] + p[w]) / 2.0; break;
    case SB_EDGE_W: p[b] = p[w]; break;
    case SB_EDGE_NW: p[b] = (p[n] + p[w]) / 2.0; break;
    }
}

// ---- K5: residual norm, partial sums over ALL interior cells (fluid or not) ------------
// One block per (row, 1024-column segment)-ish chunk; fixed tree inside the block, fixed
// order in the finalize kernel => deterministic.  The reference folds sequentially in
// row-major order; the different association is covered by the 1e-12 relative allowance.
constexpr int NORM_ROWS_PER_BLOCK = 4;
__global__ void norm_partial_kernel(Geom g, double *const *__restrict__ pbuf,
                                    const SorCtl *__restrict__ ctl, int guarded,
                                    const double *__restrict__ rhs, double *__restrict__ partial,
                                    DivC dx2, DivC dy2) {
    if (guarded && ctl->active_T == 0) return;
    const double *p = pbuf[ctl->src];
    double acc = 0.0;
    int64_t lx_base = g.own0 + (int64_t)blockIdx.x * NORM_ROWS_PER_BLOCK;
    for (int r = 0; r < NORM_ROWS_PER_BLOCK; r++) {
        int64_t lx = lx_base + r;
        int64_t gx = g.gx0 + lx;
        if (lx >= g.own1 || gx < 1 || gx > g.NX - 2) continue;
        for (int64_t y = 1 + (int64_t)blockIdx.y * blockDim.x + threadIdx.x; y <= g.NY - 2;
             y += (int64_t)gridDim.y * blockDim.x) {
            int64_t c = lx * g.pitch + y;
            double rr = residual(p[c], p[c - 1], p[c + 1], p[c - g.pitch], p[c + g.pitch], dx2,
                                 dy2, rhs[c]);
            acc = acc + (rr * rr);
        }
    }
    __shared__ double sh[TPB / 32];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < TPB / 32; k++) t += sh[k];
        partial[(int64_t)blockIdx.x * gridDim.y + blockIdx.y] = t;
    }
}

// ---- finalize: total the partials of each level, apply the exit test, advance ctl -------
// partial layout: [level][part].  Exit rule of src/simulation.rs:279:
//     norm < initial_norm || norm < eps^2   -> stop after this sweep.
// A red-black pass runs T sweeps speculatively; if the rule fires at level k < T the pass
// is repeated from the same source buffer with T = k (deterministic, so it then ends at k).
// SB_FIN_TRACE: where a slab-mode finalize spends its time (globaltimer ns, this device):
// [0] launches, [1] inside the kernel, [2] of that inside the all-gather, [3] end of the last
// one, [4] from the end of one to the start of the next (= the pass in between, launch gaps
// included)
__device__ unsigned long long g_fin_trace[8];
__device__ __forceinline__ unsigned long long fin_gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__global__ void sor_finalize_kernel(SorCtl *__restrict__ ctl, const double *__restrict__ partial,
                                    int nparts, double fluid_cells, double initial_norm,
                                    double eps2, int test_exit, double *__restrict__ norm_hist,
                                    SlabLink lk) {
    __shared__ double sh[1024 / 32];
    __shared__ double level_sum[8];
    __shared__ double level_norm[8];
    __shared__ double gathered[SB_MAX_WORLD * 8];
    int T = ctl->active_T;
    if (T == 0) return;
    const unsigned long long tr0 = fin_gtime();
    unsigned long long tr1 = tr0, tr2 = tr0;
    for (int lvl = 0; lvl < T; lvl++) {
        double acc = 0.0;
        for (int i = threadIdx.x; i < nparts; i += blockDim.x)
            acc += partial[(int64_t)lvl * nparts + i];
        acc = warp_sum(acc);
        if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int k = 0; k < (int)(blockDim.x >> 5); k++) t += sh[k];
            level_sum[lvl] = t;
        }
        __syncthreads();
    }
    if (lk.world > 1) {
        // row slabs: every rank sums the per-slab totals in rank order -> identical norms and
        // identical exit decisions on all GPUs, no host in the loop (slab_dev.cuh)
        tr1 = fin_gtime();
        const bool gathered_ok = slab_allgather(lk, level_sum, T, gathered);
        tr2 = fin_gtime();
        if (!gathered_ok) {
            if (threadIdx.x == 0) {  // a peer went missing: end the solve, the host reports it
                ctl->active_T = 0;
                ctl->finished = 1;
            }
            return;
        }
        if ((int)threadIdx.x < T) {
            double t = gathered[threadIdx.x];
            for (int r = 1; r < lk.world; r++) t += gathered[r * 8 + threadIdx.x];
            level_sum[threadIdx.x] = t;
        }
        __syncthreads();
    }
    if ((int)threadIdx.x < T) level_norm[threadIdx.x] = level_sum[threadIdx.x] / fluid_cells;
    __syncthreads();
    if (threadIdx.x != 0) return;
    sor_advance_ctl(ctl, level_norm, T, initial_norm, eps2, test_exit, norm_hist);
    const unsigned long long tr3 = fin_gtime();
    g_fin_trace[0] += 1;
    g_fin_trace[1] += tr3 - tr0;
    g_fin_trace[2] += tr2 - tr1;
    if (g_fin_trace[3] && tr0 - g_fin_trace[3] < 100000000ull) g_fin_trace[4] += tr0 - g_fin_trace[3];
    g_fin_trace[3] = tr3;
}



// out[0] = (sum of partial[0..n)) / fluid_cells, same tree as the finalize kernel
__global__ void sum_partials_kernel(const double *__restrict__ partial, int nparts,
                                    double fluid_cells, double *__restrict__ out, int divide) {
    __shared__ double sh[1024 / 32];
    double acc = 0.0;
    for (int i = threadIdx.x; i < nparts; i += blockDim.x) acc += partial[i];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < (int)(blockDim.x >> 5); k++) t += sh[k];
        out[0] = divide ? t / fluid_cells : t;
    }
}

// ---- K6: velocity update + speed range + max|u|,|v| over fluid cells ----------------------
// u, v for 0 <= x <= NX-2, 0 <= y <= NY-2, every cell (src/simulation.rs:299-311); boundary
// cells are overwritten afterwards by the restore kernel, fluid cells are final here, so
// the Fluid-only reductions of calculate_speed_range (src/grid/mod.rs:253-268) fuse in.
// PRANGE: also the Fluid-only min / max of p (calculate_pressure_range, src/grid/mod.rs:237-251,
// which the reference runs right before this stage when SOR hit its cap, simulation.rs:283):
// p is read here anyway.  Partials: 8 doubles per block (smin, smax, umax, vmax, pmin, pmax).
constexpr int RANGE_RPB = 8;  // rows per block of the velocity-update kernel
template <bool PRANGE>
__global__ void adapt_uv_kernel(Geom g, const double *__restrict__ p,
                                const double *__restrict__ f, const double *__restrict__ gq,
                                const uint8_t *__restrict__ cflag, double *__restrict__ u,
                                double *__restrict__ v, double *__restrict__ partial,
                                double delt, double delx, double dely) {
    double smin = DBL_MAX, smax = 0.0, umax = 0.0, vmax = 0.0, pmin = DBL_MAX, pmax = 0.0;
    const double dtdx = delt / delx, dtdy = delt / dely;
    for (int r = 0; r < RANGE_RPB; r++) {  // RANGE_RPB rows per block: 8x fewer partials
        const int64_t lx = g.own0 + (int64_t)blockIdx.x * RANGE_RPB + r;
        const int64_t gx = g.gx0 + lx;
        if (!(lx < g.own1 && gx >= 0 && gx < g.NX)) continue;
        for (int64_t y = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; y < g.NY;
             y += (int64_t)gridDim.y * blockDim.x) {
            int64_t c = lx * g.pitch + y;
            double un, vn, pc = 0.0;
            const bool inner = gx <= g.NX - 2 && y <= g.NY - 2;
            const bool fluid = cf_is_fluid(cflag[c]);
            if (inner || (PRANGE && fluid)) pc = p[c];
            if (inner) {
                un = f[c] - dtdx * (p[c + g.pitch] - pc);
                vn = gq[c] - dtdy * (p[c + 1] - pc);
                u[c] = un;
                v[c] = vn;
            } else {
                un = u[c];
                vn = v[c];
            }
            if (fluid) {
                double sq = (un * un) + (vn * vn);
                smin = fmin(smin, sq);
                smax = fmax(smax, sq);
                umax = fmax(umax, fabs(un));
                vmax = fmax(vmax, fabs(vn));
                if (PRANGE) {
                    pmin = fmin(pmin, pc);
                    pmax = fmax(pmax, pc);
                }
            }
        }
    }
    constexpr int NV = PRANGE ? 6 : 4;
    __shared__ double sh[NV][TPB / 32];
    smin = warp_min(smin); smax = warp_max(smax); umax = warp_max(umax); vmax = warp_max(vmax);
    if (PRANGE) { pmin = warp_min(pmin); pmax = warp_max(pmax); }
    if ((threadIdx.x & 31) == 0) {
        int w = threadIdx.x >> 5;
        sh[0][w] = smin; sh[1][w] = smax; sh[2][w] = umax; sh[3][w] = vmax;
        if (PRANGE) { sh[4][w] = pmin; sh[5][w] = pmax; }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < TPB / 32; k++) {
            smin = fmin(smin, sh[0][k]); smax = fmax(smax, sh[1][k]);
            umax = fmax(umax, sh[2][k]); vmax = fmax(vmax, sh[3][k]);
            if (PRANGE) { pmin = fmin(pmin, sh[4][k]); pmax = fmax(pmax, sh[5][k]); }
        }
        int64_t blk = (int64_t)blockIdx.x * gridDim.y + blockIdx.y;
        constexpr int ST = PRANGE ? 8 : 4;
        partial[ST * blk + 0] = smin; partial[ST * blk + 1] = smax;
        partial[ST * blk + 2] = umax; partial[ST * blk + 3] = vmax;
        if (PRANGE) { partial[ST * blk + 4] = pmin; partial[ST * blk + 5] = pmax; }
    }
}

// replay of u_v_restore (src/simulation.rs:313-320), one entry per boundary cell
__global__ void restore_uv_kernel(Geom g, double *__restrict__ u, double *__restrict__ v,
                                  BList bl) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= bl.n) return;
    int64_t b = bl.lin[k];
    int64_t lx = b / g.pitch;
    if (lx < g.own0 || lx >= g.own1) return;
    u[b] = bl.ru[k];
    v[b] = bl.rv[k];
}

// min/max reductions over Fluid cells of the whole array; mode 0: p, mode 1: u^2+v^2, |u|, |v|
__global__ void range_kernel(Geom g, const double *__restrict__ a, const double *__restrict__ b,
                             const uint8_t *__restrict__ cflag, double *__restrict__ partial,
                             int mode) {
    int64_t lx = g.own0 + blockIdx.x;
    int64_t gx = g.gx0 + lx;
    double mn = DBL_MAX, mx = 0.0, m2 = 0.0, m3 = 0.0;
    if (lx < g.own1 && gx >= 0 && gx < g.NX) {
        for (int64_t y = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; y < g.NY;
             y += (int64_t)gridDim.y * blockDim.x) {
            int64_t c = lx * g.pitch + y;
            if (!cf_is_fluid(cflag[c])) continue;
            double val;
            if (mode == 0) val = a[c];
            else {
                val = (a[c] * a[c]) + (b[c] * b[c]);
                m2 = fmax(m2, fabs(a[c]));
                m3 = fmax(m3, fabs(b[c]));
            }
            mn = fmin(mn, val);
            mx = fmax(mx, val);
        }
    }
    __shared__ double sh[4][TPB / 32];
    mn = warp_min(mn); mx = warp_max(mx); m2 = warp_max(m2); m3 = warp_max(m3);
    if ((threadIdx.x & 31) == 0) {
        int w = threadIdx.x >> 5;
        sh[0][w] = mn; sh[1][w] = mx; sh[2][w] = m2; sh[3][w] = m3;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < TPB / 32; k++) {
            mn = fmin(mn, sh[0][k]); mx = fmax(mx, sh[1][k]);
            m2 = fmax(m2, sh[2][k]); m3 = fmax(m3, sh[3][k]);
        }
        int64_t blk = (int64_t)blockIdx.x * gridDim.y + blockIdx.y;
        partial[4 * blk + 0] = mn; partial[4 * blk + 1] = mx;
        partial[4 * blk + 2] = m2; partial[4 * blk + 3] = m3;
    }
}

// final min/max/max/max (and, with stride 8, a second min/max) over block partials -> out
// (f64::MAX, 0.0 fold seeds)
__global__ void range_final_kernel(const double *__restrict__ partial, int64_t nblk,
                                   double *__restrict__ out, int stride) {
    double mn = DBL_MAX, mx = 0.0, m2 = 0.0, m3 = 0.0, pn = DBL_MAX, px = 0.0;
    for (int64_t i = threadIdx.x; i < nblk; i += blockDim.x) {
        const double *q = partial + stride * i;
        mn = fmin(mn, q[0]); mx = fmax(mx, q[1]);
        m2 = fmax(m2, q[2]); m3 = fmax(m3, q[3]);
        if (stride == 8) { pn = fmin(pn, q[4]); px = fmax(px, q[5]); }
    }
    __shared__ double sh[6][1024 / 32];
    mn = warp_min(mn); mx = warp_max(mx); m2 = warp_max(m2); m3 = warp_max(m3);
    pn = warp_min(pn); px = warp_max(px);
    if ((threadIdx.x & 31) == 0) {
        int w = threadIdx.x >> 5;
        sh[0][w] = mn; sh[1][w] = mx; sh[2][w] = m2; sh[3][w] = m3; sh[4][w] = pn; sh[5][w] = px;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < (int)(blockDim.x >> 5); k++) {
            mn = fmin(mn, sh[0][k]); mx = fmax(mx, sh[1][k]);
            m2 = fmax(m2, sh[2][k]); m3 = fmax(m3, sh[3][k]);
            pn = fmin(pn, sh[4][k]); px = fmax(px, sh[5][k]);
        }
        out[0] = mn; out[1] = mx; out[2] = m2; out[3] = m3; out[4] = pn; out[5] = px;
    }
}

// ---- cell-level operators for the C-ABI known-answer tests ------------------------------
__global__ void cellop_kernel(int op, const double *__restrict__ in, double *__restrict__ out) {
    // in: u[9], v[9], scalars[5]
    const double *ub = in, *vb = in + 9, *sc = in + 18;
    Stencil9 su = stencil_from_block(ub), sv = stencil_from_block(vb);
    double r = 0.0;
    switch (op) {
    // plain IEEE division here: these entry points pin the operators themselves
    case 0: r = du2dx(su.w, su.c, su.e, DivT{4.0 * sc[0]}, sc[1]); break;         // delx, gamma
    case 1: r = duvdx(su.c, su.s, su.w, su.sw, sv.c, sv.e, sv.w, DivT{4.0 * sc[0]}, sc[1]); break;
    case 2: r = duvdy(su.c, su.n, su.s, sv.c, sv.n, sv.e, sv.ne, DivT{4.0 * sc[0]}, sc[1]); break;
    case 3: r = dv2dy(sv.n, sv.c, sv.s, DivT{4.0 * sc[0]}, sc[1]); break;         // dely
    case 4: r = laplacian(su.c, su.n, su.s, su.w, su.e, DivT{sc[0] * sc[0]}, DivT{sc[1] * sc[1]});
        break;                                                                    // delx, dely
    case 5: r = residual(su.c, su.n, su.s, su.w, su.e, DivT{sc[0] * sc[0]}, DivT{sc[1] * sc[1]},
                         sc[2]); break;
    case 6: r = calculate_f(su, sv, fg_div_plain<DivT>(sc[0], sc[1], sc[4]), sc[2], sc[3]); break;
    case 7: r = calculate_g(su, sv, fg_div_plain<DivT>(sc[0], sc[1], sc[4]), sc[2], sc[3]); break;
    }
    out[0] = r;
}

// rows on grid.x (2^31 limit), column blocks on grid.y
dim3 row_grid(const Geom &g, int64_t rows, int64_t cols) {
    (void)g;
    return dim3((unsigned)rows, (unsigned)((cols + TPB - 1) / TPB), 1);
}

sb_status ensure_partial(sb_sim *s, size_t need) {
    if (need <= s->partial_cap) return SB_OK;
    // stream-ordered (no device-wide synchronisation, see solve() in capi.cu)
    if (s->d_partial) SB_CUDA(cudaFreeAsync(s->d_partial, s->stream));
    s->d_partial = nullptr;
    s->partial_cap = 0;
    SB_CUDA(cudaMallocAsync(&s->d_partial, need * sizeof(double), s->stream));
    s->partial_cap = need;
    return SB_OK;
}

}  // namespace

void dump_finalize_trace(int rank) {
    unsigned long long h[8];
    if (cudaMemcpyFromSymbol(h, g_fin_trace, sizeof(h)) != cudaSuccess || !h[0]) return;
    fprintf(stderr, "[sb finalize trace] rank %d: %llu finalize launches, %.1f us in the kernel "
            "(%.1f us of it in the all-gather), %.1f us from the end of one to the start of the "
            "next\n", rank, h[0], h[1] / 1e3 / h[0], h[2] / 1e3 / h[0], h[4] / 1e3 / h[0]);
}

// device array of the two pressure buffer pointers lives right after the ctl block
static double *const *pbuf_ptr(sb_sim *s) {
    return reinterpret_cast<double *const *>(reinterpret_cast<char *>(s->d_ctl) + 256);
}

sb_status launch_velocity_bc(sb_sim *s) {
    if (s->bl.n == 0) return SB_OK;
    // slab mode: redundant rows around the owned range so F/G see finished neighbours
    int64_t row0 = s->halo ? s->g.own0 - 2 : 0, row1 = s->halo ? s->g.own1 + 2 : s->g.nxl;
    int nb = (int)((s->bl.n + TPB - 1) / TPB);
    velocity_bc_gather<<<nb, TPB, 0, s->stream>>>(s->g, s->u, s->v, s->cflag, s->bl, row0, row1);
    velocity_bc_scatter<<<nb, TPB, 0, s->stream>>>(s->g, s->u, s->v, s->bl, row0, row1);
    s->launches += 2;
    s->restore_valid = true;
    SB_CUDA(cudaGetLastError());
    return SB_OK;
}

static FgrConsts fgr_consts(const sb_sim *s) {
    const sb_params &p = s->prm;
    const double d[8] = {p.delx * p.delx, p.dely * p.dely, 4.0 * p.delx, 4.0 * p.dely,
                         p.reynolds,      p.delx,          p.dely,       p.delt};
    FgrConsts kc;
    kc.fast = 1;
    for (int i = 0; i < 8; i++) {
        DivC c = make_divc(d[i]);
        kc.d[i] = c.d;
        kc.r[i] = c.r;
        kc.fast &= c.fast;
    }
    kc.delt = p.delt;
    kc.gamma = p.gamma;
    return kc;
}

static sb_status rhs_halo(sb_sim *s) {
    if (!s->slab) return SB_OK;
    // the red-black tiles re-sweep their halo rows and need rhs there: edge rows -> the
    // neighbours' halo rows, then a barrier before the first pass reads them.  (The
    // neighbours left their previous solve long ago: the barriers of the velocity update.)
    sb_status st = slab_put_rows(s, s->rhs, s->lo_rhs, s->hi_rhs, 8);
    if (st) return st;
    return slab_allreduce(s, s->d_scalars, 0, 0);
}

// what = 1: F, G (sb_calculate_f_and_g); 3: F, G and RHS in one pass (the tick)
sb_status launch_fg_rhs(sb_sim *s, int what) {
    const bool with_rhs = (what & 2) != 0;
    // slab mode without the fused rhs: F of the halo row in front is stored for rhs_kernel
    int64_t row0 = (s->halo && !with_rhs) ? s->g.own0 - 1 : s->g.own0, row1 = s->g.own1;
    const sb_params &p = s->prm;
    // performance mode (red-black): reciprocal + FMA arithmetic, bit-identical to the oracle's
    // restatement of it; reference-order mode: the reference's exact divisions
    if (p.sor_mode == SB_SOR_RED_BLACK) {
        dim3 grid((unsigned)((row1 - row0 + FGF_ROWS - 1) / FGF_ROWS),
                  (unsigned)((s->g.NY + FGF_COLS - 1) / FGF_COLS));
        fg_rhs_fast_kernel<<<grid, 32 * FGF_WARPS, 0, s->stream>>>(
            s->g, s->u, s->v, s->cflag, s->f, s->gq, s->rhs, (int)row0, (int)row1,
            make_fg_fast(p.delx, p.dely, p.delt, p.gamma, p.reynolds), 1, with_rhs ? 1 : 0);
    } else {
        dim3 grid((unsigned)((row1 - row0 + FGR_ROWS - 1) / FGR_ROWS),
                  (unsigned)((s->g.NY + FGR_COLS - 1) / FGR_COLS));
        fg_rhs_kernel<<<grid, 32 * FGR_WARPS, 0, s->stream>>>(
            s->g, s->u, s->v, s->cflag, s->f, s->gq, s->rhs, row0, row1, s->g.own0,
            fgr_consts(s), 1, with_rhs ? 1 : 0);
    }
    s->launches++;
    SB_CUDA(cudaGetLastError());
    return with_rhs ? rhs_halo(s) : SB_OK;
}

sb_status launch_fg(sb_sim *s) { return launch_fg_rhs(s, 1); }

sb_status launch_rhs(sb_sim *s) {
    int64_t row0 = s->g.own0, row1 = s->g.own1;
    rhs_kernel<<<row_grid(s->g, row1 - row0, s->g.NY), TPB, 0, s->stream>>>(
        s->g, s->f, s->gq, s->rhs, row0, row1, make_divc(s->prm.delx), make_divc(s->prm.dely),
        make_divc(s->prm.delt), s->prm.sor_mode == SB_SOR_RED_BLACK ? 1 : 0,
        make_fg_fast(s->prm.delx, s->prm.dely, s->prm.delt, s->prm.gamma, s->prm.reynolds));
    s->launches++;
    SB_CUDA(cudaGetLastError());
    return rhs_halo(s);
}

sb_status launch_pressure_bc(sb_sim *s, int guarded) {
    if (s->bl.n == 0) return SB_OK;
    int nb = (int)((s->bl.n + TPB - 1) / TPB);
    pressure_bc_kernel<<<nb, TPB, 0, s->stream>>>(s->g, pbuf_ptr(s), s->d_ctl, guarded, s->bl);
    s->launches++;
    SB_CUDA(cudaGetLastError());
    return SB_OK;
}

sb_status launch_norm_partials(sb_sim *s, int guarded, int *nblocks) {
    int64_t rows = s->g.own1 - s->g.own0;
    unsigned gy = (unsigned)((rows + NORM_ROWS_PER_BLOCK - 1) / NORM_ROWS_PER_BLOCK);
    unsigned gx = (unsigned)((s->g.NY + 4 * TPB - 1) / (4 * TPB));
    if (gx < 1) gx = 1;
    sb_status st = ensure_partial(s, (size_t)gx * gy * 4 + 64);
    if (st) return st;
    norm_partial_kernel<<<dim3(gy, gx), TPB, 0, s->stream>>>(
        s->g, pbuf_ptr(s), s->d_ctl, guarded, s->rhs, s->d_partial,
        make_divc(s->prm.delx * s->prm.delx), make_divc(s->prm.dely * s->prm.dely));
    s->launches++;
    SB_CUDA(cudaGetLastError());
    *nblocks = (int)(gx * gy);
    return SB_OK;
}

sb_status launch_sor_finalize(sb_sim *s, int nparts, double initial_norm, double eps2,
                              int test_exit, double *norm_hist) {
    sor_finalize_kernel<<<1, 1024, 0, s->stream>>>(s->d_ctl, s->d_partial, nparts,
                                                   s->fluid_cells, initial_norm, eps2,
                                                   test_exit, norm_hist, s->link);
    s->launches++;
    SB_CUDA(cudaGetLastError());
    return SB_OK;
}

sb_status reduce_norm(sb_sim *s, int nparts, double *out) {
    sum_partials_kernel<<<1, 1024, 0, s->stream>>>(s->d_partial, nparts, s->fluid_cells,
                                                   s->d_scalars, s->slab ? 0 : 1);
    s->launches++;
    sb_status st = slab_allreduce(s, s->d_scalars, 1, XR_SUM);
    if (st) return st;
    SB_CUDA(cudaMemcpyAsync(s->h_scalars, s->d_scalars, sizeof(double), cudaMemcpyDeviceToHost,
                            s->stream));
    SB_CUDA(cudaStreamSynchronize(s->stream));
    *out = s->slab ? s->h_scalars[0] / s->fluid_cells : s->h_scalars[0];
    return SB_OK;
}

// stride 4: out[0..3]; stride 8: out[0..5] (the fused pressure range in out[4..5])
static sb_status finish_ranges(sb_sim *s, int64_t nblk, double *out, int stride = 4) {
    range_final_kernel<<<1, 1024, 0, s->stream>>>(s->d_partial, nblk, s->d_scalars, stride);
    s->launches++;
    const int n = stride == 8 ? 6 : 4;
    sb_status st = slab_allreduce(s, s->d_scalars, n,
                                  XR_MIN | (XR_MAX << 2) | (XR_MAX << 4) | (XR_MAX << 6) |
                                      (XR_MIN << 8) | (XR_MAX << 10));
    if (st) return st;
    SB_CUDA(cudaMemcpyAsync(s->h_scalars, s->d_scalars, n * sizeof(double),
                            cudaMemcpyDeviceToHost, s->stream));
    SB_CUDA(cudaStreamSynchronize(s->stream));
    for (int i = 0; i < n; i++) out[i] = s->h_scalars[i];
    return SB_OK;
}

// with_prange: also grid.calculate_pressure_range() (the caller skipped it after a capped solve)
sb_status launch_adapt_uv(sb_sim *s, int with_prange) {
    const int64_t rows = (s->g.own1 - s->g.own0 + RANGE_RPB - 1) / RANGE_RPB;  // block rows
    unsigned gx = (unsigned)((s->g.NY + 4 * TPB - 1) / (4 * TPB));
    if (gx < 1) gx = 1;
    sb_status st = ensure_partial(s, (size_t)gx * rows * 8 + 64);
    if (st) return st;
    if (with_prange)
        adapt_uv_kernel<true><<<dim3((unsigned)rows, gx), TPB, 0, s->stream>>>(
            s->g, s->p[s->cur], s->f, s->gq, s->cflag, s->u, s->v, s->d_partial, s->prm.delt,
            s->prm.delx, s->prm.dely);
    else
        adapt_uv_kernel<false><<<dim3((unsigned)rows, gx), TPB, 0, s->stream>>>(
            s->g, s->p[s->cur], s->f, s->gq, s->cflag, s->u, s->v, s->d_partial, s->prm.delt,
            s->prm.delx, s->prm.dely);
    s->launches++;
    if (s->bl.n && s->restore_valid) {   // u_v_restore is empty until the first velocity BC
        int nb = (int)((s->bl.n + TPB - 1) / TPB);
        restore_uv_kernel<<<nb, TPB, 0, s->stream>>>(s->g, s->u, s->v, s->bl);
        s->launches++;
    }
    SB_CUDA(cudaGetLastError());
    if (s->slab) {
        // new u, v of my edge rows -> the neighbours' halo rows.  The barrier in front keeps a
        // slow neighbour's F/G pass from reading rows of the NEXT step; the range reduction
        // below is the barrier behind the puts.
        if ((st = slab_allreduce(s, s->d_scalars, 0, 0))) return st;
        if ((st = slab_put_rows(s, s->u, s->lo_u, s->hi_u, 8))) return st;
        if ((st = slab_put_rows(s, s->v, s->lo_v, s->hi_v, 8))) return st;
    }
    double out[6];
    st = finish_ranges(s, (int64_t)gx * rows, out, with_prange ? 8 : 4);
    if (st) return st;
    // sqrt of the folded min / max (src/grid/mod.rs:267)
    s->speed_range[0] = sqrt(out[0]);
    s->speed_range[1] = sqrt(out[1]);
    s->umax = out[2];
    s->vmax = out[3];
    s->uvmax_valid = true;
    if (with_prange) {
        s->pressure_range[0] = out[4];
        s->pressure_range[1] = out[5];
    }
    return SB_OK;
}

static sb_status launch_range(sb_sim *s, int mode, double out[4]) {
    int64_t rows = s->g.own1 - s->g.own0;
    unsigned gx = (unsigned)((s->g.NY + 4 * TPB - 1) / (4 * TPB));
    if (gx < 1) gx = 1;
    sb_status st = ensure_partial(s, (size_t)gx * rows * 4 + 64);
    if (st) return st;
    range_kernel<<<dim3((unsigned)rows, gx), TPB, 0, s->stream>>>(
        s->g, mode == 0 ? s->p[s->cur] : s->u, s->v, s->cflag, s->d_partial, mode);
    s->launches++;
    SB_CUDA(cudaGetLastError());
    return finish_ranges(s, (int64_t)gx * rows, out);
}

sb_status launch_pressure_range(sb_sim *s) {
    double out[4];
    sb_status st = launch_range(s, 0, out);
    if (st) return st;
    s->pressure_range[0] = out[0];
    s->pressure_range[1] = out[1];
    return SB_OK;
}

sb_status launch_speed_range(sb_sim *s) {
    double out[4];
    sb_status st = launch_range(s, 1, out);
    if (st) return st;
    s->speed_range[0] = sqrt(out[0]);
    s->speed_range[1] = sqrt(out[1]);
    s->umax = out[2];
    s->vmax = out[3];
    s->uvmax_valid = true;
    return SB_OK;
}

sb_status launch_cellop(int op, const double *u9, const double *v9, const double *scal,
                        double *out) {
    double h_in[23] = {0};
    for (int i = 0; i < 9; i++) {
        h_in[i] = u9 ? u9[i] : 0.0;
        h_in[9 + i] = v9 ? v9[i] : 0.0;
    }
    for (int i = 0; i < 5; i++) h_in[18 + i] = scal[i];
    double *d = nullptr;
    SB_CUDA(cudaMalloc(&d, 24 * sizeof(double)));
    SB_CUDA(cudaMemcpy(d, h_in, 23 * sizeof(double), cudaMemcpyHostToDevice));
    cellop_kernel<<<1, 1>>>(op, d, d + 23);
    cudaError_t e = cudaMemcpy(out, d + 23, sizeof(double), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) {
        set_error(std::string("cellop: ") + cudaGetErrorString(e));
        return SB_CUDA_ERROR;
    }
    return SB_OK;
}

// force-load this file's kernels (CUDA loads lazily by default, and a first launch that has
// to load code synchronises the context -- fatal while a peer slab of the same process spins
// in an all-gather on the same GPU)
void preload_stages() {
    cudaFuncAttributes a;
    cudaFuncGetAttributes(&a, velocity_bc_gather);
    cudaFuncGetAttributes(&a, velocity_bc_scatter);
    cudaFuncGetAttributes(&a, fg_rhs_kernel);
    cudaFuncGetAttributes(&a, fg_rhs_fast_kernel);
    cudaFuncGetAttributes(&a, rhs_kernel);
    cudaFuncGetAttributes(&a, pressure_bc_kernel);
    cudaFuncGetAttributes(&a, norm_partial_kernel);
    cudaFuncGetAttributes(&a, sor_finalize_kernel);
    cudaFuncGetAttributes(&a, sum_partials_kernel);
    cudaFuncGetAttributes(&a, adapt_uv_kernel<false>);
    cudaFuncGetAttributes(&a, adapt_uv_kernel<true>);
    cudaFuncGetAttributes(&a, restore_uv_kernel);
    cudaFuncGetAttributes(&a, range_kernel);
    cudaFuncGetAttributes(&a, range_final_kernel);
    cudaFuncGetAttributes(&a, cellop_kernel);
    cudaGetLastError();
}

}  // namespace sb
