// sor_rb.cu -- K4b: performance-mode SOR pass.  Red-black ordering, pressure BC, residual
// norm and up to 4 sweeps fused into ONE pass over HBM per launch (temporal blocking).
//
// Replaces, per launch, T iterations of the loop body of Simulation::solve_sor
// (/root/reference/src/simulation.rs:250-281): copy_pressure_to_boundaries
// (src/grid/mod.rs:343-412), the sweep (:253-274, here as a red half-sweep over
// (x+y) even then a black half-sweep) and calculate_norm_squared (:216-227).
//
// One CTA owns a TXR x TW tile (x rows, y columns; y contiguous).  p and rhs are staged
// into shared memory by two TMA box loads (out-of-grid cells arrive as zeros); the u8 cell
// flags are read with plain 16-bit loads while the TMA is in flight (a TMA box must start
// on a 16-byte boundary, which a halo of 1-byte cells cannot honour) and condensed into
// per-thread bit masks.  Each thread owns two adjacent columns of RPT rows and keeps their
// pressures IN REGISTERS for the whole pass; shared memory only serves the exchange with
// the neighbouring threads (one foreign column neighbour per update, two halo rows per
// half-sweep) and the rhs reads.  Row parity is a template parameter and the row loops are
// fully unrolled, so which cell of a pair has which colour is known at compile time.
//
// Shared-memory layout: after the TMA has landed a tile in natural order, every warp
// re-lays its 64-column half rows in place as [32 first cells | 32 second cells] of its 32
// column pairs.  Each lane then reads / writes single doubles at an 8-byte stride across
// the warp: the neighbour, rhs and store accesses of a half-sweep are bank-conflict free
// (in natural order every one of them is a 2-way conflict: 8 bytes at a 16-byte stride).
//
// Halo: the tile carries h = 2T+2 cells.  Boundary cells on the outermost tile ring cannot
// take their pressure BC (their fluid neighbour may lie outside the tile), so after sweep k
// the fluid values at distance >= 2k+1 from the tile edge are exact; T sweeps leave the
// inner (TXR-2h) x (TW-2h) region exact together with its residuals (which read neighbours
// at distance h-1 = 2T+1) for all T sweeps.  That region is written to the other pressure
// buffer (ping-pong); one partial sum of squared residuals per sweep and tile goes to the
// finalize kernel.
//
// Residuals ride along: a black cell's residual is taken right after its update (all its
// neighbours are final); a red cell's residual of sweep k is evaluated inside the red
// half-sweep of sweep k+1, where the same stencil sum t is needed anyway and none of its
// inputs has changed.  Only red cells next to a non-fluid cell (whose boundary pressures
// are refreshed in between), non-fluid cells, and the last sweep of a pass need an explicit
// residual evaluation.  The values are bit-identical either way.
//
// HBM traffic per launch and cell: 8 (p in) + 8 (rhs) + 1 (flag) + 8 (p out) = 25 B for
// T sweeps (halo re-reads are served by L2), i.e. 25/T B per cell-sweep.
//
// Arithmetic (identical in oracle/stroemung_oracle.c, SO_SOR_RED_BLACK):
//     t     = fma(1/dx^2, pE+pW, fma(1/dy^2, pS+pN, -rhs))
//     p_new = fma(mid, t, (1-w)*p)            mid = w / (2/dx^2 + 2/dy^2)
//     r     = fma(-(2/dx^2 + 2/dy^2), p, t)   residual, all interior cells
#include <stdlib.h>

#include "sor_rb.cuh"

namespace sb {

namespace {

constexpr int TXR = RB_TXR;             // tile rows
constexpr int TW = RB_TW;               // tile columns
#ifndef SB_RB_NTHR
#define SB_RB_NTHR 256
#endif
constexpr int NTHR = SB_RB_NTHR;
constexpr int TCOLS = TW / 2;           // thread columns (2 cells each)
constexpr int TROWS = NTHR / TCOLS;     // thread rows
constexpr int RPT = TXR / TROWS;        // rows per thread
constexpr int TMAX = RB_TMAX;
constexpr int TILE = TXR * TW;
constexpr int NWARP = NTHR / 32;
constexpr int SMEM_PAD = 128;  // in front of the p tile: the ring cells' out-of-tile reads
constexpr size_t SMEM_BYTES = SMEM_PAD + (size_t)TILE * 17 + 64 + NWARP * TMAX * sizeof(double);
static_assert(TXR % TROWS == 0 && RPT % 2 == 0, "rows must split evenly, even per thread");
static_assert(2 * RPT <= 32, "cell masks are 32-bit");

// row-slab mode: the neighbours' pressure buffers (nullptr at the chain ends / single GPU)
// and the row of THEIR array that receives my first (lo) / last (hi) H owned rows
struct RbPeers {
    double *lo_p[2], *hi_p[2];
    int64_t lo_row0, hi_row0;
    int H;
};

__device__ __forceinline__ double2 lds2(const double *sp, int idx) {
    return *reinterpret_cast<const double2 *>(sp + idx);
}

// index of tile cell (r, col) in the split layout: pair t = col/2 sits in the 64-wide half
// row t/32, at lane t%32 of its first-cell or second-cell block
__device__ __forceinline__ int cidx(int r, int col) {
    const int t = col >> 1;
    return r * TW + ((t >> 5) << 6) + ((col & 1) << 5) + (t & 31);
}

// pressure BC of one boundary cell from its fluid neighbours (src/grid/mod.rs:351-399)
__device__ __forceinline__ double bc_value(const double *sp, int r, int col, int edge) {
    const double n = sp[cidx(r, col - 1)], s = sp[cidx(r, col + 1)];
    const double e = sp[cidx(r + 1, col)], w = sp[cidx(r - 1, col)];
    switch (edge) {
    case SB_EDGE_N: return n;
    case SB_EDGE_NE: return (n + e) / 2.0;
    case SB_EDGE_E: return e;
    case SB_EDGE_SE: return (s + e) / 2.0;
    case SB_EDGE_S: return s;
    case SB_EDGE_SW: return (s + w) / 2.0;
    case SB_EDGE_W: return w;
    case SB_EDGE_NW: return (n + w) / 2.0;
    default: return sp[cidx(r, col)];
    }
}

// where a thread's cells live in the split layout
struct Thr {
    int bx;    // index of the first cell of the thread's pair in its first row
    int offL;  // from there to the left foreign neighbour (second cell of the pair before)
    int offR;  // ... to the right foreign neighbour (first cell of the pair after)
    int r_begin;
};

__device__ __forceinline__ double2 lds_pair(const double *sp, int bx) {
    return make_double2(sp[bx], sp[bx + 32]);
}

// both cell bits of the thread's rows i with r_begin + i in [lo, hi)
__device__ __forceinline__ uint32_t row_bits(int lo, int hi, int r_begin) {
    const int a = min(max(lo - r_begin, 0), RPT), b = min(max(hi - r_begin, 0), RPT);
    if (b <= a) return 0u;
    return ((1u << (2 * b)) - 1u) & ~((1u << (2 * a)) - 1u);
}

// per-thread cell masks; bit 2*i + e is cell (row r_begin + i, column col0 + e)
struct Masks {
    uint32_t upd;   // fluid and interior of the grid: swept
    uint32_t cnt;   // in the tile's exact inner region, interior, owned: counts in the norm
    uint32_t exp_;  // counted cells whose residual needs an explicit evaluation every sweep
    uint32_t bc;    // boundary cells with an edge class, not on the tile ring: take the BC
};

// bits of the cells with colour 0 (red) when the thread's first row has parity PAR
template <int PAR>
__host__ __device__ constexpr uint32_t red_mask() {
    uint32_t m = 0;
    for (int i = 0; i < RPT; i++) m |= 1u << (2 * i + ((PAR ^ i) & 1));
    return m;
}

// stencil sum t of cell (row i, component sel) from registers + one shared-memory neighbour
#define SB_T_OF_CELL(i, sel)                                                            \
    const double2 &Pm_ = (i) == 0 ? up : P[(i) == 0 ? 0 : (i)-1];                       \
    const double2 &Pn_ = (i) == RPT - 1 ? dn : P[(i) == RPT - 1 ? RPT - 1 : (i) + 1];   \
    const int bx_ = th.bx + (i)*TW;                                                     \
    const double pE_ = (sel) ? Pn_.y : Pn_.x, pW_ = (sel) ? Pm_.y : Pm_.x;              \
    const double pS_ = (sel) ? sp[bx_ + th.offR] : P[i].y;                              \
    const double pN_ = (sel) ? P[i].x : sp[bx_ + th.offL];                              \
    const double t_ = fma(k.rdx2, pE_ + pW_, fma(k.rdy2, pS_ + pN_, -sr[bx_ + (sel)*32]));

// bits of the cells with colour COLOUR when the thread's first row has parity PAR
template <int PAR, int COLOUR>
__host__ __device__ constexpr uint32_t colour_mask() {
    uint32_t m = 0;
    for (int i = 0; i < RPT; i++) m |= 1u << (2 * i + ((PAR ^ i ^ COLOUR) & 1));
    return m;
}

// one colour of one sweep.  COLOUR 0 (red): cells in m_lag also contribute the residual of
// the PREVIOUS sweep to acc_prev.  COLOUR 1 (black): counted cells contribute this sweep's
// residual to acc_cur.
//
// Two code paths, chosen per warp.  Fast path: every lane sweeps all its cells of the colour
// (the common case away from walls and obstacles) -- straight-line code, loads of GROUP rows
// issued ahead of the arithmetic, so the rows' independent dependency chains (LDS ->
// DADD -> 3 DFMA -> STS, ~60 cycles each) overlap.  Slow path: one branch per cell.
// Cells on the tile ring are swept like any other: their out-of-tile neighbour reads land in
// the padding / adjacent row of the shared-memory tile and only produce values inside the
// (already inexact) halo.
template <int PAR, int COLOUR>
__device__ __forceinline__ void half_sweep(double2 (&P)[RPT], double *sp, const double *sr,
                                           const Thr &th, const Masks &m, uint32_t m_lag,
                                           const RbConsts &k, double &acc_prev, double &acc_cur) {
    const double2 up = th.r_begin > 0 ? lds_pair(sp, th.bx - TW) : make_double2(0.0, 0.0);
    const double2 dn =
        th.r_begin + RPT < TXR ? lds_pair(sp, th.bx + RPT * TW) : make_double2(0.0, 0.0);
    constexpr uint32_t cm = colour_mask<PAR, COLOUR>();
    if (__all_sync(0xffffffffu, (m.upd & cm) == cm)) {
        constexpr int GROUP = RPT % 4 == 0 ? 4 : 3;
#pragma unroll
        for (int g0 = 0; g0 < RPT; g0 += GROUP) {
            double fo[GROUP], rh[GROUP];
#pragma unroll
            for (int j = 0; j < GROUP; j++) {
                const int i = g0 + j;
                const int sel = (PAR ^ i ^ COLOUR) & 1;
                fo[j] = sp[th.bx + i * TW + (sel ? th.offR : th.offL)];
                rh[j] = sr[th.bx + i * TW + sel * 32];
            }
#pragma unroll
            for (int j = 0; j < GROUP; j++) {
                const int i = g0 + j;
                const int sel = (PAR ^ i ^ COLOUR) & 1;
                const uint32_t bit = 1u << (2 * i + sel);
                const double2 &Pm = i == 0 ? up : P[i == 0 ? 0 : i - 1];
                const double2 &Pn = i == RPT - 1 ? dn : P[i == RPT - 1 ? RPT - 1 : i + 1];
                const double pE = sel ? Pn.y : Pn.x, pW = sel ? Pm.y : Pm.x;
                const double pS = sel ? fo[j] : P[i].y;
                const double pN = sel ? P[i].x : fo[j];
                const double pold = sel ? P[i].y : P[i].x;
                const double t = fma(k.rdx2, pE + pW, fma(k.rdy2, pS + pN, -rh[j]));
                if (COLOUR == 0) {
                    const double rr = fma(-k.diag, pold, t);
                    if (m_lag & bit) acc_prev = fma(rr, rr, acc_prev);
                }
                const double pnew = fma(k.mid, t, k.omw * pold);
                if (sel) P[i].y = pnew; else P[i].x = pnew;
                sp[th.bx + i * TW + sel * 32] = pnew;
                if (COLOUR == 1) {
                    const double rr = fma(-k.diag, pnew, t);
                    if (m.cnt & bit) acc_cur = fma(rr, rr, acc_cur);
                }
            }
        }
        return;
    }
#pragma unroll
    for (int i = 0; i < RPT; i++) {
        const int sel = (PAR ^ i ^ COLOUR) & 1;
        const uint32_t bit = 1u << (2 * i + sel);
        if (m.upd & bit) {
            SB_T_OF_CELL(i, sel)
            const double pold = sel ? P[i].y : P[i].x;
            if (COLOUR == 0 && (m_lag & bit)) {
                const double rr = fma(-k.diag, pold, t_);
                acc_prev = fma(rr, rr, acc_prev);
            }
            const double pnew = fma(k.mid, t_, k.omw * pold);
            if (sel) P[i].y = pnew; else P[i].x = pnew;
            sp[bx_ + sel * 32] = pnew;
            if (COLOUR == 1 && (m.cnt & bit)) {
                const double rr = fma(-k.diag, pnew, t_);
                acc_cur = fma(rr, rr, acc_cur);
            }
        }
    }
}

// explicit residuals of the cells in `mask` (current field) into acc.  `uniform_red`: the
// caller asks for all counted red cells (last sweep of a pass); if the whole warp sweeps all
// its red cells the straight-line path is taken.
template <int PAR>
__device__ __forceinline__ void explicit_norm(const double2 (&P)[RPT], const double *sp,
                                              const double *sr, const Thr &th, uint32_t mask,
                                              uint32_t m_upd, bool uniform_red,
                                              const RbConsts &k, double &acc) {
    constexpr uint32_t cm = colour_mask<PAR, 0>();
    const bool fast = uniform_red && (m_upd & cm) == cm && (mask & ~cm) == 0;
    if (__all_sync(0xffffffffu, fast)) {
        const double2 up = th.r_begin > 0 ? lds_pair(sp, th.bx - TW) : make_double2(0.0, 0.0);
        const double2 dn =
            th.r_begin + RPT < TXR ? lds_pair(sp, th.bx + RPT * TW) : make_double2(0.0, 0.0);
        constexpr int GROUP = RPT % 4 == 0 ? 4 : 3;
#pragma unroll
        for (int g0 = 0; g0 < RPT; g0 += GROUP) {
            double fo[GROUP], rh[GROUP];
#pragma unroll
            for (int j = 0; j < GROUP; j++) {
                const int i = g0 + j;
                const int sel = (PAR ^ i) & 1;
                fo[j] = sp[th.bx + i * TW + (sel ? th.offR : th.offL)];
                rh[j] = sr[th.bx + i * TW + sel * 32];
            }
#pragma unroll
            for (int j = 0; j < GROUP; j++) {
                const int i = g0 + j;
                const int sel = (PAR ^ i) & 1;
                const double2 &Pm = i == 0 ? up : P[i == 0 ? 0 : i - 1];
                const double2 &Pn = i == RPT - 1 ? dn : P[i == RPT - 1 ? RPT - 1 : i + 1];
                const double pE = sel ? Pn.y : Pn.x, pW = sel ? Pm.y : Pm.x;
                const double pS = sel ? fo[j] : P[i].y;
                const double pN = sel ? P[i].x : fo[j];
                const double t = fma(k.rdx2, pE + pW, fma(k.rdy2, pS + pN, -rh[j]));
                const double rr = fma(-k.diag, sel ? P[i].y : P[i].x, t);
                if (mask & (1u << (2 * i + sel))) acc = fma(rr, rr, acc);
            }
        }
        return;
    }
    if (mask == 0) return;
    const double2 up = th.r_begin > 0 ? lds_pair(sp, th.bx - TW) : make_double2(0.0, 0.0);
    const double2 dn =
        th.r_begin + RPT < TXR ? lds_pair(sp, th.bx + RPT * TW) : make_double2(0.0, 0.0);
#pragma unroll
    for (int i = 0; i < RPT; i++) {
#pragma unroll
        for (int sel = 0; sel < 2; sel++) {
            if (mask & (1u << (2 * i + sel))) {
                SB_T_OF_CELL(i, sel)
                const double rr = fma(-k.diag, sel ? P[i].y : P[i].x, t_);
                acc = fma(rr, rr, acc);
            }
        }
    }
}

__device__ __forceinline__ double warp_sum(double v) {
    for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// SLAB: this handle is one row slab of a multi-GPU run (own rows inside halo rows, neighbours
// to feed).  The single-GPU instantiation carries none of that code.
template <int PAR, bool SLAB>
__global__ void __launch_bounds__(NTHR, 2)
sor_rb_kernel(const __grid_constant__ CUtensorMap tm_p0, const __grid_constant__ CUtensorMap tm_p1,
              const __grid_constant__ CUtensorMap tm_rhs, const uint8_t *__restrict__ cflag,
              Geom g, double *const *__restrict__ pbuf, const SorCtl *__restrict__ ctl,
              double *__restrict__ partial, int tiles_y, int part_stride, int h, RbConsts k,
              int norm_only, RbPeers peers, const int32_t *__restrict__ tile_list) {
    // norm_only: no sweeps, no write-back; partial[tile] = sum of squared residuals of the
    // current field (calculate_norm_squared on its own, src/simulation.rs:216-227)
    const int T = norm_only ? 0 : ctl->active_T;
    if (!norm_only && T == 0) return;
    const int src = ctl->src;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *sp = reinterpret_cast<double *>(smem_raw + SMEM_PAD);
    double *sr = sp + TILE;
    uint8_t *sf = reinterpret_cast<uint8_t *>(sr + TILE);   // edge class of BC cells
    uint64_t *bar = reinterpret_cast<uint64_t *>(sf + TILE);
    double *sred = reinterpret_cast<double *>(sf + TILE + 64);  // [TMAX][NWARP]

    const int BX = TXR - 2 * h, BY = TW - 2 * h;  // h is even: owned columns 16-byte aligned
    // tile_list: the tiles this launch covers (nullptr = all of the lattice)
    const int tile = tile_list ? tile_list[blockIdx.x] : (int)blockIdx.x;
    const int tile_i = tile / tiles_y, tile_j = tile - tile_i * tiles_y;
    const int tx0 = (SLAB ? (int)g.own0 : 0) + tile_i * BX - h;  // local row of tile row 0 (even)
    const int ty0 = tile_j * BY - h;   // column of tile column 0 (even)

    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, (uint32_t)(TILE * 16));
        if (src) tma_load_2d(sp, &tm_p1, ty0, tx0, bar);
        else tma_load_2d(sp, &tm_p0, ty0, tx0, bar);
        tma_load_2d(sr, &tm_rhs, ty0, tx0, bar);
    }

    const int tc = threadIdx.x % TCOLS, tr = threadIdx.x / TCOLS;
    const int r_begin = tr * RPT, col0 = 2 * tc;
    const int base = r_begin * TW + col0;
    const int gy0 = ty0 + col0;  // global column of the thread's first cell (even)

    // ---- cell flags of this thread's 2 x RPT cells -> bit masks (while the TMA flies) ----
    // Geometry first (32-bit, mostly uniform per thread row): which of the thread's rows /
    // columns are swept, counted, BC-able, stored.  Then one 16-bit flag load per row.
    Masks m;
    uint32_t m_store;  // rows (bit 2i) this thread writes back
    {
        const int nxl = (int)g.nxl, NYi = (int)g.NY;
        // tile-row ranges [lo, hi): rows present in this slab / interior rows of the grid /
        // rows owned by this slab
        const int in_lo = max(0, -tx0), in_hi = min(TXR, nxl - tx0);
        const int64_t gxt = g.gx0 + tx0;  // global x of tile row 0
        const int int_lo = (int)max((int64_t)0, 1 - gxt);
        const int int_hi = (int)max((int64_t)0, min((int64_t)TXR, g.NX - 1 - gxt));
        const int own_lo = max(0, (int)g.own0 - tx0), own_hi = min(TXR, (int)g.own1 - tx0);
        const uint32_t rows_upd = row_bits(int_lo, int_hi, r_begin);
        const uint32_t rows_cnt =
            row_bits(max(max(int_lo, own_lo), h), min(min(int_hi, own_hi), TXR - h), r_begin);
        const uint32_t rows_bc = row_bits(max(in_lo, 1), min(in_hi, TXR - 1), r_begin);
        // stored: the exact inner rows this slab owns (its halo rows belong to the neighbours,
        // who write them from their own epilogue below)
        const uint32_t rows_st =
            SLAB ? row_bits(max(max(in_lo, own_lo), h), min(min(in_hi, own_hi), TXR - h), r_begin)
                 : row_bits(max(in_lo, h), min(in_hi, TXR - h), r_begin);
        // column properties of the two cells -> 0x555555 / 0xAAAAAA patterns
        uint32_t cols_upd = 0, cols_cnt = 0, cols_bc = 0;
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const int col = col0 + e, gy = gy0 + e;
            const uint32_t pat = 0x55555555u << e;
            const bool interior = gy >= 1 && gy <= NYi - 2;
            const bool tile_col = col >= 1 && col <= TW - 2;
            if (interior) cols_upd |= pat;
            if (interior && col >= h && col < TW - h) cols_cnt |= pat;
            if (tile_col) cols_bc |= pat;
        }
        const bool col_in = gy0 >= 0 && gy0 + 1 < (int)g.pitch;
        const uint8_t *fp = cflag + (int64_t)(tx0 + r_begin) * g.pitch + gy0;
        uint16_t ff[RPT];
#pragma unroll
        for (int i = 0; i < RPT; i++) {
            const int r = r_begin + i;
            ff[i] = 0;
            if (col_in && r >= in_lo && r < in_hi)
                ff[i] = *reinterpret_cast<const uint16_t *>(fp + (int64_t)i * g.pitch);
        }
        uint32_t f_valid = 0, f_fluid = 0, f_near = 0, f_edge = 0;
#pragma unroll
        for (int i = 0; i < RPT; i++) {
            const uint32_t w = ff[i];
            // per byte: valid = bit 7, fluid = (b & 0x87) == 0x80, near = bit 3, edge = bits 3-6
            const uint32_t valid = ((w >> 7) & 1u) | ((w >> 14) & 2u);
            const uint32_t kindz = (((w & 0x0007u) == 0) ? 1u : 0u) | (((w & 0x0700u) == 0) ? 2u : 0u);
            const uint32_t fluid = valid & kindz;
            const uint32_t near = ((w >> 3) & 1u) | ((w >> 10) & 2u);
            const uint32_t edge = (((w & 0x0078u) != 0) ? 1u : 0u) | (((w & 0x7800u) != 0) ? 2u : 0u);
            f_valid |= valid << (2 * i);
            f_fluid |= fluid << (2 * i);
            f_near |= (near & fluid) << (2 * i);
            f_edge |= (edge & valid & ~fluid) << (2 * i);
        }
        m.upd = f_fluid & rows_upd & cols_upd;
        m.cnt = f_valid & rows_cnt & cols_cnt;
        // explicit residual: non-fluid cells, and red fluid cells next to a non-fluid cell
        // (CF_NEAR); black fluid cells never need it
        m.exp_ = m.cnt & (~f_fluid | (f_near & red_mask<PAR>()));
        m.bc = f_edge & rows_bc & cols_bc;
        m_store = rows_st & 0x55555555u;
        if (m.bc) {  // edge classes of the BC cells, for bc_value()
#pragma unroll
            for (int i = 0; i < RPT; i++) {
                if ((m.bc >> (2 * i)) & 3u) {
                    const uint32_t w = ff[i];
                    const uint16_t edges = (uint16_t)(((w >> 3) & 0x0Fu) | (((w >> 11) & 0x0Fu) << 8));
                    *reinterpret_cast<uint16_t *>(sf + (r_begin + i) * TW + col0) = edges;
                }
            }
        }
    }
    // block-uniform: does any thread have BC cells / explicit-residual cells?
    const int has_bc = __syncthreads_or(m.bc != 0);
    const int has_exp = __syncthreads_or(m.exp_ != 0);
    const uint32_t m_red = red_mask<PAR>();
    const uint32_t m_lag_all = m.cnt & m_red & ~m.exp_;  // red residuals taken one sweep late

    mbar_wait(bar, 0);

    // own pressures -> registers (for the whole pass), then re-lay the warp's half rows of p
    // and rhs in place: [32 first cells | 32 second cells] (see the header comment).  A warp
    // only touches its own 64-column half rows here, so __syncwarp() orders reads and writes.
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    Thr th;
    th.r_begin = r_begin;
    th.bx = r_begin * TW + ((tc >> 5) << 6) + lane;
    th.offL = lane > 0 ? 31 : -1;
    th.offR = lane < 31 ? 1 : 33;
    double2 P[RPT];
#pragma unroll
    for (int i = 0; i < RPT; i++) P[i] = lds2(sp, base + i * TW);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < RPT; i++) {
        sp[th.bx + i * TW] = P[i].x;
        sp[th.bx + i * TW + 32] = P[i].y;
    }
#pragma unroll
    for (int g0 = 0; g0 < RPT; g0 += 4) {
        double2 rr[4];
#pragma unroll
        for (int j = 0; j < 4; j++) rr[j] = lds2(sr, base + (g0 + j) * TW);
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; j++) {
            sr[th.bx + (g0 + j) * TW] = rr[j].x;
            sr[th.bx + (g0 + j) * TW + 32] = rr[j].y;
        }
    }
    __syncthreads();

    double acc_prev = 0.0, acc_cur = 0.0;

    if (norm_only) {
        explicit_norm<PAR>(P, sp, sr, th, m.cnt, m.upd, false, k, acc_cur);
        const double v = warp_sum(acc_cur);
        if (lane == 0) sred[warp] = v;
    }

    for (int it = 0; it < T; it++) {
        // pressure BC: boundary cells take the (average of the) fluid neighbour(s)
        if (has_bc) {
            if (m.bc) {
#pragma unroll
                for (int i = 0; i < RPT; i++) {
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        if (m.bc & (1u << (2 * i + e))) {
                            const int r = r_begin + i, col = col0 + e;
                            const double v = bc_value(sp, r, col, sf[r * TW + col]);
                            sp[th.bx + i * TW + e * 32] = v;
                            if (e) P[i].y = v; else P[i].x = v;
                        }
                    }
                }
            }
            __syncthreads();
        }
        half_sweep<PAR, 0>(P, sp, sr, th, m, it > 0 ? m_lag_all : 0u, k, acc_prev, acc_cur);
        if (it > 0) {  // sweep it-1 is now complete in acc_prev
            const double v = warp_sum(acc_prev);
            if (lane == 0) sred[(it - 1) * NWARP + warp] = v;
        }
        __syncthreads();
        acc_prev = 0.0;
        half_sweep<PAR, 1>(P, sp, sr, th, m, 0u, k, acc_prev, acc_cur);
        __syncthreads();
        // explicit residuals of this sweep: the rare cells every sweep, all red cells after
        // the last sweep of the pass
        const bool last = it == T - 1;
        if (has_exp || last) {
            explicit_norm<PAR>(P, sp, sr, th, last ? (m.exp_ | (m.cnt & m_red)) : m.exp_, m.upd,
                               last, k, acc_cur);
            // before the next BC / red half-sweep overwrites cells these residuals read
            if (!last) __syncthreads();
        }
        acc_prev = acc_cur;
        acc_cur = 0.0;
    }
    if (T > 0) {
        const double v = warp_sum(acc_prev);
        if (lane == 0) sred[(T - 1) * NWARP + warp] = v;
    }

    // write the exact inner region to the other buffer, straight from registers
    if (!norm_only && m_store && col0 >= h && col0 < TW - h && gy0 < (int)g.NY) {
        double *pout = pbuf[src ^ 1] + (int64_t)(tx0 + r_begin) * g.pitch + gy0;
        const bool pair = gy0 + 1 < (int)g.NY;
#pragma unroll
        for (int i = 0; i < RPT; i++) {
            if (m_store & (1u << (2 * i))) {
                double *dst = pout + (int64_t)i * g.pitch;
                if (pair) *reinterpret_cast<double2 *>(dst) = P[i];
                else dst[0] = P[i].x;
            }
        }
        // halo exchange fused into the pass: rows within H of a slab edge also go straight
        // into the neighbour's halo rows of ITS target buffer (P2P stores over NVLink); the
        // all-gather of the finalize kernel that follows is the release/acquire point.  Only
        // the first and last tile rows of a slab ever get here (block-uniform test).
        if (SLAB) {
            const int own0 = (int)g.own0, own1 = (int)g.own1;
            const bool lo_tile = peers.lo_p[0] != nullptr && tx0 + h < own0 + peers.H;
            const bool hi_tile = peers.hi_p[0] != nullptr && tx0 + TXR - h > own1 - peers.H;
            if (lo_tile || hi_tile) {
                double *lo = src ? peers.lo_p[0] : peers.lo_p[1];
                double *hi = src ? peers.hi_p[0] : peers.hi_p[1];
                const int lx0 = tx0 + r_begin;
                lo += (peers.lo_row0 + (lx0 - own0)) * g.pitch + gy0;
                hi += (peers.hi_row0 + (lx0 - (own1 - peers.H))) * g.pitch + gy0;
#pragma unroll
                for (int i = 0; i < RPT; i++) {
                    if (!(m_store & (1u << (2 * i)))) continue;
                    const int lx = lx0 + i;
                    if (lo_tile && lx < own0 + peers.H) {
                        double *dst = lo + (int64_t)i * g.pitch;
                        if (pair) *reinterpret_cast<double2 *>(dst) = P[i];
                        else dst[0] = P[i].x;
                    }
                    if (hi_tile && lx >= own1 - peers.H) {
                        double *dst = hi + (int64_t)i * g.pitch;
                        if (pair) *reinterpret_cast<double2 *>(dst) = P[i];
                        else dst[0] = P[i].x;
                    }
                }
            }
        }
    }

    // per-sweep residual sums of the tile (fixed order => deterministic)
    __syncthreads();
    const int levels = norm_only ? 1 : T;
    if ((int)threadIdx.x < levels) {
        double tsum = 0.0;
        for (int w = 0; w < NWARP; w++) tsum += sred[threadIdx.x * NWARP + w];
        partial[(int64_t)threadIdx.x * part_stride + blockIdx.x] = tsum;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

sb_status make_map(EncodeTiledFn fn, CUtensorMap *map, CUtensorMapDataType dt, size_t esize,
                   void *base, const Geom &g) {
    cuuint64_t dims[2] = {(cuuint64_t)g.NY, (cuuint64_t)g.nxl};
    cuuint64_t strides[1] = {(cuuint64_t)g.pitch * esize};
    cuuint32_t box[2] = {TW, TXR};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, dt, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
        return SB_CUDA_ERROR;
    }
    return SB_OK;
}

sb_status ensure_tmaps(sb_sim *s) {
    if (s->tmaps_ready) return SB_OK;
    void *fnp = nullptr;
    cudaDriverEntryPointQueryResult qres;
    SB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &qres));
    if (!fnp || qres != cudaDriverEntryPointSuccess) {
        set_error("cuTensorMapEncodeTiled entry point not available");
        return SB_CUDA_ERROR;
    }
    EncodeTiledFn fn = reinterpret_cast<EncodeTiledFn>(fnp);
    sb_status st;
    for (int i = 0; i < 2; i++)
        if ((st = make_map(fn, &s->tm_p[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, s->p[i], s->g)))
            return st;
    if ((st = make_map(fn, &s->tm_rhs, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, s->rhs, s->g))) return st;
    SB_CUDA(cudaFuncSetAttribute(sor_rb_kernel<0, false>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    SB_CUDA(cudaFuncSetAttribute(sor_rb_kernel<1, false>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    SB_CUDA(cudaFuncSetAttribute(sor_rb_kernel<0, true>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    SB_CUDA(cudaFuncSetAttribute(sor_rb_kernel<1, true>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    s->tmaps_ready = true;
    return SB_OK;
}

}  // namespace

int rb_halo_rows(int T) { return 2 * T + 2; }


// one guarded pass: performs ctl->active_T sweeps from pbuf[ctl->src] into the other buffer
// (norm_only: just the residual partial sums of pbuf[ctl->src]).  The all-fluid part of the
// grid goes to the streaming kernel (sor_rb_stream.cu), everything else to the tile kernel.
sb_status launch_sor_rb_pass(sb_sim *s, int *nparts_out, int norm_only, const RbFin *fin,
                             int *fused) {
    sb_status st = ensure_tmaps(s);
    if (st) return st;
    const Geom &g = s->g;
    int T = s->prm.temporal_block;
    int h = rb_halo_rows(T);
    int BX = TXR - 2 * h, BY = TW - 2 * h;
    int tiles_x = (int)((g.own1 - g.own0 + BX - 1) / BX), tiles_y = (int)((g.NY + BY - 1) / BY);
    int ntiles = tiles_x * tiles_y;
    int n_tile = ntiles, n_items = 0;
    const int32_t *tile_list = nullptr;
    if (!norm_only && s->dbg.rb_stream) {
        if ((st = rb_ensure_plan(s, BX, BY, h))) return st;
        n_tile = s->plan.n_slow;
        n_items = s->plan.n_items;
        tile_list = s->plan.d_slow;
    }
    const int n_frozen = (!norm_only && s->dbg.rb_stream) ? s->plan.n_frozen : 0;
    // partial-sum slots: one per tile, one per warp of a streaming work item, one for the frozen tiles
    const int nparts = n_tile + n_items * rb_stream_slots_per_item(T) + (n_frozen ? 1 : 0);
    size_t need = (size_t)nparts * TMAX + 64;
    if (need > s->partial_cap) {  // stream-ordered: no device-wide synchronisation
        if (s->d_partial) SB_CUDA(cudaFreeAsync(s->d_partial, s->stream));
        s->d_partial = nullptr;
        s->partial_cap = 0;
        SB_CUDA(cudaMallocAsync(&s->d_partial, need * sizeof(double), s->stream));
        s->partial_cap = need;
    }
    RbPeers peers;
    peers.lo_p[0] = s->link.lo_p[0]; peers.lo_p[1] = s->link.lo_p[1];
    peers.hi_p[0] = s->link.hi_p[0]; peers.hi_p[1] = s->link.hi_p[1];
    peers.lo_row0 = s->link.lo_row0; peers.hi_row0 = s->link.hi_row0;
    peers.H = s->link.H;
    const RbConsts k = rb_consts(s);
    // tile row 0 is local row -h (even offset): the colour of the thread's first row follows
    // the parity of the slab's global row offset
    const int par = (int)(((g.gx0 % 2) + 2) % 2);
    if (n_frozen && s->plan.frozen_seq != s->solve_seq) {
        // first pass of a solve: mirror the frozen tiles' pressures into the other buffer and
        // put their residual sum (constant for the whole solve) into the extra partial slot
        if ((st = launch_frozen_mirror(s, BX, BY))) return st;
        auto nk = s->slab ? (par ? sor_rb_kernel<1, true> : sor_rb_kernel<0, true>)
                          : (par ? sor_rb_kernel<1, false> : sor_rb_kernel<0, false>);
        nk<<<n_frozen, NTHR, SMEM_BYTES, s->stream>>>(s->tm_p[0], s->tm_p[1], s->tm_rhs, s->cflag, g,
                                                      rb_pbuf_ptr(s), s->d_ctl,
                                                      s->plan.d_frozen_part, tiles_y, n_frozen, h, k,
                                                      1, peers, s->plan.d_frozen);
        s->launches++;
        if ((st = launch_frozen_fill(s, nparts, nparts - 1))) return st;
        s->plan.frozen_seq = s->solve_seq;
    }
    if (!norm_only) prof_mark(s);
    if (n_tile > 0) {
        auto kern = s->slab ? (par ? sor_rb_kernel<1, true> : sor_rb_kernel<0, true>)
                            : (par ? sor_rb_kernel<1, false> : sor_rb_kernel<0, false>);
        kern<<<n_tile, NTHR, SMEM_BYTES, s->stream>>>(s->tm_p[0], s->tm_p[1], s->tm_rhs, s->cflag,
                                                      g, rb_pbuf_ptr(s), s->d_ctl, s->d_partial,
                                                      tiles_y, nparts, h, k, norm_only, peers,
                                                      tile_list);
        s->launches++;
    }
    if (fused) *fused = 0;
    if (n_items > 0) {
        // the streaming kernel's last CTA finalizes the pass (row slabs: including the
        // all-gather of the per-slab sums); sor_finalize_kernel only runs when there is no
        // streaming item at all
        const RbFin *f = (fin && fused) ? fin : nullptr;
        if ((st = launch_sor_rb_stream(s, n_tile, nparts, h, f))) return st;
        if (f) *fused = 1;
    }
    if (!norm_only) prof_mark(s);
    SB_CUDA(cudaGetLastError());
    *nparts_out = nparts;
    return SB_OK;
}

// force-load this file's kernels (CUDA loads lazily by default, and a first launch that has
// to load code synchronises the context -- fatal while a peer slab of the same process spins
// in an all-gather on the same GPU)
void preload_sor_rb() {
    cudaFuncAttributes a;
    cudaFuncGetAttributes(&a, sor_rb_kernel<0, false>);
    cudaFuncGetAttributes(&a, sor_rb_kernel<1, false>);
    cudaFuncGetAttributes(&a, sor_rb_kernel<0, true>);
    cudaFuncGetAttributes(&a, sor_rb_kernel<1, true>);
    cudaGetLastError();
}

}  // namespace sb
