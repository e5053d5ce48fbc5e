// sor_mid.cu -- K4e: the whole red-black SOR solve of a MID-SIZE grid in ONE cooperative
// launch, with the grid resident in the shared memory of all SMs.
//
// BASELINE config 2 (1024 x 1024 cavity, up to 1000 sweeps per tick) is 8 MB per array: pass by
// pass that is ~500 dependent launches of kernels that run for 20-50 us each, and the data
// never leaves the L2 -- launch latency and the warm-up rows of the streaming items bound it,
// not HBM.  148 SMs x 227 KB hold p, rhs and a code byte of ~1.9 M cells, so here every CTA
// (one per SM, cooperative launch) owns a band of x-rows, keeps it in shared memory for the
// whole of solve_sor (/root/reference/src/simulation.rs:239-285) and talks to its two
// neighbours through a small exchange buffer in L2 with two grid-wide barriers per sweep:
//
//   per iteration        pressure BC on own rows + first halo row    (src/grid/mod.rs:343-412)
//                        red half-sweep on own rows                  (simulation.rs:253-274)
//     exchange 1         2 edge rows each way (BC + red values)      -- grid barrier
//                        black half-sweep on own rows + first halo row (redundant, so that
//                        the residual of the edge rows needs no third exchange)
//                        residual over ALL interior cells of own rows (simulation.rs:216-227)
//     exchange 2         2 edge rows each way + the CTA's partial sum -- grid barrier
//                        every CTA totals the partial sums in the same order and applies the
//                        exit rule (simulation.rs:279): identical decisions, no broadcast
//
// The two halo rows per side are what makes one exchange per half-sweep enough: a boundary
// cell in the first halo row takes its BC from cells of the second.  Arithmetic is that of
// the tile / streaming / small kernels and of the oracle's red-black restatement (sor_rb.cuh):
// p is bit-identical whichever kernel ran; the norm is summed in another order (1e-12
// allowance).
#include <stdlib.h>

#include <algorithm>

#include "sor_rb.cuh"

namespace sb {

namespace {

constexpr int MID_THREADS = 1024;
constexpr int MID_MAX_CTAS = 1024;

struct MidParams {
    int nx, ny;
    int64_t pitch;
    int rows_base, rows_extra;   // CTA c owns rows_base + (c < rows_extra) rows
    int rmax;                    // most rows any CTA owns
    double *const *pbuf;
    const double *rhs;
    const uint8_t *cflag;
    SorCtl *ctl;
    RbConsts k;
    double fluid_cells, initial_norm, eps2;
    int test_exit;
    double *norm_hist;
    double *xbuf;                // [2 exchanges][CTA][4 rows][ny]
    double *partial;             // [CTA]
    unsigned long long *bar;     // arrivals so far (zeroed before the launch)
};

// only the neighbours the edge class names are read (cf. sor_small.cu)
__device__ __forceinline__ double mid_bc(const double *sp, int c, int ny, int edge) {
    switch (edge) {
    case SB_EDGE_N: return sp[c - 1];
    case SB_EDGE_NE: return (sp[c - 1] + sp[c + ny]) / 2.0;
    case SB_EDGE_E: return sp[c + ny];
    case SB_EDGE_SE: return (sp[c + 1] + sp[c + ny]) / 2.0;
    case SB_EDGE_S: return sp[c + 1];
    case SB_EDGE_SW: return (sp[c + 1] + sp[c - ny]) / 2.0;
    case SB_EDGE_W: return sp[c - ny];
    default: return (sp[c - 1] + sp[c - ny]) / 2.0;  // SB_EDGE_NW
    }
}

// All CTAs are co-resident (cooperative launch).  One thread per CTA arrives and polls; the
// fences make the CTA's plain stores visible before the arrival and order the reads after it.
// A wait that lasts ~2 s gives up (never a hung GPU): the solve is then reported as failed.
__device__ __forceinline__ bool grid_barrier(unsigned long long *bar, unsigned long long target,
                                             int *s_fail) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(bar, 1ULL);
        const long long t0 = clock64();
        for (;;) {
            unsigned long long v;
            asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(bar) : "memory");
            if (v >= target) break;
            if (clock64() - t0 > (1LL << 32)) {
                *s_fail = 1;
                break;
            }
        }
        __threadfence();
    }
    __syncthreads();
    return *s_fail == 0;
}

// per-cell code in shared memory: bits 0-3 edge class of a boundary cell that takes the BC,
// bit 4 interior cell (counts in the norm), bit 5 fluid interior cell (swept)
__global__ void __launch_bounds__(MID_THREADS, 1) sor_mid_kernel(const MidParams a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double s_warp[MID_THREADS / 32];
    __shared__ double s_norm;
    __shared__ int s_fail;
    const int NX = a.nx, NY = a.ny, G = (int)gridDim.x, cta = (int)blockIdx.x;
    const int R = a.rows_base + (cta < a.rows_extra ? 1 : 0);
    const int r0 = cta * a.rows_base + min(cta, a.rows_extra);
    const int xl0 = r0 - 2;  // global x of local row 0; local rows: 2 halo + R own + 2 halo
    double *sp = reinterpret_cast<double *>(smem_raw);       // [rmax + 4][NY], local row l
    double *srl = sp + (size_t)(a.rmax + 4) * NY - NY;       // rhs of local rows 1 .. R+2
    uint8_t *sc = reinterpret_cast<uint8_t *>(sp + (size_t)(2 * a.rmax + 6) * NY);
    double *p = a.pbuf[a.ctl->src];
    const uint32_t max_it = a.ctl->max_iterations;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const RbConsts k = a.k;

    // walking a range of rows with all threads: thread t starts at cell t and steps by
    // MID_THREADS cells; (row, column) advance without a division
    const int q0 = tid / NY, y0 = tid - q0 * NY;
    const int dq = MID_THREADS / NY, dy = MID_THREADS - dq * NY;
    // the same over the cells of one colour: H slots per row, column 2j + (row parity)
    const int H = (NY + 1) >> 1;
    const int hq0 = tid / H, j0 = tid - hq0 * H;
    const int hdq = MID_THREADS / H, hdj = MID_THREADS - hdq * H;

    if (tid == 0) s_fail = 0;
    for (int l = q0, y = y0; l < R + 4;) {
        const int x = xl0 + l, ci = l * NY + y;
        double pv = 0.0;
        uint8_t code = 0;
        if (x >= 0 && x < NX) {
            const int64_t gc = (int64_t)x * a.pitch + y;
            pv = p[gc];
            const uint8_t f = a.cflag[gc];
            const bool interior = x >= 1 && x <= NX - 2 && y >= 1 && y <= NY - 2;
            if (cf_is_boundary(f)) code = (uint8_t)cf_edge(f);
            if (interior) code |= 16;
            if (interior && cf_is_fluid(f)) code |= 32;
            if (l >= 1 && l < R + 3) srl[ci] = a.rhs[gc];
        } else if (l >= 1 && l < R + 3) {
            srl[ci] = 0.0;
        }
        sp[ci] = pv;
        sc[ci] = code;
        y += dy; l += dq;
        if (y >= NY) { y -= NY; l++; }
    }
    __syncthreads();

    double *const xb1 = a.xbuf, *const xb2 = a.xbuf + (size_t)G * 4 * NY;
    unsigned long long arrivals = 0;
    uint32_t it = 0;
    double norm = 0.0;
    int cap = 1;
    bool ok = true;

    // my 2 first and 2 last own rows -> exchange buffer xb (slots 0,1 and 2,3)
    auto put_edges = [&](double *xb) {
        double *mine = xb + (size_t)cta * 4 * NY;
        for (int i = tid; i < 4 * NY; i += MID_THREADS) {
            const int slot = i / NY, y = i - slot * NY;
            const int l = slot < 2 ? 2 + slot : R + slot - 2;
            mine[i] = sp[l * NY + y];
        }
    };
    // the neighbours' edge rows -> my halo rows (L2 reads: the lines were written by other SMs)
    auto get_halos = [&](const double *xb) {
        for (int i = tid; i < 4 * NY; i += MID_THREADS) {
            const int slot = i / NY, y = i - slot * NY;
            if (slot < 2) {
                if (cta > 0)
                    sp[slot * NY + y] = __ldcg(xb + ((size_t)(cta - 1) * 4 + 2 + slot) * NY + y);
            } else if (cta < G - 1) {
                sp[(R + slot) * NY + y] = __ldcg(xb + ((size_t)(cta + 1) * 4 + slot - 2) * NY + y);
            }
        }
    };
    // one colour of local rows [la, lb)
    auto half_sweep = [&](int la, int lb, int colour) {
        for (int l = la + hq0, j = j0; l < lb;) {
            const int y = 2 * j + ((xl0 + l + colour) & 1);
            if (y < NY) {
                const int ci = l * NY + y;
                if (sc[ci] & 32) {
                    const double t = fma(k.rdx2, sp[ci + NY] + sp[ci - NY],
                                         fma(k.rdy2, sp[ci + 1] + sp[ci - 1], -srl[ci]));
                    sp[ci] = fma(k.mid, t, k.omw * sp[ci]);
                }
            }
            j += hdj; l += hdq;
            if (j >= H) { j -= H; l++; }
        }
    };

    while (it < max_it) {
        // pressure BC on own rows and the first halo row of each side
        for (int l = 1 + q0, y = y0; l < R + 3;) {
            const int ci = l * NY + y;
            const int edge = sc[ci] & 15;
            double nv = 0.0;
            if (edge) nv = mid_bc(sp, ci, NY, edge);
            // reads of fluid cells, writes of boundary cells: no hazard within the pass
            if (edge) sp[ci] = nv;
            y += dy; l += dq;
            if (y >= NY) { y -= NY; l++; }
        }
        __syncthreads();
        half_sweep(2, R + 2, 0);
        __syncthreads();
        put_edges(xb1);
        arrivals += G;
        if (!(ok = grid_barrier(a.bar, arrivals, &s_fail))) break;
        get_halos(xb1);
        __syncthreads();
        half_sweep(1, R + 3, 1);
        __syncthreads();
        // residual norm over ALL interior cells of own rows
        double acc = 0.0;
        for (int l = 2 + q0, y = y0; l < R + 2;) {
            const int ci = l * NY + y;
            if (sc[ci] & 16) {
                const double t = fma(k.rdx2, sp[ci + NY] + sp[ci - NY],
                                     fma(k.rdy2, sp[ci + 1] + sp[ci - 1], -srl[ci]));
                const double r = fma(-k.diag, sp[ci], t);
                acc = fma(r, r, acc);
            }
            y += dy; l += dq;
            if (y >= NY) { y -= NY; l++; }
        }
        acc = warp_sum_down(acc);
        if (lane == 0) s_warp[warp] = acc;
        put_edges(xb2);
        __syncthreads();
        if (tid == 0) {
            double t = 0.0;
            for (int w = 0; w < MID_THREADS / 32; w++) t += s_warp[w];
            a.partial[cta] = t;
        }
        arrivals += G;
        if (!(ok = grid_barrier(a.bar, arrivals, &s_fail))) break;
        get_halos(xb2);
        if (warp == 0) {
            double t = 0.0;
            for (int i = lane; i < G; i += 32) t += __ldcg(a.partial + i);
            for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            if (lane == 0) s_norm = t / a.fluid_cells;
        }
        __syncthreads();
        norm = s_norm;
        if (cta == 0 && tid == 0 && a.norm_hist) a.norm_hist[it] = norm;
        it++;
        if (a.test_exit && ((norm < a.initial_norm) || (norm < a.eps2))) { cap = 0; break; }
    }

    for (int l = 2 + q0, y = y0; l < R + 2;) {
        p[(int64_t)(xl0 + l) * a.pitch + y] = sp[l * NY + y];
        y += dy; l += dq;
        if (y >= NY) { y -= NY; l++; }
    }
    if (tid == 0 && (cta == 0 || !ok)) {
        if (!ok) a.ctl->pad = 1;  // a grid barrier timed out
        if (cta == 0) {
            a.ctl->iters_done = it;
            a.ctl->last_norm = norm;
            a.ctl->norms[0] = norm;
            a.ctl->active_T = 0;
            a.ctl->finished = 1;
            a.ctl->cap_hit = (cap && max_it > 0 && ok) ? 1 : 0;
        }
    }
}

// ---- register-window variant (NY <= 1024, 4..8 rows per CTA) -----------------------------
// The generic kernel above spends its time on index arithmetic and shared-memory traffic
// (~40 instructions per cell visit).  Here thread j owns the column pair (2j, 2j+1) of its
// CTA's band for the whole solve and keeps those pressures in REGISTERS: the x-neighbours of
// a cell are registers of the same thread, one y-neighbour is the thread's other column, the
// second comes from the neighbouring thread through a shared-memory mirror (split into even
// and odd columns: consecutive threads touch consecutive words).  Local rows are anchored at
// an even global x, so the colour of (row l, column c) is the compile-time (l + c) & 1 of the
// unrolled window and every cell's role -- swept, counted, boundary, halo -- is a bit of a
// per-thread mask.  FOUR halo rows per side let the whole sweep run between two exchanges
// (BC on 3 halo rows, red on 2, black on 1, residual on own rows): ONE grid barrier per sweep.
constexpr int REG_THREADS = 512;
constexpr int REG_CPP = REG_THREADS + 2;   // mirror: column pairs + one pad slot per end
constexpr int REG_ROWP = 2 * REG_CPP;      // mirror row: [pad E(0..511) pad][pad O(0..511) pad]
constexpr int REG_RROW = 2 * REG_THREADS;  // rhs row: [E(0..511)][O(0..511)]
// strides are compile-time constants: inside the unrolled window every shared-memory access
// is one base register (this thread's column pair) plus an immediate offset

template <int RM>
__global__ void __launch_bounds__(REG_THREADS, 1) sor_mid_reg_kernel(const MidParams a) {
    constexpr int LR = RM + 9;   // local rows: 4 + RM + 4, +1 for the even anchor
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double s_warp[REG_THREADS / 32];
    __shared__ double s_norm;
    __shared__ int s_fail;
    __shared__ long long s_t[8];          // phase timers of thread 0 (SB_MID_TRACE)
    const int NX = a.nx, NY = a.ny, G = (int)gridDim.x, cta = (int)blockIdx.x;
    const int R = a.rows_base + (cta < a.rows_extra ? 1 : 0);
    const int r0 = cta * a.rows_base + min(cta, a.rows_extra);
    const int lo = (r0 - 4) & 1;          // first valid local row (global x = r0 - 4)
    const int hi = lo + R + 8;            // one past the last valid local row
    const int x0 = r0 - 4 - lo;           // global x of local row 0: even
    const int CP = (NY + 1) >> 1;         // column pairs in use
    double *sp = reinterpret_cast<double *>(smem_raw);     // mirror rows lo .. hi-1 (a.rmax + 8)
    double *sr = sp + (size_t)(a.rmax + 8) * REG_ROWP;     // rhs rows lo+2 .. hi-3 (a.rmax + 4)
    const uint32_t max_it = a.ctl->max_iterations;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int j = tid;
    const bool active = j < CP;
    const RbConsts k = a.k;
    // this thread's slots, indexed by local row: mirror E at pj[l*ROWP + 1], O at
    // pj[l*ROWP + CPP + 1]; rhs E at rj[l*RROW], O at rj[l*RROW + 512]
    double *const pj = sp - (ptrdiff_t)lo * REG_ROWP + j;
    double *const rj = sr - (ptrdiff_t)(lo + 2) * REG_RROW + j;

    // Register window: rows 1 .. LR-2.  The outermost halo rows only feed the BC (mirror).
    double P[LR][2];
    // bit l of: u_red / u_blk = the red / black cell of row l is swept in that phase;
    // r_red / r_blk = it counts in the norm; bm[c] = cell (l, c) takes the pressure BC;
    // halo = row l is a halo row with a neighbour behind it; own = row l is mine
    uint32_t u_red = 0, u_blk = 0, r_red = 0, r_blk = 0, bm[2] = {0, 0}, halo = 0, own = 0;
    unsigned long long e4[2] = {0, 0};    // edge class of BC cells, nibble l - lo - 1
    // BC cells whose one fluid neighbour is this thread's other column (the walls y = 0 and
    // y = NY-1 of a channel): a register move, no mirror read
    uint32_t bs0 = 0, bn1 = 0;

    if (tid == 0) s_fail = 0;
    if (tid < 8) s_t[tid] = 0;
    for (int i = tid; i < (a.rmax + 8) * REG_ROWP; i += REG_THREADS) sp[i] = 0.0;
    __syncthreads();
    double *const pg = a.pbuf[a.ctl->src];
#pragma unroll
    for (int l = 0; l < LR; l++) {
        const int x = x0 + l;
        const bool rowok = l >= lo && l < hi && x >= 0 && x < NX;
        if (l >= lo + 4 && l < hi - 4) own |= 1u << l;
        if (rowok && ((l < lo + 4 && cta > 0) || (l >= hi - 4 && cta < G - 1))) halo |= 1u << l;
#pragma unroll
        for (int c = 0; c < 2; c++) {
            const int y = 2 * j + c;
            double pv = 0.0;
            if (rowok && active && y < NY) {
                const int64_t gc = (int64_t)x * a.pitch + y;
                pv = pg[gc];
                const uint8_t f = a.cflag[gc];
                const bool interior = x >= 1 && x <= NX - 2 && y >= 1 && y <= NY - 2;
                const bool fluid = interior && cf_is_fluid(f);
                const int edge = cf_is_boundary(f) ? cf_edge(f) : 0;
                const bool red = ((l + c) & 1) == 0;
                if (fluid && red && l >= lo + 2 && l < hi - 2) u_red |= 1u << l;
                if (fluid && !red && l >= lo + 3 && l < hi - 3) u_blk |= 1u << l;
                if (interior && l >= lo + 4 && l < hi - 4) {
                    if (red) r_red |= 1u << l;
                    else r_blk |= 1u << l;
                }
                if (edge && l >= lo + 1 && l < hi - 1) {
                    if (c == 0 && edge == SB_EDGE_S) bs0 |= 1u << l;
                    else if (c == 1 && edge == SB_EDGE_N) bn1 |= 1u << l;
                    else {
                        bm[c] |= 1u << l;
                        e4[c] |= (unsigned long long)edge << (4 * (l - lo - 1));
                    }
                }
                pj[l * REG_ROWP + c * REG_CPP + 1] = pv;
                if (l >= lo + 2 && l < hi - 2) rj[l * REG_RROW + c * REG_THREADS] = a.rhs[gc];
            }
            if (l >= 1 && l <= LR - 2) P[l][c] = pv;
        }
    }
    __syncthreads();

    const size_t xrow = (size_t)2 * CP;            // doubles per exchanged row
    const size_t xhalf = (size_t)G * 8 * xrow;     // one parity of the exchange buffer
    unsigned long long arrivals = 0;
    uint32_t it = 0;
    double norm = 0.0;
    int cap = 1;
    bool ok = true;
    const uint32_t any_bc = bm[0] | bm[1];
    long long t_last = clock64();
#define MID_PHASE(n)                                  \
    if (tid == 0) {                                   \
        const long long t_now = clock64();            \
        s_t[n] += t_now - t_last;                     \
        t_last = t_now;                               \
    }
    // t = rdx2 (pE + pW) + rdy2 (pS + pN) - rhs for cell (l, c): x-neighbours in registers, one
    // y-neighbour is my other column, the other one the neighbouring thread's (mirror)
#define MID_T(l, c)                                                                         \
    fma(k.rdx2, P[(l) + 1][c] + P[(l) - 1][c],                                              \
        fma(k.rdy2,                                                                         \
            P[l][1 - (c)] + ((c) == 0 ? pj[(l) * REG_ROWP + REG_CPP] : pj[(l) * REG_ROWP + 2]), \
            -rj[(l) * REG_RROW + (c) * REG_THREADS]))

    while (it < max_it) {
        // pressure BC (src/grid/mod.rs:343-412): fluid cells -> boundary cells
        if (bs0 | bn1) {
#pragma unroll
            for (int l = 1; l <= LR - 2; l++) {
                if ((bs0 >> l) & 1) {
                    P[l][0] = P[l][1];
                    pj[l * REG_ROWP + 1] = P[l][1];
                }
                if ((bn1 >> l) & 1) {
                    P[l][1] = P[l][0];
                    pj[l * REG_ROWP + REG_CPP + 1] = P[l][0];
                }
            }
        }
        if (any_bc) {  // the general case, in the mirror
#pragma unroll
            for (int c = 0; c < 2; c++) {
                uint32_t m = bm[c];
                while (m) {
                    const int l = __ffs(m) - 1;
                    m &= m - 1;
                    const int edge = (int)((e4[c] >> (4 * (l - lo - 1))) & 15);
                    double *me = pj + l * REG_ROWP + c * REG_CPP + 1;
                    // y-1 / y+1 of column c: the other column of this pair or of the next pair
                    const double *yn = c == 0 ? me + REG_CPP - 1 : me - REG_CPP;
                    const double *ys = c == 0 ? me + REG_CPP : me - REG_CPP + 1;
                    // branch-free: one neighbour along y, one along x, both loads in flight
                    // (edge classes: N 1, NE 2, E 3, SE 4, S 5, SW 6, W 7, NW 8)
                    const bool has_n = (0x106u >> edge) & 1, has_s = (0x070u >> edge) & 1;
                    const bool has_e = (0x01cu >> edge) & 1, has_w = (0x1c0u >> edge) & 1;
                    const double yv = *(has_n ? yn : ys);
                    const double xv = *(has_e ? me + REG_ROWP : me - REG_ROWP);
                    const bool has_y = has_n || has_s, has_x = has_e || has_w;
                    const double v = (has_y && has_x) ? (yv + xv) / 2.0 : (has_y ? yv : xv);
                    *me = v;
                }
            }
        }
        __syncthreads();
        MID_PHASE(0)
        if (any_bc) {
#pragma unroll
            for (int l = 1; l <= LR - 2; l++) {
                if ((bm[0] >> l) & 1) P[l][0] = pj[l * REG_ROWP + 1];
                if ((bm[1] >> l) & 1) P[l][1] = pj[l * REG_ROWP + REG_CPP + 1];
            }
        }
        // red half-sweep: own rows and two halo rows per side
#pragma unroll
        for (int l = 2; l <= LR - 3; l++) {
            constexpr int dummy = 0;
            (void)dummy;
            const int c = l & 1;
            if ((u_red >> l) & 1) {
                const double t = MID_T(l, c);
                const double pn = fma(k.mid, t, k.omw * P[l][c]);
                P[l][c] = pn;
                pj[l * REG_ROWP + c * REG_CPP + 1] = pn;
            }
        }
        __syncthreads();
        MID_PHASE(1)
        // black half-sweep: own rows and one halo row per side; the residual of a black cell
        // sees its final neighbours already (math.rs:176-186 on the swept field)
        double acc = 0.0;
#pragma unroll
        for (int l = 2; l <= LR - 3; l++) {
            const int c = 1 - (l & 1);
            const bool upd = (u_blk >> l) & 1, res = (r_blk >> l) & 1;
            if (upd || res) {
                const double t = MID_T(l, c);
                double pc = P[l][c];
                if (upd) {
                    pc = fma(k.mid, t, k.omw * pc);
                    P[l][c] = pc;
                    pj[l * REG_ROWP + c * REG_CPP + 1] = pc;
                }
                if (res) {
                    const double r = fma(-k.diag, pc, t);
                    acc = fma(r, r, acc);
                }
            }
        }
        __syncthreads();
        MID_PHASE(2)
        // my four first and four last own rows -> the exchange buffer of this sweep's parity
        double *xb = a.xbuf + (it & 1) * xhalf;
        if (active) {
            double *mine = xb + (size_t)cta * 8 * xrow + 2 * j;
            const double *first = pj + (lo + 4) * REG_ROWP + 1, *last = pj + (lo + R) * REG_ROWP + 1;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                *reinterpret_cast<double2 *>(mine + i * xrow) =
                    make_double2(first[i * REG_ROWP], first[i * REG_ROWP + REG_CPP]);
                *reinterpret_cast<double2 *>(mine + (4 + i) * xrow) =
                    make_double2(last[i * REG_ROWP], last[i * REG_ROWP + REG_CPP]);
            }
        }
        // residual of the red cells of own rows, now that their black neighbours are final
#pragma unroll
        for (int l = 2; l <= LR - 3; l++) {
            const int c = l & 1;
            if ((r_red >> l) & 1) {
                const double t = MID_T(l, c);
                const double r = fma(-k.diag, P[l][c], t);
                acc = fma(r, r, acc);
            }
        }
        acc = warp_sum_down(acc);
        if (lane == 0) s_warp[warp] = acc;
        __syncthreads();   // edge rows stored by all threads, warp sums in place
        MID_PHASE(3)
        arrivals += G;
        if (warp == 0) {
            // the CTA's sum (fixed order), then arrive: the release orders the CTA's edge rows
            // and the sum before the arrival; one thread polls, its acquire orders the reads
            double t = lane < REG_THREADS / 32 ? s_warp[lane] : 0.0;
            for (int o = 8; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            if (lane == 0) {
                a.partial[(it & 1) * G + cta] = t;
                asm volatile("red.release.gpu.global.add.u64 [%0], 1;" ::"l"(a.bar) : "memory");
                const long long t0 = clock64();
                for (;;) {
                    unsigned long long v;
                    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(a.bar) : "memory");
                    if (v >= arrivals) break;
                    if (clock64() - t0 > (1LL << 32)) { s_fail = 1; break; }
                }
            }
        }
        __syncthreads();
        if (s_fail) { ok = false; break; }
        MID_PHASE(4)
        // the CTAs' partial sums (warp 0; loads first, they fly with the halo loads) ...
        double pv[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
        if (warp == 0) {
            const double *part = a.partial + (it & 1) * G;
#pragma unroll
            for (int q = 0; q < 5; q++)
                if (32 * q + lane < G) pv[q] = __ldcg(part + 32 * q + lane);
        }
        // ... and the neighbours' edge rows -> my halo rows (mirror first, registers after the
        // barrier)
        if (active) {
            double *lower = pj + lo * REG_ROWP + 1, *upper = pj + (hi - 4) * REG_ROWP + 1;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                if (cta > 0) {
                    const double2 v = __ldcg(reinterpret_cast<const double2 *>(
                        xb + ((size_t)(cta - 1) * 8 + 4 + i) * xrow + 2 * j));
                    lower[i * REG_ROWP] = v.x;
                    lower[i * REG_ROWP + REG_CPP] = v.y;
                }
            }
#pragma unroll
            for (int i = 0; i < 4; i++) {
                if (cta < G - 1) {
                    const double2 v = __ldcg(reinterpret_cast<const double2 *>(
                        xb + ((size_t)(cta + 1) * 8 + i) * xrow + 2 * j));
                    upper[i * REG_ROWP] = v.x;
                    upper[i * REG_ROWP + REG_CPP] = v.y;
                }
            }
        }
        if (warp == 0) {
            // same order on every CTA: identical norms, identical decisions
            double t = 0.0;
#pragma unroll
            for (int q = 0; q < 5; q++) t += pv[q];
            for (int base = 160; base < G; base += 32)   // more than 160 CTAs: the rest
                if (base + lane < G) t += __ldcg(a.partial + (it & 1) * G + base + lane);
            for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            if (lane == 0) s_norm = t / a.fluid_cells;
        }
        __syncthreads();
        if (active) {
#pragma unroll
            for (int l = 1; l <= LR - 2; l++) {
                if ((halo >> l) & 1) {
                    P[l][0] = pj[l * REG_ROWP + 1];
                    P[l][1] = pj[l * REG_ROWP + REG_CPP + 1];
                }
            }
        }
        norm = s_norm;
        if (cta == 0 && tid == 0 && a.norm_hist) a.norm_hist[it] = norm;
        it++;
        MID_PHASE(5)
        if (a.test_exit && ((norm < a.initial_norm) || (norm < a.eps2))) { cap = 0; break; }
    }
#undef MID_T
#undef MID_PHASE

    if (active) {
#pragma unroll
        for (int l = 1; l <= LR - 2; l++) {
            if ((own >> l) & 1) {
                double *dst = pg + (int64_t)(x0 + l) * a.pitch + 2 * j;
                dst[0] = P[l][0];
                if (2 * j + 1 < NY) dst[1] = P[l][1];
            }
        }
    }
    if (tid == 0 && (cta == 0 || !ok)) {
        if (!ok) a.ctl->pad = 1;  // a grid barrier timed out
        if (cta == 0) {
            a.ctl->iters_done = it;
            a.ctl->last_norm = norm;
            a.ctl->norms[0] = norm;
            a.ctl->active_T = 0;
            a.ctl->finished = 1;
            a.ctl->cap_hit = (cap && max_it > 0 && ok) ? 1 : 0;
            // diagnostics (SB_MID_TRACE): cycles of CTA 0 per phase -- BC, red, black,
            // red residual + publish, grid barrier, halo + norm
            for (int i = 0; i < 6; i++) a.ctl->norms[1 + i] = (double)s_t[i];
        }
    }
}

struct MidPlan {
    int variant;   // 1 generic (shared memory only), 2 register window
    int ctas, rows_base, rows_extra, rmax;
    size_t smem, xdoubles;
};

struct DevInfo {
    int sms = 0, smem_optin = 0, coop = 0;
    bool ready = false, attr_set = false;
};
DevInfo g_dev[64];

void split_rows(MidPlan *pl, int64_t nx, int ctas) {
    pl->ctas = ctas;
    pl->rows_base = (int)(nx / ctas);
    pl->rows_extra = (int)(nx % ctas);
    pl->rmax = pl->rows_base + (pl->rows_extra ? 1 : 0);
}

bool mid_plan(const sb_sim *s, MidPlan *pl) {
    DevInfo &d = g_dev[s->device & 63];
    if (!d.ready) {
        cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, s->device);
        cudaDeviceGetAttribute(&d.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, s->device);
        cudaDeviceGetAttribute(&d.coop, cudaDevAttrCooperativeLaunch, s->device);
        d.ready = true;
    }
    if (!d.coop || d.sms < 1) return false;
    const int64_t nx = s->g.NX, ny = s->g.NY;
    if (nx < 3 || ny < 3 || nx > (1 << 20) || ny > (1 << 20)) return false;
    const size_t budget = (size_t)d.smem_optin - 1024;
    int want = std::min(d.sms, MID_MAX_CTAS), variant = 0;
    if (s->dbg.sor_mid_ctas >= 1) want = std::min(want, s->dbg.sor_mid_ctas);   // tests: force a
    variant = s->dbg.sor_mid_variant;                                        // decomposition / kernel
    // register window: column pairs on 512 threads, 4..8 rows per CTA (4 halo rows per side
    // must be own rows of the direct neighbour); more CTAs than asked for if bands get too tall
    if (variant != 1 && ny <= 2 * REG_THREADS && nx >= 4) {
        const int cp = (int)((ny + 1) / 2);
        for (int ctas = (int)std::min<int64_t>(want, nx / 4); ctas <= std::min<int64_t>(d.sms, nx / 4);
             ctas++) {
            split_rows(pl, nx, ctas);
            if (pl->rmax > 8) continue;
            pl->smem = ((size_t)(pl->rmax + 8) * REG_ROWP + (size_t)(pl->rmax + 4) * REG_RROW) * 8;
            if (pl->smem > budget) continue;
            pl->variant = 2;
            pl->xdoubles = (size_t)2 * ctas * 8 * 2 * cp;
            return true;
        }
    }
    if (variant == 2) return false;
    // generic: p rmax+4 rows, rhs rmax+2 rows, code rmax+4 rows of bytes
    for (int ctas = (int)std::min<int64_t>(want, std::max<int64_t>(nx / 2, 1));
         ctas <= std::min<int64_t>(d.sms, std::max<int64_t>(nx / 2, 1)); ctas++) {
        split_rows(pl, nx, ctas);
        pl->smem = (size_t)(2 * pl->rmax + 6) * ny * 8 + (size_t)(pl->rmax + 4) * ny + 16;
        if (pl->smem > budget) continue;
        pl->variant = 1;
        pl->xdoubles = (size_t)2 * ctas * 4 * ny;
        return true;
    }
    return false;
}

template <int RM>
sb_status launch_reg(const MidPlan &pl, MidParams &a, int smem_optin, cudaStream_t stream) {
    static bool attr_set[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_set[dev & 63]) {
        SB_CUDA(cudaFuncSetAttribute(sor_mid_reg_kernel<RM>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin - 1024));
        attr_set[dev & 63] = true;
    }
    void *args[] = {&a};
    const cudaError_t e = cudaLaunchCooperativeKernel((const void *)sor_mid_reg_kernel<RM>,
                                                      dim3(pl.ctas), dim3(REG_THREADS), args,
                                                      pl.smem, stream);
    if (e == cudaErrorCooperativeLaunchTooLarge) {
        cudaGetLastError();
        return SB_INVALID_ARGUMENT;   // "not now": the caller falls back to the pass kernels
    }
    SB_CUDA(e);
    return SB_OK;
}

}  // namespace

// grids this path takes: single GPU, not small enough for one SM, the whole of p, rhs and the
// cell codes in the shared memory of the SMs together
bool sor_mid_fits(const sb_sim *s) {
    if (!s->dbg.sor_mid) return false;   // tests: keep such grids on the pass kernels
    if (s->slab) return false;
    MidPlan pl;
    return mid_plan(s, &pl);
}

// the whole solve; the host has initialised *d_ctl (src, max_iterations) and reads it back
sb_status launch_sor_mid(sb_sim *s, double initial_norm, double eps2, int test_exit,
                         double *norm_hist) {
    MidPlan pl;
    if (!mid_plan(s, &pl)) {
        set_error("launch_sor_mid: grid does not fit (internal error)");
        return SB_CUDA_ERROR;
    }
    DevInfo &d = g_dev[s->device & 63];
    if (!d.attr_set) {
        SB_CUDA(cudaFuncSetAttribute(sor_mid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     d.smem_optin - 1024));
        d.attr_set = true;
    }
    // barrier word, exchange rows, partial sums (two parities)
    const size_t need = (pl.xdoubles + 2 * pl.ctas) * sizeof(double) + 128;
    if (s->mid_cap < need) {
        if (s->d_mid) SB_CUDA(cudaFreeAsync(s->d_mid, s->stream));
        s->d_mid = nullptr;
        s->mid_cap = 0;
        SB_CUDA(cudaMallocAsync(&s->d_mid, need, s->stream));
        s->mid_cap = need;
    }
    MidParams a{};
    a.nx = (int)s->g.NX;
    a.ny = (int)s->g.NY;
    a.pitch = s->g.pitch;
    a.rows_base = pl.rows_base;
    a.rows_extra = pl.rows_extra;
    a.rmax = pl.rmax;
    a.pbuf = rb_pbuf_ptr(s);
    a.rhs = s->rhs;
    a.cflag = s->cflag;
    a.ctl = s->d_ctl;
    a.k = rb_consts(s);
    a.fluid_cells = s->fluid_cells;
    a.initial_norm = initial_norm;
    a.eps2 = eps2;
    a.test_exit = test_exit;
    a.norm_hist = norm_hist;
    a.bar = reinterpret_cast<unsigned long long *>(s->d_mid);
    a.xbuf = reinterpret_cast<double *>(reinterpret_cast<char *>(s->d_mid) + 128);
    a.partial = a.xbuf + pl.xdoubles;
    SB_CUDA(cudaMemsetAsync(a.bar, 0, 128, s->stream));
    void *args[] = {&a};
    prof_mark(s);
    sb_status st = SB_OK;
    if (pl.variant == 2) {
        if (pl.rmax <= 5) st = launch_reg<5>(pl, a, d.smem_optin, s->stream);
        else if (pl.rmax == 6) st = launch_reg<6>(pl, a, d.smem_optin, s->stream);
        else if (pl.rmax == 7) st = launch_reg<7>(pl, a, d.smem_optin, s->stream);
        else st = launch_reg<8>(pl, a, d.smem_optin, s->stream);
    } else {
        const cudaError_t e = cudaLaunchCooperativeKernel((const void *)sor_mid_kernel,
                                                          dim3(pl.ctas), dim3(MID_THREADS), args,
                                                          pl.smem, s->stream);
        if (e == cudaErrorCooperativeLaunchTooLarge) {
            cudaGetLastError();
            st = SB_INVALID_ARGUMENT;
        } else {
            SB_CUDA(e);
        }
    }
    if (st == SB_INVALID_ARGUMENT) {  // not all CTAs can be resident here: never try again
        s->mid_unavailable = true;
        prof_mark(s);   // closes the pair opened above (an empty entry)
        return SB_OK;
    }
    if (st) return st;
    s->launches++;
    s->last_sor_ctas = pl.ctas;
    s->last_sor_path = pl.variant == 2 ? 3 : 2;
    prof_mark(s);
    SB_CUDA(cudaGetLastError());
    return SB_OK;
}

void preload_sor_mid() {
    cudaFuncAttributes a;
    cudaFuncGetAttributes(&a, sor_mid_kernel);
    cudaFuncGetAttributes(&a, sor_mid_reg_kernel<5>);
    cudaFuncGetAttributes(&a, sor_mid_reg_kernel<6>);
    cudaFuncGetAttributes(&a, sor_mid_reg_kernel<7>);
    cudaFuncGetAttributes(&a, sor_mid_reg_kernel<8>);
    cudaGetLastError();
}

}  // namespace sb
