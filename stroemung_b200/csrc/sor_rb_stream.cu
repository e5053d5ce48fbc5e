// sor_rb_stream.cu -- K4c: the performance-mode SOR pass on all-fluid regions, as a
// register-resident row pipeline ("2.5-D" temporal blocking).
//
// Same arithmetic as the tile kernel (sor_rb.cuh; the red-black restatement of the sweep of
// /root/reference/src/simulation.rs:253-274 and of calculate_norm_squared, :216-227), other
// schedule.  Where the footprint of a piece of the grid holds nothing but interior fluid
// cells with fluid neighbours, no pressure BC and no cell test is needed, and T sweeps can
// be pipelined along x by ONE WARP without any block-level synchronisation:
//
//   * a warp owns a 128-column strip (4 adjacent columns per lane) and walks down x;
//   * in the tick in which row R arrives it runs, for k = 0..T-1, the red half-sweep of
//     sweep k on row R-(2k+1) and the black half-sweep on row R-(2k+2) (each needs its two
//     neighbour rows one half-sweep behind -- true in this order), then the residual of the
//     red cells of the last sweep on row R-(2T+1), and retires row R-(2T+2) to HBM;
//   * the 2T+4 rows in flight live in REGISTERS (the tick loop is unrolled over the window so
//     every row has a fixed register name); the only exchange between lanes is one shuffle
//     per half-sweep (the y-neighbour across the lane boundary);
//   * rows of p and rhs are fetched PF rows ahead by 1-D bulk copies of the TMA engine into
//     per-warp shared-memory rings (mbarrier per slot); rhs stays in its ring for the 2T+2
//     ticks a row is worked on;
//   * residuals ride along exactly as in the tile kernel: black cells at their update, red
//     cells of sweep k inside the red half-sweep of sweep k+1.
//
// Halo: the strip carries h = 2T+2 columns per side like a tile, but along x only the two
// ends of a work item pay 2T+2 warm-up rows -- a tile pays them every 48 rows.  Useful work
// per cell update rises from 58 % (T = 3 tile) to ~80 %, and there is no load phase: loads,
// arithmetic and stores of different rows overlap all the time.
//
// HBM traffic per item and cell: 8 (p) + 8 (rhs) + 8 (p out) = 24 B for T sweeps (+ the
// strip halo re-reads, L2 hits when neighbouring strips run side by side).
//
// Which tiles are "plain" is decided per lattice tile by rb_plain_kernel; the host turns
// runs of plain tiles along x into work items (RbPlan).
#include <algorithm>

#include "sor_rb.cuh"

namespace sb {

namespace {

constexpr int SW = RB_TW;      // strip columns: 32 lanes x 2 pairs of columns
constexpr int ROW_BYTES = SW * 8;

// Everything that indexes a ring is a compile-time constant inside the unrolled window:
//   * the tick loop is unrolled over NW = 2T+4 ticks (the register window), U = tick mod NW;
//   * row R lands in slot U of the rhs ring (NW slots, one mbarrier each) and in slot
//     U mod (T+2) of the p ring (NW/2 slots: a p row is consumed in its arrival tick);
//   * the rhs of row q is read at ticks q+1, q+3, .. q+2T-1 (the red half-sweeps; the black
//     cells' values are carried one tick in registers) -- the red values read at q+2T-1 are
//     carried two more ticks for the residual of the last sweep, so a slot is free 2T ticks
//     after its row arrived and rows can be requested PF <= 4 ticks ahead.
__host__ __device__ constexpr int stream_nw(int T) { return 2 * T + 4; }
__host__ __device__ constexpr int stream_np(int T) { return T + 2; }
// rows requested ahead: <= T+1 (p ring) and <= 4 (rhs ring); measured best at 4 (profiles/)
#ifndef SB_STREAM_PF
#define SB_STREAM_PF 4
#endif
__host__ __device__ constexpr int stream_pf(int T) {
    return SB_STREAM_PF < T + 1 ? SB_STREAM_PF : T + 1;
}
__host__ __device__ constexpr int stream_smem(int TB) {
    return (stream_np(TB) + stream_nw(TB)) * ROW_BYTES + stream_nw(TB) * 8;
}
// warps per SM the register budget of the TB instantiation is cut for.  The register file
// is split per SM sub-partition (16 K registers each), so the steps are 16 warps (128
// registers per thread), 12 (168) and 8 (255): TB = 4 keeps 48 pressures per lane in flight
// and needs the last one.
__host__ __device__ constexpr int stream_min_ctas(int TB) {
    return TB == 1 ? 16 : TB <= 3 ? 12 : 8;
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void stg_f64x2(double *p, double a, double b) {
    asm volatile("st.global.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(a), "d"(b) : "memory");
}

// Lane l owns column pairs A = (2l, 2l+1) and B = (64+2l, 64+2l+1) of the strip: 16-byte
// shared-memory reads at a 16-byte lane stride are bank-conflict free and the stores of a
// row are two fully coalesced 512-byte segments.  W[row][0..3] = {A0, A1, B0, B1}.
struct SCtx {
    const double *pin, *rhs;   // first column of the strip in local row 0
    double *pout;
    int64_t pitch;
    int x0, x1;                // rows counted and stored
    int first, re;             // rows loaded: [first, re)
    int lane, lane_m1, lane_p1;
    bool cmA, cmB;             // the pair lies in the strip's inner columns [h, SW-h)
    double *pring, *rring;     // ring bases
    const double *pl, *rl;     // this lane's pair A in slot 0 of the p / rhs ring
    uint64_t *bar;
    RbConsts k;
};

// request row `row` (both arrays) into rhs slot `slot` / p slot `slot % NP`; lane 0 only
template <int NP>
__device__ __forceinline__ void issue_row(const SCtx &c, int64_t off, int slot) {
    uint64_t *bar = c.bar + slot;
    mbar_expect_tx(bar, 2 * ROW_BYTES);
    bulk_load(c.pring + (slot % NP) * SW, c.pin + off, ROW_BYTES, bar);
    bulk_load(c.rring + slot * SW, c.rhs + off, ROW_BYTES, bar);
}

// stencil sums t of the two cells of one colour in a lane's two column pairs.
// SET 0: the first cells A0, B0 (y-neighbours: the lane's own second cell and the second cell
// of the pair to the left), SET 1: the second cells A1, B1 (own first cell, first cell of the
// pair to the right).  Pair B of lane 0 continues pair A of lane 31 and vice versa.
template <int SET>
__device__ __forceinline__ void stencil2(const SCtx &c, const double (&me)[4],
                                         const double (&up)[4], const double (&dn)[4],
                                         double rha, double rhb, double &ta, double &tb) {
    const RbConsts &k = c.k;
    if (SET == 0) {
        const double la = __shfl_sync(0xffffffffu, me[1], c.lane_m1);
        const double lb0 = __shfl_sync(0xffffffffu, me[3], c.lane_m1);
        const double lb = c.lane == 0 ? la : lb0;
        ta = fma(k.rdx2, dn[0] + up[0], fma(k.rdy2, me[1] + la, -rha));
        tb = fma(k.rdx2, dn[2] + up[2], fma(k.rdy2, me[3] + lb, -rhb));
    } else {
        const double ra0 = __shfl_sync(0xffffffffu, me[0], c.lane_p1);
        const double rb = __shfl_sync(0xffffffffu, me[2], c.lane_p1);
        const double ra = c.lane == 31 ? rb : ra0;
        ta = fma(k.rdx2, dn[1] + up[1], fma(k.rdy2, ra + me[0], -rha));
        tb = fma(k.rdx2, dn[3] + up[3], fma(k.rdy2, rb + me[2], -rhb));
    }
}

// One tick: row R has been requested PF ticks ago.  U = (R - rs) mod NW is the register slot
// of row R and its ring slot; after inlining into the unrolled loop every index below is a
// constant.  roff = R * pitch.  ph = parity of the mbarrier phase of this loop iteration.
// STEADY: every row this tick touches is a counted row (no row tests at all).
template <int T, bool STEADY>
__device__ __forceinline__ void stream_tick(double (&W)[2 * T + 4][4], double (&C)[T][2],
                                            double (&FR)[2][2], double (&accA)[T],
                                            double (&accB)[T], const int U, const int R,
                                            const int64_t roff, const uint32_t ph,
                                            const SCtx &c) {
    constexpr int NW = stream_nw(T), NP = stream_np(T), PF = stream_pf(T);
    const RbConsts &k = c.k;
    const unsigned nrows = (unsigned)(c.x1 - c.x0);
    // ---- request row R + PF; row R: shared-memory ring -> registers ------------------------
    __syncwarp();  // every lane is done with the slots the request overwrites
    if (c.lane == 0) {
        if (STEADY || R + PF < c.re) issue_row<NP>(c, roff + PF * c.pitch, (U + PF) % NW);
        if (!STEADY && R < c.first) mbar_arrive(c.bar + U);  // keeps the phases in step
    }
    if (STEADY || R >= c.first) {
        mbar_wait(c.bar + U, ph);
        const double *src = c.pl + (U % NP) * SW;
        const double2 a = *reinterpret_cast<const double2 *>(src);
        const double2 b = *reinterpret_cast<const double2 *>(src + 64);
        W[U][0] = a.x; W[U][1] = a.y; W[U][2] = b.x; W[U][3] = b.y;
    } else {
        W[U][0] = W[U][1] = W[U][2] = W[U][3] = 0.0;
    }
    double fr_a = 0.0, fr_b = 0.0;  // red rhs of the row of the last red half-sweep
#pragma unroll
    for (int kk = 0; kk < T; kk++) {
        // ---- red half-sweep of sweep kk on row R - (2kk+1) ---------------------------------
        double carry_a, carry_b;
        const int lv = kk > 0 ? kk - 1 : 0;  // level of the late red residuals
        {
            const int lag = 2 * kk + 1;
            const int s = (U + 2 * NW - lag) % NW, sm = (s + NW - 1) % NW, sp = (s + 1) % NW;
            const int par = s & 1;  // row parity (rs has even global x, NW is even)
            const int q = R - lag;
            const double *rp = c.rl + s * SW;
            const double2 rA = *reinterpret_cast<const double2 *>(rp);
            const double2 rB = *reinterpret_cast<const double2 *>(rp + 64);
            const bool rv = STEADY || (unsigned)(q - c.x0) < nrows;
            double ta, tb;
            if (par == 0) {
                stencil2<0>(c, W[s], W[sm], W[sp], rA.x, rB.x, ta, tb);
                if (kk > 0 && rv) {  // residuals of sweep kk-1, one sweep late
                    const double ra = fma(-k.diag, W[s][0], ta), rb = fma(-k.diag, W[s][2], tb);
                    accA[lv] = fma(ra, ra, accA[lv]);
                    accB[lv] = fma(rb, rb, accB[lv]);
                }
                W[s][0] = fma(k.mid, ta, k.omw * W[s][0]);
                W[s][2] = fma(k.mid, tb, k.omw * W[s][2]);
                carry_a = rA.y; carry_b = rB.y;
                if (kk == T - 1) { fr_a = rA.x; fr_b = rB.x; }
            } else {
                stencil2<1>(c, W[s], W[sm], W[sp], rA.y, rB.y, ta, tb);
                if (kk > 0 && rv) {
                    const double ra = fma(-k.diag, W[s][1], ta), rb = fma(-k.diag, W[s][3], tb);
                    accA[lv] = fma(ra, ra, accA[lv]);
                    accB[lv] = fma(rb, rb, accB[lv]);
                }
                W[s][1] = fma(k.mid, ta, k.omw * W[s][1]);
                W[s][3] = fma(k.mid, tb, k.omw * W[s][3]);
                carry_a = rA.x; carry_b = rB.x;
                if (kk == T - 1) { fr_a = rA.y; fr_b = rB.y; }
            }
        }
        // ---- black half-sweep of sweep kk on row R - (2kk+2); rhs carried from last tick ---
        {
            const int lag = 2 * kk + 2;
            const int s = (U + 2 * NW - lag) % NW, sm = (s + NW - 1) % NW, sp = (s + 1) % NW;
            const int par = s & 1;
            const int q = R - lag;
            const bool rv = STEADY || (unsigned)(q - c.x0) < nrows;
            double ta, tb;
            if (par == 0) {  // black cells of an even row: the second cells
                stencil2<1>(c, W[s], W[sm], W[sp], C[kk][0], C[kk][1], ta, tb);
                const double pa = fma(k.mid, ta, k.omw * W[s][1]);
                const double pb = fma(k.mid, tb, k.omw * W[s][3]);
                W[s][1] = pa; W[s][3] = pb;
                if (rv) {
                    const double ra = fma(-k.diag, pa, ta), rb = fma(-k.diag, pb, tb);
                    accA[kk] = fma(ra, ra, accA[kk]);
                    accB[kk] = fma(rb, rb, accB[kk]);
                }
            } else {
                stencil2<0>(c, W[s], W[sm], W[sp], C[kk][0], C[kk][1], ta, tb);
                const double pa = fma(k.mid, ta, k.omw * W[s][0]);
                const double pb = fma(k.mid, tb, k.omw * W[s][2]);
                W[s][0] = pa; W[s][2] = pb;
                if (rv) {
                    const double ra = fma(-k.diag, pa, ta), rb = fma(-k.diag, pb, tb);
                    accA[kk] = fma(ra, ra, accA[kk]);
                    accB[kk] = fma(rb, rb, accB[kk]);
                }
            }
        }
        C[kk][0] = carry_a;
        C[kk][1] = carry_b;
    }
    // ---- residual of the red cells of the last sweep on row R - (2T+1); its red rhs values
    //      were read two ticks ago (FR[U & 1]) ------------------------------------------------
    {
        const int lag = 2 * T + 1;
        const int s = (U + 2 * NW - lag) % NW, sm = (s + NW - 1) % NW, sp = (s + 1) % NW;
        const int par = s & 1;
        const int q = R - lag;
        if (STEADY || (unsigned)(q - c.x0) < nrows) {  // warp-uniform
            double ta, tb;
            if (par == 0) {
                stencil2<0>(c, W[s], W[sm], W[sp], FR[U & 1][0], FR[U & 1][1], ta, tb);
                const double ra = fma(-k.diag, W[s][0], ta), rb = fma(-k.diag, W[s][2], tb);
                accA[T - 1] = fma(ra, ra, accA[T - 1]);
                accB[T - 1] = fma(rb, rb, accB[T - 1]);
            } else {
                stencil2<1>(c, W[s], W[sm], W[sp], FR[U & 1][0], FR[U & 1][1], ta, tb);
                const double ra = fma(-k.diag, W[s][1], ta), rb = fma(-k.diag, W[s][3], tb);
                accA[T - 1] = fma(ra, ra, accA[T - 1]);
                accB[T - 1] = fma(rb, rb, accB[T - 1]);
            }
        }
        FR[U & 1][0] = fr_a;
        FR[U & 1][1] = fr_b;
    }
    // ---- retire row R - (2T+2): nothing reads it any more ----------------------------------
    {
        const int lag = 2 * T + 2;
        const int s = (U + 2 * NW - lag) % NW;
        const int q = R - lag;
        if (STEADY || (unsigned)(q - c.x0) < nrows) {
            double *dst = c.pout + (roff - lag * c.pitch);
            if (c.cmA) stg_f64x2(dst, W[s][0], W[s][1]);
            if (c.cmB) stg_f64x2(dst + 64, W[s][2], W[s][3]);
        }
    }
}

template <int T>
__device__ __forceinline__ void stream_item(SCtx &c, int gpar, double *__restrict__ partial,
                                            int64_t part_stride) {
    constexpr int NW = stream_nw(T), NP = stream_np(T), PF = stream_pf(T);
    constexpr int HP = 2 * T + 2;
    c.first = c.x0 - HP;
    c.re = c.x1 + HP;
    // the tick loop starts on a row of even global x so that register slot parity = row parity
    const int rs = c.first - ((gpar + c.first) & 1);
    // ticks R in [st_lo, st_hi]: all rows R-1 .. R-(2T+2) lie in [x0, x1)
    const int st_lo = c.x0 + 2 * T + 2, st_hi = c.x1;
    if (c.lane == 0) {
        for (int i = 0; i < NW; i++) mbar_init(c.bar + i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // tick R requests row R + PF: rows below rs + PF are requested here
        for (int r = c.first; r < rs + PF; r++)
            if (r < c.re) issue_row<NP>(c, (int64_t)r * c.pitch, r - rs);
    }
    __syncwarp();
    double W[NW][4], C[T][2], FR[2][2], accA[T], accB[T];
#pragma unroll
    for (int i = 0; i < NW; i++) W[i][0] = W[i][1] = W[i][2] = W[i][3] = 0.0;
#pragma unroll
    for (int i = 0; i < T; i++) C[i][0] = C[i][1] = accA[i] = accB[i] = 0.0;
    FR[0][0] = FR[0][1] = FR[1][0] = FR[1][1] = 0.0;
    uint32_t ph = 0;
    int64_t roff = (int64_t)rs * c.pitch;
    for (int R0 = rs; R0 < c.re;) {
        if (R0 >= st_lo && R0 + NW - 1 <= st_hi) {
            do {  // the steady state: straight-line code, no row tests
#pragma unroll
                for (int U = 0; U < NW; U++) {
                    stream_tick<T, true>(W, C, FR, accA, accB, U, R0 + U, roff, ph, c);
                    roff += c.pitch;
                }
                ph ^= 1u;
                R0 += NW;
            } while (R0 + NW - 1 <= st_hi);
        } else {
#pragma unroll
            for (int U = 0; U < NW; U++) {
                if (R0 + U >= c.re) break;
                stream_tick<T, false>(W, C, FR, accA, accB, U, R0 + U, roff, ph, c);
                roff += c.pitch;
            }
            ph ^= 1u;
            R0 += NW;
        }
    }
#pragma unroll
    for (int i = 0; i < T; i++) {
        const double v = warp_sum_down((c.cmA ? accA[i] : 0.0) + (c.cmB ? accB[i] : 0.0));
        if (c.lane == 0) partial[(int64_t)i * part_stride] = v;
    }
}

// one warp per CTA, one work item per warp; TB = the configured temporal block (the lattice
// and the shared-memory budget follow it), the pass itself runs ctl->active_T <= TB sweeps
template <int TB>
__global__ void __launch_bounds__(32, stream_min_ctas(TB))
sor_rb_stream_kernel(const RbItem *__restrict__ items, double *const *__restrict__ pbuf,
                     const double *__restrict__ rhs, const SorCtl *__restrict__ ctl,
                     double *__restrict__ partial, int part_base, int part_stride, int64_t pitch,
                     int gpar, int h, RbConsts k) {
    const int T = ctl->active_T;
    if (T == 0) return;
    const int src = ctl->src;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const RbItem it = items[blockIdx.x];
    SCtx c;
    c.lane = threadIdx.x;
    c.lane_m1 = (c.lane + 31) & 31;
    c.lane_p1 = (c.lane + 1) & 31;
    // rings of the T actually run: p ring, rhs ring, one mbarrier per rhs slot
    c.pring = reinterpret_cast<double *>(smem_raw);
    c.rring = c.pring + stream_np(T) * SW;
    c.bar = reinterpret_cast<uint64_t *>(c.rring + stream_nw(T) * SW);
    c.pl = c.pring + 2 * c.lane;
    c.rl = c.rring + 2 * c.lane;
    c.pin = pbuf[src] + it.ty0;
    c.pout = pbuf[src ^ 1] + it.ty0 + 2 * c.lane;
    c.rhs = rhs + it.ty0;
    c.pitch = pitch;
    c.x0 = it.x0;
    c.x1 = it.x1;
    c.k = k;
    c.cmA = 2 * c.lane >= h && 2 * c.lane < SW - h;
    c.cmB = 64 + 2 * c.lane >= h && 64 + 2 * c.lane < SW - h;
    double *part = partial + part_base + blockIdx.x;
    if (T == TB) stream_item<TB>(c, gpar, part, part_stride);
    else if (TB > 1 && T == 1) stream_item<1>(c, gpar, part, part_stride);
    else if (TB > 2 && T == 2) stream_item<2>(c, gpar, part, part_stride);
    else if (TB > 3 && T == 3) stream_item<3>(c, gpar, part, part_stride);
}

using StreamKernel = void (*)(const RbItem *, double *const *, const double *, const SorCtl *,
                              double *, int, int, int64_t, int, int, RbConsts);
StreamKernel stream_kernel(int TB) {
    switch (TB) {
    case 1: return sor_rb_stream_kernel<1>;
    case 2: return sor_rb_stream_kernel<2>;
    case 3: return sor_rb_stream_kernel<3>;
    default: return sor_rb_stream_kernel<4>;
    }
}
int stream_smem_bytes(int TB) {
    switch (TB) {
    case 1: return stream_smem(1);
    case 2: return stream_smem(2);
    case 3: return stream_smem(3);
    default: return stream_smem(4);
    }
}

// plain[tile] = 1 iff every cell of the tile's footprint (inner region + h cells around it,
// 128 columns wide) is an interior fluid cell without a non-fluid neighbour
__global__ void rb_plain_kernel(const uint8_t *__restrict__ cflag, Geom g, int tiles_y, int BX,
                                int BY, int h, uint8_t *__restrict__ plain) {
    const int tile = blockIdx.x;
    const int ti = tile / tiles_y, tj = tile - ti * tiles_y;
    const int64_t x0 = g.own0 + (int64_t)ti * BX;
    const int64_t x1 = min(x0 + BX, g.own1);
    const int64_t col = (int64_t)tj * BY - h + threadIdx.x;
    int ok = col >= 1 && col <= g.NY - 2;
    if (ok) {
        for (int64_t r = x0 - h; r < x1 + h; r++) {
            const int64_t gx = g.gx0 + r;
            if (r < 0 || r >= g.nxl || gx < 1 || gx > g.NX - 2 ||
                cflag[r * g.pitch + col] != CF_FLUID) {
                ok = 0;
                break;
            }
        }
    }
    ok = __syncthreads_and(ok);
    if (threadIdx.x == 0) plain[tile] = (uint8_t)ok;
}

}  // namespace

void rb_plan_release(sb_sim *s) {
    cudaFree(s->plan.d_slow);
    cudaFree(s->plan.d_items);
    cudaFree(s->plan.d_plain);
    s->plan = RbPlan();
}

// (Re)build the split of the tile lattice into slow tiles and streaming work items when the
// cell flags or the temporal block have changed since the last build.
sb_status rb_ensure_plan(sb_sim *s, int BX, int BY, int h) {
    RbPlan &pl = s->plan;
    const int T = s->prm.temporal_block;
    if (pl.epoch == s->flag_epoch && pl.T == T) return SB_OK;
    const Geom &g = s->g;
    const int tiles_x = (int)((g.own1 - g.own0 + BX - 1) / BX);
    const int tiles_y = (int)((g.NY + BY - 1) / BY);
    const int ntiles = tiles_x * tiles_y;
    if ((size_t)ntiles > pl.cap_tiles) {
        if (pl.d_slow) SB_CUDA(cudaFreeAsync(pl.d_slow, s->stream));
        if (pl.d_plain) SB_CUDA(cudaFreeAsync(pl.d_plain, s->stream));
        if (pl.d_items) SB_CUDA(cudaFreeAsync(pl.d_items, s->stream));
        pl.d_slow = nullptr; pl.d_plain = nullptr; pl.d_items = nullptr;
        pl.cap_tiles = 0;
        SB_CUDA(cudaMallocAsync(&pl.d_slow, (size_t)ntiles * sizeof(int32_t), s->stream));
        SB_CUDA(cudaMallocAsync(&pl.d_plain, (size_t)ntiles, s->stream));
        SB_CUDA(cudaMallocAsync(&pl.d_items, (size_t)ntiles * sizeof(RbItem), s->stream));
        pl.cap_tiles = (size_t)ntiles;
    }
    rb_plain_kernel<<<ntiles, SW, 0, s->stream>>>(s->cflag, g, tiles_y, BX, BY, h, pl.d_plain);
    s->launches++;
    SB_CUDA(cudaGetLastError());
    std::vector<uint8_t> plain((size_t)ntiles);
    SB_CUDA(cudaMemcpyAsync(plain.data(), pl.d_plain, (size_t)ntiles, cudaMemcpyDeviceToHost,
                            s->stream));
    SB_CUDA(cudaStreamSynchronize(s->stream));
    if (s->slab) {
        // tiles that own rows within H of a slab edge also feed the neighbour's halo rows:
        // that code lives in the tile kernel
        const int H = s->link.H;
        for (int ti = 0; ti < tiles_x; ti++) {
            const int64_t x0 = g.own0 + (int64_t)ti * BX, x1 = std::min<int64_t>(x0 + BX, g.own1);
            const bool lo = s->link.lo_p[0] != nullptr && x0 < g.own0 + H;
            const bool hi = s->link.hi_p[0] != nullptr && x1 > g.own1 - H;
            if (lo || hi)
                for (int tj = 0; tj < tiles_y; tj++) plain[(size_t)ti * tiles_y + tj] = 0;
        }
    }
    // runs of plain tiles along x, per strip
    struct Run { int tj, ti0, len; };
    std::vector<Run> runs;
    std::vector<int32_t> slow;
    for (int tj = 0; tj < tiles_y; tj++) {
        int ti = 0;
        while (ti < tiles_x) {
            if (!plain[(size_t)ti * tiles_y + tj]) { ti++; continue; }
            int t0 = ti;
            while (ti < tiles_x && plain[(size_t)ti * tiles_y + tj]) ti++;
            runs.push_back({tj, t0, ti - t0});
        }
    }
    for (int t = 0; t < ntiles; t++)
        if (!plain[(size_t)t]) slow.push_back(t);
    // tiles per item: as many items as keep every SM's warps busy in whole waves, as long as
    // possible otherwise (every item pays 2(2T+2) warm-up rows)
    static bool carveout_set = false;
    if (!carveout_set) {
        for (int tb = 1; tb <= RB_TMAX; tb++)
            cudaFuncSetAttribute(stream_kernel(tb), cudaFuncAttributePreferredSharedMemoryCarveout,
                                 cudaSharedmemCarveoutMaxShared);
        carveout_set = true;
    }
    int dev_sms = 148, ctas = 8;
    cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, s->device);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, stream_kernel(T), 32,
                                                  stream_smem_bytes(T));
    if (ctas < 1) ctas = 1;
    const int64_t resident = (int64_t)dev_sms * ctas;
    int best_seg = 1;
    double best_cost = 1e300;
    for (int seg = 1; seg <= 256; seg++) {
        int64_t n = 0;
        int longest = 0;
        for (const Run &r : runs) {
            const int m = (r.len + seg - 1) / seg;
            n += m;
            longest = std::max(longest, (r.len + m - 1) / m);
        }
        if (n == 0) break;
        const int64_t waves = (n + resident - 1) / resident;
        const double cost = (double)waves * ((double)longest * BX + 2.0 * h);
        if (cost < best_cost) { best_cost = cost; best_seg = seg; }
    }
    std::vector<RbItem> items;
    for (const Run &r : runs) {
        const int m = (r.len + best_seg - 1) / best_seg;
        for (int i = 0; i < m; i++) {
            const int a = (int)((int64_t)r.len * i / m), b = (int)((int64_t)r.len * (i + 1) / m);
            RbItem it;
            it.x0 = (int32_t)(g.own0 + (int64_t)(r.ti0 + a) * BX);
            it.x1 = (int32_t)std::min<int64_t>(g.own0 + (int64_t)(r.ti0 + b) * BX, g.own1);
            it.ty0 = r.tj * BY - h;
            it.pad = 0;
            items.push_back(it);
        }
    }
    // neighbouring strips of the same rows run side by side: their shared halo columns are
    // then fetched from HBM once
    std::stable_sort(items.begin(), items.end(), [](const RbItem &a, const RbItem &b) {
        return a.x0 != b.x0 ? a.x0 < b.x0 : a.ty0 < b.ty0;
    });
    if (!slow.empty())
        SB_CUDA(cudaMemcpyAsync(pl.d_slow, slow.data(), slow.size() * sizeof(int32_t),
                                cudaMemcpyHostToDevice, s->stream));
    if (!items.empty())
        SB_CUDA(cudaMemcpyAsync(pl.d_items, items.data(), items.size() * sizeof(RbItem),
                                cudaMemcpyHostToDevice, s->stream));
    SB_CUDA(cudaStreamSynchronize(s->stream));  // the vectors go out of scope
    pl.n_slow = (int)slow.size();
    pl.n_items = (int)items.size();
    pl.tiles_x = tiles_x;
    pl.tiles_y = tiles_y;
    pl.T = T;
    pl.epoch = s->flag_epoch;
    return SB_OK;
}

sb_status launch_sor_rb_stream(sb_sim *s, int part_base, int part_stride, int h) {
    const Geom &g = s->g;
    const int gpar = (int)(((g.gx0 % 2) + 2) % 2);
    const int TB = s->prm.temporal_block;
    stream_kernel(TB)<<<s->plan.n_items, 32, stream_smem_bytes(TB), s->stream>>>(
        s->plan.d_items, rb_pbuf_ptr(s), s->rhs, s->d_ctl, s->d_partial, part_base, part_stride,
        g.pitch, gpar, h, rb_consts(s));
    s->launches++;
    SB_CUDA(cudaGetLastError());
    return SB_OK;
}

void preload_sor_rb_stream() {
    cudaFuncAttributes a;
    for (int tb = 1; tb <= RB_TMAX; tb++) cudaFuncGetAttributes(&a, stream_kernel(tb));
    cudaFuncGetAttributes(&a, rb_plain_kernel);
    cudaGetLastError();
}

}  // namespace sb
