// sor_rb_stream.cu -- K4c: the performance-mode SOR pass on regular regions, as a
// register-resident row pipeline ("2.5-D" temporal blocking).
//
// Same arithmetic as the tile kernel (sor_rb.cuh; the red-black restatement of the sweep of
// /root/reference/src/simulation.rs:253-274, of copy_pressure_to_boundaries,
// src/grid/mod.rs:343-412, and of calculate_norm_squared, simulation.rs:216-227), other
// schedule.  Where a piece of the grid holds nothing but fluid cells -- plus, optionally, ONE
// straight wall along x on the strip's first or last column and a boundary row (inflow,
// outflow, wall) at either end in x -- T sweeps can be pipelined along x by ONE WARP without
// any block-level synchronisation:
//
//   * a warp owns a 128-column strip (4 columns per lane) and walks down x;
//   * in the tick in which row R arrives it runs, for k = 0..T-1, the red half-sweep of
//     sweep k on row R-(2k+1) and the black half-sweep on row R-(2k+2) (each needs its two
//     neighbour rows one half-sweep behind -- true in this order), then the residual of the
//     red cells of the last sweep on row R-(2T+1), and retires row R-(2T+2) to HBM;
//   * the 2T+4 rows in flight live in REGISTERS (the tick loop is unrolled over the window so
//     every row has a fixed register name); the only exchange between lanes is one shuffle
//     per half-sweep (the y-neighbour across the lane boundary);
//   * rows of p and rhs are fetched PF rows ahead by 1-D bulk copies of the TMA engine into
//     per-item shared-memory rings (one mbarrier per slot);
//   * at T = 4 the strip is walked by a CHAIN of two such warps (HEAD: sweeps 0-1, TAIL: sweeps
//     2-3, rows handed over through a shared-memory ring; two rows per tick) -- see "two warps
//     per work item" below;
//   * residuals ride along exactly as in the tile kernel: black cells at their update, red
//     cells of sweep k inside the red half-sweep of sweep k+1.
//
// Boundary cells inside an item.  The reference copies pressures into the boundary cells
// once per iteration, BEFORE the sweep (simulation.rs:251).  A wall cell (x, 0) takes
// p(x, 1); it sits in the register window like any other cell (lane 0's first / lane 31's
// last cell), is refreshed at the start of the red half-sweep of its row, never updated and
// never counted.  A boundary row at the end of an item (grid row 0 or NX-1) is refreshed from
// its neighbour row at the start of that row's red half-sweep.  The one subtlety is the late
// red residual: the residual of sweep k-1 of a red cell next to a boundary cell must see the
// boundary value of sweep k-1, so in those half-sweeps the stencil sum is taken first with
// the old boundary value (residual), then the boundary cell is refreshed and the affected sum
// retaken (update).  Walls are a template parameter (plain strips carry none of this code);
// boundary rows only exist in the warm-up / drain path of an item.
//
// Halo: a strip carries h = 2T+2 columns per open side like a tile, but along x only the
// open ends of a work item pay 2T+2 warm-up rows -- a tile pays them every 48 rows.
//
// HBM traffic per item and cell: 8 (p) + 8 (rhs) + 8 (p out) = 24 B for T sweeps (+ the
// strip halo re-reads, L2 hits when neighbouring strips run side by side).
//
// What kind of item a lattice tile can join is decided per tile by rb_class_kernel; the host
// turns runs of equal tiles along x into work items (RbPlan).  Everything else -- obstacles,
// corners of walls, slab edges -- stays with the tile kernel.
#include <limits.h>
#include <stdlib.h>

#include <algorithm>

#include "slab_dev.cuh"
#include "sor_rb.cuh"

namespace sb {

namespace {

constexpr int SW = RB_TW;      // strip columns: 32 lanes x 2 pairs of columns
constexpr int ROW_BYTES = SW * 8;

// item kinds (RbItem::pad bits 0-1) and flags
constexpr int IT_PLAIN = 0, IT_WALL_LO = 1, IT_WALL_HI = 2;
constexpr int IT_BC_LO = 4;    // the item's first row is a boundary row fed from the next row
constexpr int IT_BC_HI = 8;    // the item's last row is a boundary row fed from the row before
// tile classes of rb_class_kernel: 0 = tile kernel, 1 + item kind otherwise, plus
constexpr int TC_BC_LO = 4, TC_BC_HI = 8;  // the tile owns grid row 0 / NX-1 as such a row
constexpr int TC_FROZEN = 16;  // not streamable, but nothing in or next to the tile ever changes
// the footprint reaches grid row NX-1 without the tile owning it (the last tile row of the
// lattice is shorter than the halo): streamable only as part of a run that goes on into the
// next tile, whose item then owns that boundary row
constexpr int TC_NEED_NEXT = 32;

// Everything that indexes a ring is a compile-time constant inside the unrolled window:
//   * the tick loop is unrolled over NW = 2T+4 ticks (the register window), U = tick mod NW;
//   * row R lands in slot U of the rhs ring (one mbarrier per slot) and in slot U mod (T+2)
//     of the p ring (a p row is consumed in its arrival tick);
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

// ---- two warps per work item at TB = 4 ("chain") -------------------------------------------
// A warp that keeps the 2T+4 rows of T = 4 sweeps in registers needs 246 of them (8 warps per
// SM, two per scheduler) and an unrolled loop of 46 KB against 32 KB of instruction cache; its
// eight half-sweeps per tick form one dependent chain.  Measured on the B200 (profiles/r2_*):
// the memory pipeline of the pass alone takes 0.29 ms at 8192^2, the whole pass 0.39 ms, the
// difference being warps that wait on their own arithmetic (stall samples: no instruction
// 22 %, fixed-latency wait 22 %).  So at TB = 4 an item is run by TWO warps, each with the
// window of T = 2:
//   HEAD  requests the rows (TMA), runs sweeps 0-1 and hands every row it retires to
//   TAIL  through a 4-row shared-memory ring (two mbarriers per slot: full / empty); TAIL runs
//         sweeps 2-3, takes the late residuals and stores the row (and a slab's edge rows).
// TAIL reads its rhs rows from HEAD's rhs ring, which therefore keeps three windows' worth of
// rows (banks that rotate once per loop iteration; the hand-over ring's back pressure keeps
// HEAD from overwriting a row TAIL still reads).  Both run the T = 2 code: 20 KB of loop,
// 164 registers, 12 warps per SM, chains of four half-sweeps -- and the HBM traffic of T = 4.
constexpr int ROLE_SOLO = 0, ROLE_HEAD = 1, ROLE_TAIL = 2;
constexpr int CHAIN_TW = 2;        // sweeps per warp of a chain
constexpr int CHAIN_D = 4;         // rows in the hand-over ring (divides stream_nw(CHAIN_TW))
constexpr int CHAIN_BANKS = 3;     // rhs ring = 3 windows of HEAD
// A chain warp's tick takes TWO rows (stream_tick2): half the ticks, and the two rows of a
// half-sweep are independent instruction streams inside one warp.  Rows are requested and
// handed over in aligned pairs (one mbarrier per pair).
constexpr int CHAIN_NP = stream_nw(CHAIN_TW);   // p ring of a chain's HEAD: pairs stay aligned
#ifndef SB_CHAIN_PF
#define SB_CHAIN_PF 4
#endif
constexpr int CHAIN_PF = SB_CHAIN_PF;           // rows requested ahead (even)
static_assert(CHAIN_PF % 2 == 0 && CHAIN_PF + 2 <= CHAIN_NP, "p ring: rows in flight + the arriving pair");
static_assert(CHAIN_D % 2 == 0 && stream_nw(CHAIN_TW) % CHAIN_D == 0, "hand-over ring of whole pairs");
// HEAD in tick R requests rows up to R + PF + 1 and may run 2 TW + D + 2 rows ahead of TAIL,
// which still reads the rhs of row R' - 2 TW + 1 in its tick R'
static_assert(CHAIN_PF + 2 + 2 * CHAIN_TW + CHAIN_D + 2 + 2 * CHAIN_TW <= CHAIN_BANKS * stream_nw(CHAIN_TW),
              "rhs ring too short for the chain");
__host__ __device__ constexpr bool stream_chained(int TB) { return TB == 4; }
// warps of a CTA / work items of a CTA.  The register file is split per SM sub-partition
// (16 K registers each), so the steps are 16 warps (128 registers per thread), 12 (168) and
// 8 (255)
#ifndef SB_CHAIN_WARPS
#define SB_CHAIN_WARPS 12
#endif
__host__ __device__ constexpr int stream_warps(int TB) {
    return TB == 1 ? 16 : TB == 4 ? SB_CHAIN_WARPS : 12;
}
__host__ __device__ constexpr int stream_items_per_cta(int TB) {
    return stream_chained(TB) ? stream_warps(TB) / 2 : stream_warps(TB);
}
// partial-sum slots per item (one per warp that works on it)
__host__ __device__ constexpr int stream_slots_per_item(int TB) { return stream_chained(TB) ? 2 : 1; }
// shared memory of one work item (rings + mbarriers), padded to the 128-byte ring alignment.
//   solo:  p ring T+2 rows, rhs ring 2T+4 rows, 2T+4 mbarriers
//   chain: p ring, 3 x rhs window, hand-over ring, arrival + full + empty mbarriers; the
//          shortened passes (T < 4: one warp of the pair runs them alone) fit inside
__host__ __device__ constexpr int stream_smem(int TB) {
    return stream_chained(TB)
               ? ((CHAIN_NP + CHAIN_BANKS * stream_nw(CHAIN_TW) + CHAIN_D) * ROW_BYTES +
                  (stream_nw(CHAIN_TW) + 2 * CHAIN_D) * 8 + 127) / 128 * 128
               : ((stream_np(TB) + stream_nw(TB)) * ROW_BYTES + stream_nw(TB) * 8 + 127) / 128 * 128;
}
// All warps of an SM form ONE CTA (apart from the pairs of a chain they never synchronise with
// each other) whose items are all of the same kind (plain / wall strip), so an SM runs one
// copy of the unrolled window (profiles/r1_stream_kinds_ab.txt: +8 % against one-warp CTAs).

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
    return pred != 0;
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
    int x0, x1;                // rows stored
    int cx0, cx1;              // rows swept and counted (the stored rows minus boundary rows)
    int bc_lo, bc_hi;          // boundary rows of the item (far out of range if none)
    int first, re;             // rows loaded: [first, re)
    int rend;                  // ticks run for rows [rs, rend)
    int rs, hend;              // chain: rows [rs, hend) go through the hand-over ring, in order
    int group_bar;             // chain: named barrier of the warp pair
    int lane, lane_m1, lane_p1;
    bool cmA, cmB;             // the pair lies in the strip's stored / counted columns
    bool keepA, keepB;         // A0 / B1 of this lane is a wall cell
    double *pring;             // p ring base (HEAD / SOLO)
    double *rring;             // rhs ring base: bank 0
    const double *pl;          // this lane's pair A in slot 0 of the p ring
    // this lane's pair A in slot 0 of the rhs bank of this loop iteration / the one before;
    // the bank the requests of this iteration's last PF ticks go to (HEAD / SOLO)
    const double *rl_cur, *rl_prev;
    double *rr_cur, *rr_next;
    uint64_t *bar;             // arrival barriers of the rhs slots of one window
    double *hring;             // hand-over ring (chain), this lane's pair A in slot 0
    uint64_t *hfull, *hempty;
    RbConsts k;
};

// What the streaming kernel needs to know about the neighbouring slabs (SlabLink, slab.cu).
// Row slabs: the rows within H of a slab edge are also stored into the neighbour's halo rows
// of ITS target buffer (P2P over NVLink).  Lives in the kernel's parameter space (constant
// bank) and is only touched on the warm-up / drain path: no registers in the steady loop.
struct StreamPeers {
    double *lo_p[2], *hi_p[2];   // the neighbours' two pressure buffers (nullptr: none)
    int64_t lo_row0, hi_row0;    // local row of THEIR array that receives my first / last H rows
    int own0, own1, H;
    double *const *pbuf;         // my own two buffers and the control block (for src)
    const SorCtl *ctl;
};

// request row at offset `off` (both arrays): rhs into slot `slot` of bank `rbank`, p into slot
// `slot % NP`; lane 0 only
template <int NP>
__device__ __forceinline__ void issue_row(const SCtx &c, double *rbank, int64_t off, int slot) {
    uint64_t *bar = c.bar + slot;
    mbar_expect_tx(bar, 2 * ROW_BYTES);
    bulk_load(c.pring + (slot % NP) * SW, c.pin + off, ROW_BYTES, bar);
    bulk_load(rbank + slot * SW, c.rhs + off, ROW_BYTES, bar);
}

// request the rows at offset `off` and one pitch further (both arrays) into rhs slots `slot`,
// `slot + 1` of bank `rbank` and the same slots of the p ring: ONE mbarrier (that of `slot`)
// for the four copies; v0 / v1: the row exists.  One lane only.
__device__ __forceinline__ void issue_rows2(const SCtx &c, double *rbank, int64_t off, int slot,
                                            bool v0, bool v1) {
    uint64_t *bar = c.bar + slot;
    mbar_expect_tx(bar, ((v0 ? 1 : 0) + (v1 ? 1 : 0)) * 2 * ROW_BYTES);
    if (v0) {
        bulk_load(c.pring + slot * SW, c.pin + off, ROW_BYTES, bar);
        bulk_load(rbank + slot * SW, c.rhs + off, ROW_BYTES, bar);
    }
    if (v1) {
        bulk_load(c.pring + (slot + 1) * SW, c.pin + off + c.pitch, ROW_BYTES, bar);
        bulk_load(rbank + (slot + 1) * SW, c.rhs + off + c.pitch, ROW_BYTES, bar);
    }
}

// The two cells of one colour in a lane's two column pairs.  SET 0: the first cells A0, B0
// (cells 0, 2; y-neighbours: the lane's own second cell and the second cell of the pair to
// the left), SET 1: the second cells A1, B1 (cells 1, 3; own first cell, first cell of the
// pair to the right).  Pair B of lane 0 continues pair A of lane 31 and vice versa.
// nbr2: the two foreign y-neighbours.
template <int SET>
__device__ __forceinline__ void nbr2(const SCtx &c, const double (&me)[4], double &na,
                                     double &nb) {
    if (SET == 0) {
        const double la = __shfl_sync(0xffffffffu, me[1], c.lane_m1);
        const double lb0 = __shfl_sync(0xffffffffu, me[3], c.lane_m1);
        na = la;
        nb = c.lane == 0 ? la : lb0;
    } else {
        const double ra0 = __shfl_sync(0xffffffffu, me[0], c.lane_p1);
        const double rb = __shfl_sync(0xffffffffu, me[2], c.lane_p1);
        na = c.lane == 31 ? rb : ra0;
        nb = rb;
    }
}
// t = fma(1/dx^2, pE + pW, fma(1/dy^2, pS + pN, -rhs))
__device__ __forceinline__ double tsum(const RbConsts &k, double xs, double ys, double rh) {
    return fma(k.rdx2, xs, fma(k.rdy2, ys, -rh));
}

// One half-sweep (RED: first colour) on the row in register slot s, cells SET; g = the sweep's
// index in the PASS (a TAIL warp starts at g = 2).
//   RED:   rhs from the ring (the other colour's values are carried to the black half-sweep
//          of the next tick); the residuals of sweep g-1 of these cells are taken here, one
//          sweep late (LATE = g > 0, into acc_*); boundary cells are refreshed.
//   black: rhs carried; residuals of sweep g right after the update.
// q = the row; STEADY: q is an ordinary counted row (no row tests).
template <int T, bool STEADY, int WALL, int SET, bool RED>
__device__ __forceinline__ void half_sweep(double (&W)[2 * T + 4][4], const int s, const int q,
                                           const int g, const double rha, const double rhb,
                                           double &acc_a, double &acc_b, const SCtx &c) {
#ifdef SB_STREAM_NOCOMPUTE   // experiment: the memory pipeline of the pass alone (wrong results)
    return;
#endif
    constexpr int NW = stream_nw(T);
    constexpr int ia = SET, ib = SET + 2, oa = SET ^ 1, ob = (SET ^ 1) + 2;
    // the wall cell is one of these cells / the fluid cell next to the wall is
    // WALL: the strip has a wall column -- cell 0 of lane 0 (keepA) or cell 3 of lane 31
    // (keepB), told apart at run time so that both share one copy of the code
    constexpr bool wall_a = WALL && SET == 0;   // cell 0 may be the wall cell
    constexpr bool wall_b = WALL && SET == 1;   // cell 3 may be the wall cell
    constexpr bool adj_a = WALL && SET == 1;    // cell 1 may be the fluid cell next to it
    constexpr bool adj_b = WALL && SET == 0;    // cell 2 may be
    const RbConsts &k = c.k;
    const int sm = (s + NW - 1) % NW, sp = (s + 1) % NW;
    if (!STEADY && (q == c.bc_lo || q == c.bc_hi)) return;   // a boundary row: not swept
    const bool rv = STEADY || (unsigned)(q - c.cx0) < (unsigned)(c.cx1 - c.cx0);
    const bool late = RED && g > 0;
    // boundary refresh of this iteration (red half-sweep only): wall cell <- its fluid
    // neighbour; boundary row <- this row (corner cells on the wall column keep their value)
    const bool row_lo = !STEADY && RED && q == c.bc_lo + 1;
    const bool row_hi = !STEADY && RED && q == c.bc_hi - 1;
    auto refresh_wall = [&]() {
        if (WALL) {
            W[s][0] = c.keepA ? W[s][1] : W[s][0];
            W[s][3] = c.keepB ? W[s][2] : W[s][3];
        }
    };
    auto refresh_rows = [&]() {
        if (row_lo) {
            W[sm][0] = (WALL && c.keepA) ? W[sm][0] : W[s][0];
            W[sm][1] = W[s][1];
            W[sm][2] = W[s][2];
            W[sm][3] = (WALL && c.keepB) ? W[sm][3] : W[s][3];
        }
        if (row_hi) {
            W[sp][0] = (WALL && c.keepA) ? W[sp][0] : W[s][0];
            W[sp][1] = W[s][1];
            W[sp][2] = W[s][2];
            W[sp][3] = (WALL && c.keepB) ? W[sp][3] : W[s][3];
        }
    };
    if (RED && !late) {  // first sweep of the pass: nothing is late, refresh first
        refresh_wall();
        if (!STEADY) refresh_rows();
    }
    double na, nb;
    nbr2<SET>(c, W[s], na, nb);
    double xa = W[sp][ia] + W[sm][ia], xb = W[sp][ib] + W[sm][ib];
    double ta = tsum(k, xa, W[s][oa] + na, rha), tb = tsum(k, xb, W[s][ob] + nb, rhb);
    if (late) {
        if (rv) {  // residuals of sweep g-1 with the boundary values of sweep g-1
            double ra = fma(-k.diag, W[s][ia], ta), rb = fma(-k.diag, W[s][ib], tb);
            if (wall_a) ra = c.keepA ? 0.0 : ra;
            if (wall_b) rb = c.keepB ? 0.0 : rb;
            acc_a = fma(ra, ra, acc_a);
            acc_b = fma(rb, rb, acc_b);
        }
        if (adj_a || adj_b) {  // now the wall cell moves on to sweep g: retake the sum
            refresh_wall();
            if (adj_a) ta = tsum(k, xa, W[s][oa] + na, rha);
            if (adj_b) tb = tsum(k, xb, W[s][ob] + nb, rhb);
        }
        if (!STEADY && (row_lo || row_hi)) {  // warp-uniform
            refresh_rows();
            xa = W[sp][ia] + W[sm][ia];
            xb = W[sp][ib] + W[sm][ib];
            ta = tsum(k, xa, W[s][oa] + na, rha);
            tb = tsum(k, xb, W[s][ob] + nb, rhb);
        }
    }
    double pa = fma(k.mid, ta, k.omw * W[s][ia]), pb = fma(k.mid, tb, k.omw * W[s][ib]);
    if (wall_a) pa = c.keepA ? W[s][ia] : pa;
    if (wall_b) pb = c.keepB ? W[s][ib] : pb;
    W[s][ia] = pa;
    W[s][ib] = pb;
    if (!RED && rv) {
        double ra = fma(-k.diag, pa, ta), rb = fma(-k.diag, pb, tb);
        if (wall_a) ra = c.keepA ? 0.0 : ra;
        if (wall_b) rb = c.keepB ? 0.0 : rb;
        acc_a = fma(ra, ra, acc_a);
        acc_b = fma(rb, rb, acc_b);
    }
}

// One tick: row R has been requested PF ticks ago.  U = (R - rs) mod NW is the register slot
// of row R and its ring slot; after inlining into the unrolled loop every index below is a
// constant.  roff = R * pitch.  ph = parity of the loop iteration `it`; KB = index in the
// pass of this warp's first sweep.  acc*[g - AB] collects level g, AB = max(KB - 1, 0).
// STEADY: every row this tick touches is an ordinary counted row (no row tests at all) and
// (chain) every hand-over slot has been used before.
template <int T, bool STEADY, int WALL, int ROLE, int KB>
__device__ __forceinline__ void stream_tick(double (&W)[2 * T + 4][4], double (&C)[T][2],
                                            double (&FR)[2][2], double (&accA)[T + 1],
                                            double (&accB)[T + 1], const int U, const int R,
                                            const int64_t roff, const uint32_t ph, const int it,
                                            const SCtx &c, const StreamPeers &pe) {
    constexpr int NW = stream_nw(T), NP = stream_np(T), PF = stream_pf(T);
    constexpr int AB = KB > 0 ? KB - 1 : 0;
    constexpr int D = CHAIN_D;
    const RbConsts &k = c.k;
    // ---- row R: requested (HEAD / SOLO: row R + PF goes out) or handed over -> registers ------
    __syncwarp();  // every lane is done with the slots the request overwrites
    if constexpr (ROLE != ROLE_TAIL) {
        if (STEADY) {
            // the warp is converged here: elect.sync lets the copies issue from straight-line
            // code (a `lane == 0` branch makes ptxas wrap each UBLKCP in an election loop)
            if (elect_one())
                issue_row<NP>(c, U + PF < NW ? c.rr_cur : c.rr_next, roff + PF * c.pitch, (U + PF) % NW);
        } else if (c.lane == 0) {
            if (R + PF < c.re)
                issue_row<NP>(c, U + PF < NW ? c.rr_cur : c.rr_next, roff + PF * c.pitch, (U + PF) % NW);
            if (R < c.first) mbar_arrive(c.bar + U);  // keeps the phases in step
        }
        if (STEADY || (R >= c.first && R < c.re)) {
            mbar_wait(c.bar + U, ph);
            const double *src = c.pl + (U % NP) * SW;
            const double2 a = *reinterpret_cast<const double2 *>(src);
            const double2 b = *reinterpret_cast<const double2 *>(src + 64);
            W[U][0] = a.x; W[U][1] = a.y; W[U][2] = b.x; W[U][3] = b.y;
        } else {
            W[U][0] = W[U][1] = W[U][2] = W[U][3] = 0.0;
        }
    } else {
        // every row from rs on takes its turn in the ring (rows before `first` carry zeros),
        // so that use n of a slot is row rs + n D + slot for both warps
        if (STEADY || R < c.hend) {
            // the n-th use of slot U % D, n = it * (NW / D) + U / D: parity (U / D) & 1
            static_assert((NW / D) % 2 == 0, "static hand-over phases need an even NW / D");
            mbar_wait(c.hfull + U % D, (uint32_t)((U / D) & 1));
            const double *src = c.hring + (U % D) * SW;
            const double2 a = *reinterpret_cast<const double2 *>(src);
            const double2 b = *reinterpret_cast<const double2 *>(src + 64);
            W[U][0] = a.x; W[U][1] = a.y; W[U][2] = b.x; W[U][3] = b.y;
            __syncwarp();
            if (c.lane == 0) mbar_arrive(c.hempty + U % D);
        } else {
            W[U][0] = W[U][1] = W[U][2] = W[U][3] = 0.0;
        }
    }
    double fr_a = 0.0, fr_b = 0.0;  // red rhs of the row of the last red half-sweep
#pragma unroll
    for (int kk = 0; kk < T; kk++) {
        double carry_a, carry_b;
        const int g = KB + kk;                      // sweep index in the pass
        const int lv = (g > 0 ? g - 1 : 0) - AB;    // accumulator of the late red residuals
        {   // ---- red half-sweep of sweep kk on row R - (2kk+1) -----------------------------
            const int lag = 2 * kk + 1;
            const int s = (U + 2 * NW - lag) % NW;
            const double *rp = (U >= lag ? c.rl_cur : c.rl_prev) + s * SW;
            const double2 rA = *reinterpret_cast<const double2 *>(rp);
            const double2 rB = *reinterpret_cast<const double2 *>(rp + 64);
            if ((s & 1) == 0) {  // row parity (rs has even global x, NW is even)
                half_sweep<T, STEADY, WALL, 0, true>(W, s, R - lag, g, rA.x, rB.x, accA[lv],
                                                     accB[lv], c);
                carry_a = rA.y; carry_b = rB.y;
                if (kk == T - 1) { fr_a = rA.x; fr_b = rB.x; }
            } else {
                half_sweep<T, STEADY, WALL, 1, true>(W, s, R - lag, g, rA.y, rB.y, accA[lv],
                                                     accB[lv], c);
                carry_a = rA.x; carry_b = rB.x;
                if (kk == T - 1) { fr_a = rA.y; fr_b = rB.y; }
            }
        }
        {   // ---- black half-sweep of sweep kk on row R - (2kk+2); rhs carried one tick ------
            const int lag = 2 * kk + 2;
            const int s = (U + 2 * NW - lag) % NW;
            if ((s & 1) == 0)  // black cells of an even row: the second cells
                half_sweep<T, STEADY, WALL, 1, false>(W, s, R - lag, g, C[kk][0], C[kk][1],
                                                      accA[g - AB], accB[g - AB], c);
            else
                half_sweep<T, STEADY, WALL, 0, false>(W, s, R - lag, g, C[kk][0], C[kk][1],
                                                      accA[g - AB], accB[g - AB], c);
        }
        C[kk][0] = carry_a;
        C[kk][1] = carry_b;
    }
    // ---- residual of the red cells of the LAST sweep of the pass on row R - (2T+1); its red
    //      rhs values were read two ticks ago (FR[U & 1]); no boundary cell has been refreshed
    //      since.  (A HEAD warp leaves the late residuals of its last sweep to TAIL.) ----------
    if constexpr (ROLE != ROLE_HEAD) {
        const int lag = 2 * T + 1;
        const int s = (U + 2 * NW - lag) % NW, sm = (s + NW - 1) % NW, sp = (s + 1) % NW;
        const int q = R - lag;
#ifdef SB_STREAM_NOCOMPUTE
        if (false) {
#else
        if (STEADY || (unsigned)(q - c.cx0) < (unsigned)(c.cx1 - c.cx0)) {  // warp-uniform
#endif
            double na, nb, ra, rb;
            if ((s & 1) == 0) {
                nbr2<0>(c, W[s], na, nb);
                ra = fma(-k.diag, W[s][0],
                         tsum(k, W[sp][0] + W[sm][0], W[s][1] + na, FR[U & 1][0]));
                rb = fma(-k.diag, W[s][2],
                         tsum(k, W[sp][2] + W[sm][2], W[s][3] + nb, FR[U & 1][1]));
                if (WALL) ra = c.keepA ? 0.0 : ra;
            } else {
                nbr2<1>(c, W[s], na, nb);
                ra = fma(-k.diag, W[s][1],
                         tsum(k, W[sp][1] + W[sm][1], W[s][0] + na, FR[U & 1][0]));
                rb = fma(-k.diag, W[s][3],
                         tsum(k, W[sp][3] + W[sm][3], W[s][2] + nb, FR[U & 1][1]));
                if (WALL) rb = c.keepB ? 0.0 : rb;
            }
            accA[KB + T - 1 - AB] = fma(ra, ra, accA[KB + T - 1 - AB]);
            accB[KB + T - 1 - AB] = fma(rb, rb, accB[KB + T - 1 - AB]);
        }
        FR[U & 1][0] = fr_a;
        FR[U & 1][1] = fr_b;
    }
    // ---- retire row R - (2T+2): nothing reads it any more ----------------------------------
    if constexpr (ROLE == ROLE_HEAD) {
        // to TAIL through the hand-over ring: every row that was loaded, in order.  Row q is the
        // n-th use of slot (q - rs) % D with n = it * (NW / D) + floor((U - lag) / D); the slot is
        // free once TAIL has read use n - 1
        constexpr int lag = 2 * T + 2;
        const int s = (U + 2 * NW - lag) % NW;
        const int q = R - lag;
        const int d = U - lag;                                   // in [-lag, NW - 1 - lag]
        const int hs = ((d % D) + D) % D;
        const int nrel = (d - hs) / D;                           // floor(d / D)
        if (STEADY || (q >= c.rs && q < c.hend)) {
            if (STEADY || it * (NW / D) + nrel >= 1)
                mbar_wait(c.hempty + hs, (uint32_t)((nrel + 2 * (NW / D) - 1) & 1));
            double *dst = c.hring + hs * SW;
            *reinterpret_cast<double2 *>(dst) = make_double2(W[s][0], W[s][1]);
            *reinterpret_cast<double2 *>(dst + 64) = make_double2(W[s][2], W[s][3]);
            __syncwarp();
            if (c.lane == 0) mbar_arrive(c.hfull + hs);
        }
    } else {
        const int lag = 2 * T + 2;
        const int s = (U + 2 * NW - lag) % NW;
        const int q = R - lag;
        if (STEADY || (unsigned)(q - c.x0) < (unsigned)(c.x1 - c.x0)) {
            double *dst = c.pout + (roff - lag * c.pitch);
            if (c.cmA) stg_f64x2(dst, W[s][0], W[s][1]);
            if (c.cmB) stg_f64x2(dst + 64, W[s][2], W[s][3]);
            // halo exchange fused into the pass: the rows within H of a slab edge also go
            // straight into the neighbour's halo rows (the steady range keeps clear of them)
            if (!STEADY) {
                const bool to_lo = pe.lo_p[0] != nullptr && q < pe.own0 + pe.H;
                const bool to_hi = pe.hi_p[0] != nullptr && q >= pe.own1 - pe.H;
                if (to_lo || to_hi) {  // warp-uniform, rare: 2H rows per slab and strip
                    const int dstbuf = pe.ctl->src ^ 1;
                    const int64_t rowshift = to_lo ? pe.lo_row0 - pe.own0
                                                   : pe.hi_row0 - (pe.own1 - pe.H);
                    double *peer = (to_lo ? pe.lo_p[dstbuf] : pe.hi_p[dstbuf]) +
                                   (dst - pe.pbuf[dstbuf]) + rowshift * c.pitch;
                    if (c.cmA) stg_f64x2(peer, W[s][0], W[s][1]);
                    if (c.cmB) stg_f64x2(peer + 64, W[s][2], W[s][3]);
                }
            }
        }
    }
}


// One tick of a chain warp: TWO rows.  Rows R (register slot U, even) and R + 1 arrive; for
// k = 0..T-1 the red half-sweep of sweep k runs on rows R - 2k and R - 2k - 1, then the black
// one on rows R - 2k - 1 and R - 2k - 2 (the single-row ticks R and R + 1 merged: the two
// rows of a half-sweep do not depend on each other).  TAIL then takes the late red residuals of
// the last sweep on rows R - 2T, R - 2T - 1 and stores rows R - 2T - 2, R - 2T - 1; HEAD
// hands the pair R - 2T, R - 2T + 1 to TAIL.  Same window of 2T + 4 rows as the single-row tick.
template <int T, bool STEADY, int WALL, int ROLE, int KB>
__device__ __forceinline__ void stream_tick2(double (&W)[2 * T + 4][4], double (&C)[T][2],
                                             double (&FR)[2][2], double (&accA)[T + 1],
                                             double (&accB)[T + 1], const int U, const int R,
                                             const int64_t roff, const uint32_t ph, const int it,
                                             const SCtx &c, const StreamPeers &pe) {
    constexpr int NW = stream_nw(T), PF = CHAIN_PF;
    constexpr int AB = KB > 0 ? KB - 1 : 0;
    constexpr int D = CHAIN_D;
    static_assert(NW == CHAIN_NP, "p ring slot = rhs slot");
    const RbConsts &k = c.k;
    // use n of hand-over slot h holds rows rs + n D + h (+1); for the pair in register slot U'
    // of this loop iteration n = it (NW / D) + floor(U' / D): parity needs `ph` iff NW / D is odd
    const uint32_t itpar = (NW / D) % 2 ? ph : 0u;
    auto load2 = [&](const double *src, bool v0, bool v1) {
        if (v0) {
            const double2 a = *reinterpret_cast<const double2 *>(src);
            const double2 b = *reinterpret_cast<const double2 *>(src + 64);
            W[U][0] = a.x; W[U][1] = a.y; W[U][2] = b.x; W[U][3] = b.y;
        } else {
            W[U][0] = W[U][1] = W[U][2] = W[U][3] = 0.0;
        }
        if (v1) {
            const double2 a = *reinterpret_cast<const double2 *>(src + SW);
            const double2 b = *reinterpret_cast<const double2 *>(src + SW + 64);
            W[U + 1][0] = a.x; W[U + 1][1] = a.y; W[U + 1][2] = b.x; W[U + 1][3] = b.y;
        } else {
            W[U + 1][0] = W[U + 1][1] = W[U + 1][2] = W[U + 1][3] = 0.0;
        }
    };
    // ---- rows R, R + 1: requested (HEAD: rows R + PF, R + PF + 1 go out) or handed over ------
    __syncwarp();  // every lane is done with the slots the request overwrites
    if constexpr (ROLE != ROLE_TAIL) {
        double *bank = U + PF < NW ? c.rr_cur : c.rr_next;
        if (STEADY) {
            if (elect_one()) issue_rows2(c, bank, roff + PF * c.pitch, (U + PF) % NW, true, true);
        } else if (c.lane == 0) {
            const bool v0 = R + PF < c.re, v1 = R + PF + 1 < c.re;
            if (v0 || v1) issue_rows2(c, bank, roff + PF * c.pitch, (U + PF) % NW, v0, v1);
        }
        const bool l0 = STEADY || (R >= c.first && R < c.re);
        const bool l1 = STEADY || (R + 1 >= c.first && R + 1 < c.re);
        if (l0 || l1) mbar_wait(c.bar + U, ph);
        load2(c.pl + U * SW, l0, l1);
    } else {
        if (STEADY || R < c.hend) {
            mbar_wait(c.hfull + U % D, (itpar ^ (uint32_t)(U / D)) & 1u);
            load2(c.hring + (U % D) * SW, true, true);
            __syncwarp();
            if (c.lane == 0) mbar_arrive(c.hempty + U % D);
        } else {
            load2(c.hring, false, false);
        }
    }
    double fre_a = 0.0, fre_b = 0.0, fro_a = 0.0, fro_b = 0.0;  // red rhs of the last red half-sweep
#pragma unroll
    for (int kk = 0; kk < T; kk++) {
        const int g = KB + kk;                      // sweep index in the pass
        const int lv = (g > 0 ? g - 1 : 0) - AB;    // accumulator of the late red residuals
        const int se = (U + 2 * NW - 2 * kk) % NW;          // even row R - 2kk
        const int so = (U + 2 * NW - 2 * kk - 1) % NW;      // odd row R - 2kk - 1
        const int se2 = (U + 2 * NW - 2 * kk - 2) % NW;     // even row R - 2kk - 2
        const double *rpe = (U >= 2 * kk ? c.rl_cur : c.rl_prev) + se * SW;
        const double *rpo = (U >= 2 * kk + 1 ? c.rl_cur : c.rl_prev) + so * SW;
        const double2 rAe = *reinterpret_cast<const double2 *>(rpe);
        const double2 rBe = *reinterpret_cast<const double2 *>(rpe + 64);
        const double2 rAo = *reinterpret_cast<const double2 *>(rpo);
        const double2 rBo = *reinterpret_cast<const double2 *>(rpo + 64);
        // ---- red half-sweep of sweep kk: first cells of the even row, second cells of the odd
        half_sweep<T, STEADY, WALL, 0, true>(W, se, R - 2 * kk, g, rAe.x, rBe.x, accA[lv], accB[lv], c);
        half_sweep<T, STEADY, WALL, 1, true>(W, so, R - 2 * kk - 1, g, rAo.y, rBo.y, accA[lv], accB[lv], c);
        // ---- black half-sweep: first cells of the odd row (rhs read above), second cells of the
        //      even row below it (rhs carried from the tick before)
        half_sweep<T, STEADY, WALL, 0, false>(W, so, R - 2 * kk - 1, g, rAo.x, rBo.x, accA[g - AB],
                                              accB[g - AB], c);
        half_sweep<T, STEADY, WALL, 1, false>(W, se2, R - 2 * kk - 2, g, C[kk][0], C[kk][1],
                                              accA[g - AB], accB[g - AB], c);
        C[kk][0] = rAe.y;
        C[kk][1] = rBe.y;
        if (kk == T - 1) { fre_a = rAe.x; fre_b = rBe.x; fro_a = rAo.y; fro_b = rBo.y; }
    }
    // ---- residual of the red cells of the LAST sweep of the pass on rows R - 2T (even) and
    //      R - 2T - 1 (odd); their red rhs values were read one tick ago (FR) -----------------
    if constexpr (ROLE != ROLE_HEAD) {
#ifndef SB_STREAM_NOCOMPUTE
        {
            const int s = (U + 2 * NW - 2 * T) % NW, sm = (s + NW - 1) % NW, sp = (s + 1) % NW;
            const int q = R - 2 * T;
            if (STEADY || (unsigned)(q - c.cx0) < (unsigned)(c.cx1 - c.cx0)) {  // warp-uniform
                double na, nb;
                nbr2<0>(c, W[s], na, nb);
                double ra = fma(-k.diag, W[s][0], tsum(k, W[sp][0] + W[sm][0], W[s][1] + na, FR[0][0]));
                double rb = fma(-k.diag, W[s][2], tsum(k, W[sp][2] + W[sm][2], W[s][3] + nb, FR[0][1]));
                if (WALL) ra = c.keepA ? 0.0 : ra;
                accA[KB + T - 1 - AB] = fma(ra, ra, accA[KB + T - 1 - AB]);
                accB[KB + T - 1 - AB] = fma(rb, rb, accB[KB + T - 1 - AB]);
            }
        }
        {
            const int s = (U + 2 * NW - 2 * T - 1) % NW, sm = (s + NW - 1) % NW, sp = (s + 1) % NW;
            const int q = R - 2 * T - 1;
            if (STEADY || (unsigned)(q - c.cx0) < (unsigned)(c.cx1 - c.cx0)) {  // warp-uniform
                double na, nb;
                nbr2<1>(c, W[s], na, nb);
                double ra = fma(-k.diag, W[s][1], tsum(k, W[sp][1] + W[sm][1], W[s][0] + na, FR[1][0]));
                double rb = fma(-k.diag, W[s][3], tsum(k, W[sp][3] + W[sm][3], W[s][2] + nb, FR[1][1]));
                if (WALL) rb = c.keepB ? 0.0 : rb;
                accA[KB + T - 1 - AB] = fma(ra, ra, accA[KB + T - 1 - AB]);
                accB[KB + T - 1 - AB] = fma(rb, rb, accB[KB + T - 1 - AB]);
            }
        }
#endif
        FR[0][0] = fre_a; FR[0][1] = fre_b;
        FR[1][0] = fro_a; FR[1][1] = fro_b;
    }
    // ---- retire ---------------------------------------------------------------------------
    if constexpr (ROLE == ROLE_HEAD) {
        // rows R - 2T, R - 2T + 1 have seen all of HEAD's half-sweeps: to TAIL, every pair from
        // rs on in order.  Pair q is use n = it (NW / D) + floor((U - 2T) / D) of its slot, which
        // is free once TAIL has read use n - 1
        const int s0 = (U + 2 * NW - 2 * T) % NW, s1 = (s0 + 1) % NW;
        const int q = R - 2 * T;
        const int d = U - 2 * T;
        const int hs = ((d % D) + D) % D;
        const int nrel = (d - hs) / D;                           // floor(d / D)
        if (STEADY || (q >= c.rs && q < c.hend)) {
            if (STEADY || it * (NW / D) + nrel >= 1)
                mbar_wait(c.hempty + hs, (itpar + (uint32_t)(nrel + 2 * (NW / D) + 1)) & 1u);
            double *dst = c.hring + hs * SW;
            *reinterpret_cast<double2 *>(dst) = make_double2(W[s0][0], W[s0][1]);
            *reinterpret_cast<double2 *>(dst + 64) = make_double2(W[s0][2], W[s0][3]);
            *reinterpret_cast<double2 *>(dst + SW) = make_double2(W[s1][0], W[s1][1]);
            *reinterpret_cast<double2 *>(dst + SW + 64) = make_double2(W[s1][2], W[s1][3]);
            __syncwarp();
            if (c.lane == 0) mbar_arrive(c.hfull + hs);
        }
    } else {
#pragma unroll
        for (int j = 0; j < 2; j++) {   // rows R - 2T - 2, R - 2T - 1: nothing reads them any more
            const int lag = 2 * T + 2 - j;
            const int s = (U + 2 * NW - lag) % NW;
            const int q = R - lag;
            if (STEADY || (unsigned)(q - c.x0) < (unsigned)(c.x1 - c.x0)) {
                double *dst = c.pout + (roff - lag * c.pitch);
                if (c.cmA) stg_f64x2(dst, W[s][0], W[s][1]);
                if (c.cmB) stg_f64x2(dst + 64, W[s][2], W[s][3]);
                // halo exchange fused into the pass: the rows within H of a slab edge also go
                // straight into the neighbour's halo rows (the steady range keeps clear of them)
                if (!STEADY) {
                    const bool to_lo = pe.lo_p[0] != nullptr && q < pe.own0 + pe.H;
                    const bool to_hi = pe.hi_p[0] != nullptr && q >= pe.own1 - pe.H;
                    if (to_lo || to_hi) {  // warp-uniform, rare: 2H rows per slab and strip
                        const int dstbuf = pe.ctl->src ^ 1;
                        const int64_t rowshift = to_lo ? pe.lo_row0 - pe.own0
                                                       : pe.hi_row0 - (pe.own1 - pe.H);
                        double *peer = (to_lo ? pe.lo_p[dstbuf] : pe.hi_p[dstbuf]) +
                                       (dst - pe.pbuf[dstbuf]) + rowshift * c.pitch;
                        if (c.cmA) stg_f64x2(peer, W[s][0], W[s][1]);
                        if (c.cmB) stg_f64x2(peer + 64, W[s][2], W[s][3]);
                    }
                }
            }
        }
    }
}

// One work item (ROLE_SOLO) or one warp's half of it (chain).  TP = sweeps of the whole pass
// (levels written); `partial` = this warp's slot of level 0.
template <int T, int WALL, int ROLE, int KB, int TP>
__device__ __forceinline__ void stream_item(SCtx &c, int flags, int gpar,
                                            double *__restrict__ partial, int64_t part_stride,
                                            const StreamPeers &pe) {
    constexpr bool ROWS2 = ROLE != ROLE_SOLO;   // chain warps: two rows per tick
    constexpr int NW = stream_nw(T), NP = stream_np(T);
    constexpr int PF = ROWS2 ? CHAIN_PF : stream_pf(T);
    constexpr int HP = 2 * TP + 2;          // warm-up rows of the whole pass
    constexpr int AB = KB > 0 ? KB - 1 : 0;
    constexpr int NBANK = ROLE == ROLE_SOLO ? 1 : CHAIN_BANKS;
    const bool lo = flags & IT_BC_LO, hi = flags & IT_BC_HI;
    c.first = lo ? c.x0 : c.x0 - HP;
    c.re = hi ? c.x1 : c.x1 + HP;
    // the tick loop starts on a row of even global x so that register slot parity = row parity
    const int rs = c.first - ((gpar + c.first) & 1);
    c.rs = rs;
    // TAIL reads rows up to its last tick; HEAD hands over exactly those, the others retire
    // the stored rows
    c.hend = min(c.re, c.x1 + 2 * T + 2);
    c.rend = ROLE == ROLE_HEAD ? c.hend + 2 * T : c.x1 + 2 * T + 2;
    c.bc_lo = lo ? c.x0 : -(1 << 29);
    c.bc_hi = hi ? c.x1 - 1 : (1 << 29);
    c.cx0 = c.x0 + (lo ? 1 : 0);
    c.cx1 = c.x1 - (hi ? 1 : 0);
    c.keepA = (flags & 3) == IT_WALL_LO && c.lane == 0;
    c.keepB = (flags & 3) == IT_WALL_HI && c.lane == 31;
    // steady ticks R in [st_lo, st_hi]: rows R-1 .. R-(2T+2) are ordinary counted rows (not
    // next to a boundary row either), row R + PF is still to be requested, and a steady tick
    // retires row R - (2T+2) without tests: not a row a neighbour slab gets.  In a chain the
    // counted rows of sweep g are the stored rows widened by 2 (TP - 1 - g) on either side,
    // and every hand-over slot must have been used (two loop iterations in).
    const int lo_end = pe.lo_p[0] != nullptr ? pe.own0 + pe.H : INT_MIN;
    const int hi_beg = pe.hi_p[0] != nullptr ? pe.own1 - pe.H : INT_MAX - 64;
    int st_lo = max(c.x0 + (lo ? 2 : 0), lo_end) + 2 * T + 2;
    int st_hi = min(min(c.x1 - (hi ? 2 : 0), c.re - PF - 1), hi_beg + 2 * T + 1);
    if (ROWS2) {
        // a two-row tick R sweeps rows R .. R - 2T - 1, stores R - 2T - 2 and R - 2T - 1 and
        // requests rows up to R + PF + 1; the last tick of a loop iteration is R0 + NW - 2
        st_lo = max(st_lo, rs + 2 * NW);
        st_hi = min(c.x1 - 1 - (hi ? 2 : 0), hi_beg + 2 * T) + 1;
        if (ROLE == ROLE_HEAD) st_hi = min(st_hi, c.re - PF - 2 + 1);
    }
    if (ROLE != ROLE_TAIL && c.lane == 0) {
        for (int i = 0; i < NW; i++) mbar_init(c.bar + i, 1);
        if (ROLE == ROLE_HEAD)
            for (int i = 0; i < CHAIN_D; i++) {
                mbar_init(c.hfull + i, 1);
                mbar_init(c.hempty + i, 1);
            }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // tick R requests row R + PF: rows below rs + PF are requested here (bank 0)
        if (ROWS2) {
            for (int r = rs; r < rs + PF; r += 2) {
                const bool v0 = r >= c.first && r < c.re, v1 = r + 1 >= c.first && r + 1 < c.re;
                if (v0 || v1) issue_rows2(c, c.rring, (int64_t)r * c.pitch, r - rs, v0, v1);
            }
        } else {
            for (int r = c.first; r < rs + PF; r++)
                if (r < c.re) issue_row<NP>(c, c.rring, (int64_t)r * c.pitch, r - rs);
        }
    }
    __syncwarp();
    // TAIL may touch the rings once HEAD has initialised the mbarriers
    if (ROLE == ROLE_HEAD) asm volatile("bar.arrive %0, 64;" ::"r"(c.group_bar) : "memory");
    if (ROLE == ROLE_TAIL) asm volatile("bar.sync %0, 64;" ::"r"(c.group_bar) : "memory");
    double W[NW][4], C[T][2], FR[2][2], accA[T + 1], accB[T + 1];
#pragma unroll
    for (int i = 0; i < NW; i++) W[i][0] = W[i][1] = W[i][2] = W[i][3] = 0.0;
#pragma unroll
    for (int i = 0; i < T; i++) C[i][0] = C[i][1] = 0.0;
#pragma unroll
    for (int i = 0; i <= T; i++) accA[i] = accB[i] = 0.0;
    FR[0][0] = FR[0][1] = FR[1][0] = FR[1][1] = 0.0;
    uint32_t ph = 0;
    int it = 0, bank = 0;
    int64_t roff = (int64_t)rs * c.pitch;
    const double *rl0 = c.rring + 2 * c.lane;
    auto set_banks = [&]() {   // rhs banks of loop iteration `it`: bank = it % NBANK
        const int prev = bank == 0 ? NBANK - 1 : bank - 1, next = bank == NBANK - 1 ? 0 : bank + 1;
        c.rl_cur = rl0 + bank * (NW * SW);
        c.rl_prev = rl0 + prev * (NW * SW);
        c.rr_cur = c.rring + bank * (NW * SW);
        c.rr_next = c.rring + next * (NW * SW);
        bank = next;
    };
    for (int R0 = rs; R0 < c.rend;) {
        set_banks();
        if (R0 >= st_lo && R0 + NW - 1 <= st_hi) {
            do {  // the steady state: straight-line code, no row tests
#pragma unroll
                for (int U = 0; U < NW; U += ROWS2 ? 2 : 1) {
                    if constexpr (ROWS2) {
                        stream_tick2<T, true, WALL, ROLE, KB>(W, C, FR, accA, accB, U, R0 + U, roff, ph,
                                                              it, c, pe);
                        roff += 2 * c.pitch;
                    } else {
                        stream_tick<T, true, WALL, ROLE, KB>(W, C, FR, accA, accB, U, R0 + U, roff, ph,
                                                             it, c, pe);
                        roff += c.pitch;
                    }
                }
                ph ^= 1u;
                it++;
                R0 += NW;
                if (R0 + NW - 1 <= st_hi) set_banks();
                else break;
            } while (true);
        } else {
#pragma unroll
            for (int U = 0; U < NW; U += ROWS2 ? 2 : 1) {
                if (R0 + U >= c.rend) break;
                if constexpr (ROWS2) {
                    stream_tick2<T, false, WALL, ROLE, KB>(W, C, FR, accA, accB, U, R0 + U, roff, ph, it,
                                                           c, pe);
                    roff += 2 * c.pitch;
                } else {
                    stream_tick<T, false, WALL, ROLE, KB>(W, C, FR, accA, accB, U, R0 + U, roff, ph, it,
                                                          c, pe);
                    roff += c.pitch;
                }
            }
            ph ^= 1u;
            it++;
            R0 += NW;
        }
    }
    // this warp's levels: HEAD 0 .. T-1 (the late red residuals of its last sweep are TAIL's),
    // TAIL KB-1 .. KB+T-1, SOLO 0 .. T-1; every other level of the pass gets a zero
#pragma unroll
    for (int g = 0; g < TP; g++) {
        double v = 0.0;
        if (g >= AB && g - AB <= T && (ROLE == ROLE_TAIL || g < T))
            v = warp_sum_down((c.cmA ? accA[g - AB] : 0.0) + (c.cmB ? accB[g - AB] : 0.0));
        if (c.lane == 0) partial[(int64_t)g * part_stride] = v;
    }
}

template <int T, int ROLE, int KB, int TP>
__device__ __forceinline__ void stream_item_any(SCtx &c, int flags, int gpar, double *partial,
                                                int64_t part_stride, const StreamPeers &pe) {
    if ((flags & 3) == IT_PLAIN) stream_item<T, 0, ROLE, KB, TP>(c, flags, gpar, partial, part_stride, pe);
    else stream_item<T, 1, ROLE, KB, TP>(c, flags, gpar, partial, part_stride, pe);
}

// One CTA per SM, all its items of one kind.  TB = the configured temporal block (the lattice
// and the shared-memory budget follow it), the pass itself runs ctl->active_T <= TB sweeps.
// TB <= 3: one warp per item; TB = 4: two (HEAD / TAIL), and a shortened pass (T < 4) is run
// by the first warp of each pair alone.  Items with x1 <= x0 pad a CTA to one kind.
template <int TB>
__global__ void __launch_bounds__(32 * stream_warps(TB), 1)
sor_rb_stream_kernel(const RbItem *__restrict__ items, double *const *__restrict__ pbuf,
                     const double *__restrict__ rhs, SorCtl *ctl, double *partial, int part_base,
                     int part_stride, int64_t pitch, int gpar, RbConsts k, RbFin fin,
                     const __grid_constant__ StreamPeers pe, const __grid_constant__ SlabLink lk) {
    constexpr bool CH = stream_chained(TB);
    constexpr int IPC = stream_items_per_cta(TB), SPI = stream_slots_per_item(TB);
    const int T = ctl->active_T;
    if (T == 0) return;
    const int src = ctl->src;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // broadcast from lane 0: tells the compiler the warp index (and all that follows from
    // it: ring addresses, item fields) is warp-uniform, so it stays on the uniform datapath
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    // chain: warps 0 .. IPC-1 are the HEADs, IPC .. 2 IPC-1 the TAILs, so that every scheduler
    // (warp index mod 4) carries both roles (HEAD is the slower one: -3 % per pass against
    // HEAD / TAIL on alternating warps, profiles/r2_stream_chain_ab.txt)
    const int slot = CH ? warp % IPC : warp;         // item of this CTA
    const int role = CH ? warp / IPC : 0;            // chain: 0 HEAD, 1 TAIL
    const int idx = blockIdx.x * IPC + slot;
    const RbItem it = items[idx];
    const int flags = it.pad & 0xff, st0 = (it.pad >> 8) & 0xff, st1 = (it.pad >> 16) & 0xff;
    SCtx c;
    c.lane = threadIdx.x & 31;
    double *part = partial + part_base + idx * SPI + role;
    if (it.x1 <= it.x0) {  // padding: contributes nothing
        if (c.lane == 0)
            for (int i = 0; i < T; i++) part[(int64_t)i * part_stride] = 0.0;
    } else {
        c.lane_m1 = (c.lane + 31) & 31;
        c.lane_p1 = (c.lane + 1) & 31;
        unsigned char *base = smem_raw + slot * stream_smem(TB);
        const bool chain_now = CH && T == TB;
        // rings of the T actually run: p ring, rhs ring (x 3 banks in a chain), hand-over ring,
        // then the mbarriers
        const int tw = chain_now ? CHAIN_TW : T;
        c.pring = reinterpret_cast<double *>(base);
        c.rring = c.pring + (chain_now ? CHAIN_NP : stream_np(tw)) * SW;
        c.hring = c.rring + (chain_now ? CHAIN_BANKS : 1) * stream_nw(tw) * SW;
        c.bar = reinterpret_cast<uint64_t *>(c.hring + (chain_now ? CHAIN_D : 0) * SW);
        c.hfull = c.bar + stream_nw(tw);
        c.hempty = c.hfull + CHAIN_D;
        c.hring += 2 * c.lane;
        c.pl = c.pring + 2 * c.lane;
        c.pin = pbuf[src] + it.ty0;
        c.pout = pbuf[src ^ 1] + it.ty0 + 2 * c.lane;
        c.rhs = rhs + it.ty0;
        c.pitch = pitch;
        c.x0 = it.x0;
        c.x1 = it.x1;
        c.k = k;
        c.group_bar = 1 + slot;
        // stored = counted columns [st0, st1) of the strip (pair-aligned; a wall cell inside is
        // stored but not counted: keepA / keepB)
        c.cmA = 2 * c.lane >= st0 && 2 * c.lane < st1;
        c.cmB = 64 + 2 * c.lane >= st0 && 64 + 2 * c.lane < st1;
        if (CH) {
            if (T == TB) {
                if (role == 0) {
                    stream_item_any<CHAIN_TW, ROLE_HEAD, 0, TB>(c, flags, gpar, part, part_stride, pe);
                } else {
                    stream_item_any<CHAIN_TW, ROLE_TAIL, CHAIN_TW, TB>(c, flags, gpar, part, part_stride, pe);
                }
            } else if (role == 0) {
                if (T == 1) stream_item_any<1, ROLE_SOLO, 0, 1>(c, flags, gpar, part, part_stride, pe);
                else if (T == 2) stream_item_any<2, ROLE_SOLO, 0, 2>(c, flags, gpar, part, part_stride, pe);
                else stream_item_any<3, ROLE_SOLO, 0, 3>(c, flags, gpar, part, part_stride, pe);
            } else if (c.lane == 0) {
                for (int i = 0; i < T; i++) part[(int64_t)i * part_stride] = 0.0;
            }
        } else {
            if (T == TB) stream_item_any<TB, ROLE_SOLO, 0, TB>(c, flags, gpar, part, part_stride, pe);
            else if (TB > 1 && T == 1) stream_item_any<1, ROLE_SOLO, 0, 1>(c, flags, gpar, part, part_stride, pe);
            else if (TB > 2 && T == 2) stream_item_any<2, ROLE_SOLO, 0, 2>(c, flags, gpar, part, part_stride, pe);
        }
    }
    // ---- the last CTA to finish totals the partials of the whole pass (tile kernel's
    //      included: it ran before), applies the exit rule and advances the control block --
    //      what sor_finalize_kernel does as a separate launch.  Row slabs: the per-slab sums of
    //      all ranks are gathered here too (slab_dev.cuh) and added in rank order, so every
    //      rank takes the same exit decision; the round is also the release / acquire point of
    //      the halo rows this kernel stored into the neighbours (system-scope fence below) ------
    if (!fin.enabled) return;
    __shared__ int s_last;
    __shared__ double s_sum[stream_warps(TB)];
    __shared__ double s_norm[RB_TMAX];
    __shared__ double s_gathered[SB_MAX_WORLD * 8];
    if (lk.world > 1) __threadfence_system();
    else __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(fin.counter, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int lvl = 0; lvl < T; lvl++) {
        double acc = 0.0;
        for (int i = threadIdx.x; i < part_stride; i += blockDim.x)
            acc += __ldcg(partial + (int64_t)lvl * part_stride + i);
        acc = warp_sum_down(acc);
        if (c.lane == 0) s_sum[warp] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int w = 0; w < stream_warps(TB); w++) t += s_sum[w];
            s_norm[lvl] = t;
        }
        __syncthreads();
    }
    if (lk.world > 1) {
        const bool ok = slab_allgather(lk, s_norm, T, s_gathered);
        if (!ok) {   // a peer went missing: end the solve, the host reports it
            if (threadIdx.x == 0) {
                ctl->active_T = 0;
                ctl->finished = 1;
                *fin.counter = 0u;
            }
            return;
        }
        if ((int)threadIdx.x < T) {
            double t = s_gathered[threadIdx.x];
            for (int r = 1; r < lk.world; r++) t += s_gathered[r * 8 + threadIdx.x];
            s_norm[threadIdx.x] = t;
        }
        __syncthreads();
    }
    if ((int)threadIdx.x < T) s_norm[threadIdx.x] = s_norm[threadIdx.x] / fin.fluid_cells;
    __syncthreads();
    if (threadIdx.x == 0) {
        sor_advance_ctl(ctl, s_norm, T, fin.initial_norm, fin.eps2, fin.test_exit, fin.norm_hist);
        *fin.counter = 0u;
    }
}

using StreamKernel = void (*)(const RbItem *, double *const *, const double *, SorCtl *,
                              double *, int, int, int64_t, int, RbConsts, RbFin, StreamPeers,
                              SlabLink);
StreamKernel stream_kernel(int TB) {
    switch (TB) {
    case 1: return sor_rb_stream_kernel<1>;
    case 2: return sor_rb_stream_kernel<2>;
    case 3: return sor_rb_stream_kernel<3>;
    default: return sor_rb_stream_kernel<4>;
    }
}
int stream_cta_warps(int TB) {
    switch (TB) {
    case 1: return stream_warps(1);
    case 2: return stream_warps(2);
    case 3: return stream_warps(3);
    default: return stream_warps(4);
    }
}
// work items of one CTA
int stream_cta_items(int TB) {
    switch (TB) {
    case 1: return stream_items_per_cta(1);
    case 2: return stream_items_per_cta(2);
    case 3: return stream_items_per_cta(3);
    default: return stream_items_per_cta(4);
    }
}
// dynamic shared memory of one CTA
int stream_smem_bytes(int TB) {
    switch (TB) {
    case 1: return stream_smem(1) * stream_items_per_cta(1);
    case 2: return stream_smem(2) * stream_items_per_cta(2);
    case 3: return stream_smem(3) * stream_items_per_cta(3);
    default: return stream_smem(4) * stream_items_per_cta(4);
    }
}

// Class of every lattice tile (BX x BY inner region, h cells of footprint around it), decided
// on the cells of the 128-column strip an item of that tile would run on:
//   tiles of the first / last tile column: strip [0, 128) / [NY-128, NY) -- the wall column
//   must hold boundary cells fed from their y-neighbour (edge class S / N) and nothing else
//   may be a boundary cell; other tiles: strip [tj BY - h, ..+128), all fluid;
//   grid rows 0 and NX-1, if the tile owns them, must hold boundary cells fed from the next /
//   previous row (edge class E / W; the cells on a wall column: no edge class).
// Only the kinds matter: what lies outside the footprint cannot reach the owned cells within
// T sweeps.  0 = none of this holds: tile kernel.
__global__ void rb_class_kernel(const uint8_t *__restrict__ cflag, Geom g, int tiles_y, int BX,
                                int BY, int h, uint8_t *__restrict__ cls) {
    const int tile = blockIdx.x;
    const int ti = tile / tiles_y, tj = tile - ti * tiles_y;
    const int64_t x0 = g.own0 + (int64_t)ti * BX;
    const int64_t x1 = min(x0 + BX, g.own1);
    int kind = IT_PLAIN;
    int64_t c0 = (int64_t)tj * BY - h;
    if (tj == 0) { kind = IT_WALL_LO; c0 = 0; }
    else if (tj == tiles_y - 1) { kind = IT_WALL_HI; c0 = g.NY - SW; }
    const int64_t col = c0 + threadIdx.x;
    const bool wall_col = (kind == IT_WALL_LO && col == 0) || (kind == IT_WALL_HI && col == g.NY - 1);
    int ok = tiles_y >= 2 && c0 >= 0 && (c0 & 1) == 0 && col < g.NY;
    // the owned columns of the last tile column need h columns of strip to their left
    if (kind == IT_WALL_HI && (int64_t)tj * BY - c0 < h) ok = 0;
    if (kind == IT_PLAIN && (col < 1 || col > g.NY - 2)) ok = 0;
    int bc_lo = 0, bc_hi = 0, need_next = 0;
    if (ok) {
        for (int64_t r = x0 - h; r < x1 + h; r++) {
            const int64_t gx = g.gx0 + r;
            if (gx < 0 || gx >= g.NX) continue;      // beyond the grid: nothing there
            if (r < 0 || r >= g.nxl) { ok = 0; break; }  // beyond this slab's rows
            const uint8_t f = cflag[r * g.pitch + col];
            if (gx == 0 || gx == g.NX - 1) {
                // a boundary row: must be owned by this tile and be fed from the row inside
                const int want = gx == 0 ? SB_EDGE_E : SB_EDGE_W;
                if (!cf_is_boundary(f) || cf_edge(f) != (wall_col ? SB_EDGE_NONE : want)) { ok = 0; break; }
                if (r >= x0 && r < x1) { if (gx == 0) bc_lo = 1; else bc_hi = 1; }
                else if (gx == g.NX - 1 && r >= x1) need_next = 1;   // owned by the next tile
                else { ok = 0; break; }
            } else if (wall_col) {
                if (!cf_is_boundary(f) ||
                    cf_edge(f) != (kind == IT_WALL_LO ? SB_EDGE_S : SB_EDGE_N)) { ok = 0; break; }
            } else if (!cf_is_fluid(f)) { ok = 0; break; }
        }
    }
    ok = __syncthreads_and(ok);
    // frozen: the inner region and one cell around it lie inside the grid (and this slab's
    // rows) and hold nothing but boundary cells without an edge class
    int frozen = 1;
    {
        const int64_t fc = (int64_t)tj * BY - 1 + threadIdx.x;   // columns tj BY - 1 .. + BY
        const int64_t fc1 = min((int64_t)tj * BY + BY, g.NY) + 1;
        if (fc < fc1) {
            if (fc < 0 || fc >= g.NY) frozen = 0;
            for (int64_t r = x0 - 1; frozen && r < x1 + 1; r++) {
                const int64_t gx = g.gx0 + r;
                if (gx < 0 || gx >= g.NX || r < 0 || r >= g.nxl) { frozen = 0; break; }
                const uint8_t f = cflag[r * g.pitch + fc];
                if (!cf_is_boundary(f) || cf_edge(f) != SB_EDGE_NONE) frozen = 0;
            }
        }
    }
    frozen = __syncthreads_and(frozen);
    if (threadIdx.x == 0)
        cls[tile] = ok ? (uint8_t)(1 + kind + (bc_lo ? TC_BC_LO : 0) + (bc_hi ? TC_BC_HI : 0) +
                                   (need_next ? TC_NEED_NEXT : 0))
                       : (frozen ? (uint8_t)TC_FROZEN : (uint8_t)0);
}

// p[other] := p[current] on the inner regions of the frozen tiles (once per solve)
__global__ void frozen_mirror_kernel(const int32_t *__restrict__ tiles, Geom g, int tiles_y, int BX,
                                     int BY, double *const *__restrict__ pbuf,
                                     const SorCtl *__restrict__ ctl) {
    const int tile = tiles[blockIdx.x];
    const int ti = tile / tiles_y, tj = tile - ti * tiles_y;
    const int64_t x0 = g.own0 + (int64_t)ti * BX, x1 = min(x0 + BX, g.own1);
    const int64_t y0 = (int64_t)tj * BY, y1 = min(y0 + BY, g.NY);
    const double *src = pbuf[ctl->src];
    double *dst = pbuf[ctl->src ^ 1];
    for (int64_t r = x0; r < x1; r++)
        for (int64_t y = y0 + threadIdx.x; y < y1; y += blockDim.x)
            dst[r * g.pitch + y] = src[r * g.pitch + y];
}

// the frozen tiles' residual sum (fixed order) -> the extra partial slot of every level
__global__ void frozen_fill_kernel(const double *__restrict__ part, int n, double *__restrict__ partial,
                                   int part_stride, int slot, int levels) {
    __shared__ double sh[32];
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += part[i];
    acc = warp_sum_down(acc);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < (int)(blockDim.x >> 5); k++) t += sh[k];
        for (int l = 0; l < levels; l++) partial[(int64_t)l * part_stride + slot] = t;
    }
}

}  // namespace

int rb_stream_slots_per_item(int T) { return T >= 4 ? stream_slots_per_item(4) : 1; }

void rb_plan_release(sb_sim *s) {
    cudaFree(s->plan.d_slow);
    cudaFree(s->plan.d_items);
    cudaFree(s->plan.d_plain);
    cudaFree(s->plan.d_counter);
    cudaFree(s->plan.d_frozen);
    cudaFree(s->plan.d_frozen_part);
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
        pl.d_slow = nullptr; pl.d_plain = nullptr;
        pl.cap_tiles = 0;
        SB_CUDA(cudaMallocAsync(&pl.d_slow, (size_t)ntiles * sizeof(int32_t), s->stream));
        SB_CUDA(cudaMallocAsync(&pl.d_plain, (size_t)ntiles, s->stream));
        pl.cap_tiles = (size_t)ntiles;
    }
    rb_class_kernel<<<ntiles, SW, 0, s->stream>>>(s->cflag, g, tiles_y, BX, BY, h, pl.d_plain);
    s->launches++;
    SB_CUDA(cudaGetLastError());
    std::vector<uint8_t> cls((size_t)ntiles);
    SB_CUDA(cudaMemcpyAsync(cls.data(), pl.d_plain, (size_t)ntiles, cudaMemcpyDeviceToHost,
                            s->stream));
    SB_CUDA(cudaStreamSynchronize(s->stream));
    {
        // Which kinds the streaming kernel takes: bit 0 wall strips, bit 1 tiles with boundary
        // rows.  The wall copy of the window costs 0.81 us per row-tick against 0.51 for the
        // plain one at T = 4 (per-item trace, profiles/r1_stream_kinds_ab.txt): its items get
        // half the rows (wall_weight) and fill ~8 SMs, which on a wide grid only breaks even
        // with leaving the wall tiles to the tile kernel; up to T = 3 it wins (0.343 vs
        // 0.394 ms per pass at 8192^2).  SB_RB_STREAM_KINDS overrides (A/B runs).
        const int keep = s->dbg.rb_stream_kinds;   // 3: everything (tests narrow it)
        for (auto &c : cls) {
            if (!(keep & 1) && (c & 3) != 1 + IT_PLAIN) c = 0;
            if (!(keep & 2) && (c & (TC_BC_LO | TC_BC_HI))) c = 0;
        }
    }
    // a tile that needs its successor in the same run (TC_NEED_NEXT) and does not get it
    // stays with the tile kernel
    for (int tj = 0; tj < tiles_y; tj++)
        for (int ti = tiles_x - 1; ti >= 0; ti--) {
            uint8_t &c = cls[(size_t)ti * tiles_y + tj];
            if (c == TC_FROZEN || !(c & TC_NEED_NEXT)) continue;
            const uint8_t nxt = ti + 1 < tiles_x ? cls[(size_t)(ti + 1) * tiles_y + tj] : (uint8_t)0;
            if (nxt == TC_FROZEN || (nxt & 3) != (c & 3)) c = 0;
        }
    // runs of tiles of one item kind along x, per strip.  weight: how much longer a wall
    // strip takes per row than a plain one; its items get that much fewer rows
    struct Run { int tj, ti0, len, kind; double weight; };
    // Wall items run ~1.5x slower per row than plain ones (their copy of the window carries
    // the boundary refresh); weight 2 balances them on one GPU and in row slabs alike.  (With
    // the one-warp T = 4 kernel of round 1 a slab needed weight 6 -- wall items were 2.5x slower
    // there; with the two-warp chain weight 6 costs 10 % at two slabs:
    // profiles/r2_wall_weight_slabs.txt.)
    double wall_weight = 2.0;
    if (s->dbg.wall_weight > 0.0) wall_weight = s->dbg.wall_weight;
    std::vector<Run> runs;
    std::vector<int32_t> slow;
    for (int tj = 0; tj < tiles_y; tj++) {
        int ti = 0;
        while (ti < tiles_x) {
            const int k = cls[(size_t)ti * tiles_y + tj] & 3;
            if (!k) { ti++; continue; }
            int t0 = ti;
            while (ti < tiles_x && (cls[(size_t)ti * tiles_y + tj] & 3) == k) ti++;
            runs.push_back({tj, t0, ti - t0, k - 1, k - 1 == IT_PLAIN ? 1.0 : wall_weight});
        }
    }
    std::vector<int32_t> frozen;
    {
        const bool use_frozen = s->dbg.rb_frozen;   // tests: frozen tiles on the tile kernel
        for (int t = 0; t < ntiles; t++) {
            if (cls[(size_t)t] != TC_FROZEN) continue;
            bool keep_frozen = use_frozen;
            if (s->slab) {  // next to a slab edge the tile kernel also feeds the neighbour's halo
                const int ti = t / tiles_y;
                const int64_t x0 = g.own0 + (int64_t)ti * BX, x1 = std::min<int64_t>(x0 + BX, g.own1);
                if ((s->link.lo_p[0] && x0 < g.own0 + s->link.H) ||
                    (s->link.hi_p[0] && x1 > g.own1 - s->link.H))
                    keep_frozen = false;
            }
            if (keep_frozen) frozen.push_back(t);
            else cls[(size_t)t] = 0;
        }
    }
    for (int t = 0; t < ntiles; t++)
        if (!cls[(size_t)t]) slow.push_back(t);
    // Tiles per item: as many items as fill the SMs in whole waves of CTAs, as long as
    // possible otherwise (every open end of an item pays 2T+2 warm-up rows).  A CTA holds
    // `nwarp` items of ONE kind (at T = 4 two warps run an item).
    const int nwarp = stream_cta_items(T);
    {
        // per device: the attributes live with the function in each context
        static bool attrs_set[64] = {false};
        const int dev = s->device & 63;
        if (!attrs_set[dev]) {
            for (int tb = 1; tb <= RB_TMAX; tb++) {
                cudaFuncSetAttribute(stream_kernel(tb),
                                     cudaFuncAttributePreferredSharedMemoryCarveout,
                                     cudaSharedmemCarveoutMaxShared);
                cudaFuncSetAttribute(stream_kernel(tb), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     stream_smem_bytes(tb));
            }
            attrs_set[dev] = true;
        }
    }
    int dev_sms = 148, ctas = 1;
    cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, s->device);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, stream_kernel(T), 32 * stream_cta_warps(T),
                                                  stream_smem_bytes(T));
    if (ctas < 1) ctas = 1;
    const int64_t resident = (int64_t)dev_sms * ctas;  // CTAs at a time
    // what an item costs beyond its rows and warm-up rows, in rows: pipeline fill and drain on
    // the row-tested copy of the tick, ring set-up, and the CTA's warps waiting for its slowest
    // item before the next wave's CTA can start (profiles/r2_plan_overhead.txt)
    const double item_overhead = s->dbg.plan_overhead >= 0.0 ? s->dbg.plan_overhead : 100.0;
    auto pieces = [](const Run &r, int seg) {
        return std::max(1, std::min(r.len, (int)((r.len * r.weight + seg - 1) / seg)));
    };
    int best_seg = 1;
    double best_cost = 1e300;
    for (int seg = 1; seg <= 256; seg++) {
        int64_t n[2] = {0, 0};
        double longest = 0;
        for (const Run &r : runs) {
            const int m = pieces(r, seg);
            n[r.kind != IT_PLAIN] += m;
            longest = std::max(longest, ((r.len + m - 1) / m) * r.weight);
        }
        const int64_t nc = (n[0] + nwarp - 1) / nwarp + (n[1] + nwarp - 1) / nwarp;
        if (nc == 0) break;
        const int64_t waves = (nc + resident - 1) / resident;
        const double cost = (double)waves * (longest * BX + 2.0 * h + item_overhead);
        if (cost < best_cost) { best_cost = cost; best_seg = seg; }
    }
    // Row-granular plan for ONE wave: split the resident CTAs between the two kinds, hand the
    // item slots of a kind to its runs one by one (always to the run whose pieces are longest)
    // and cut every run into equal pieces at arbitrary (even) rows.  The tile-granular search
    // above can only cut at tile rows and wastes up to a tile per item: with the two wall
    // strips of an 8192^2 channel in the stream it ended 12 % above the ideal and only broke
    // even with leaving them to the tile kernel; this one ends 5 % above (the walls' share).
    auto run_rows = [&](const Run &r) {
        const int64_t a = g.own0 + (int64_t)r.ti0 * BX;
        return (int)(std::min<int64_t>(a + (int64_t)r.len * BX, g.own1) - a);
    };
    std::vector<int> row_pieces;   // pieces per run; empty = keep the tile-granular plan
    {
        const bool enabled = true;
        int nrun[2] = {0, 0};
        for (const Run &r : runs) nrun[r.kind != IT_PLAIN]++;
        double best = best_cost;
        // waves: grids with many strips (32768 columns: 304 runs for 1184 slots) balance better
        // with two to four rounds of shorter items than with one round of 3 or 4 per strip
        for (int waves = 1; enabled && waves <= 6; waves++)
        for (int64_t cw = 0; cw <= resident; cw++) {
            const int64_t c_k[2] = {resident - cw, cw};
            if ((nrun[0] > 0) != (c_k[0] > 0) && nrun[0] > 0) continue;
            if ((nrun[1] > 0) != (c_k[1] > 0)) continue;
            if (c_k[0] * nwarp * waves < nrun[0] || c_k[1] * nwarp * waves < nrun[1]) continue;
            std::vector<int> m(runs.size(), 1);
            double cost = 0.0;
            for (int k = 0; k < 2; k++) {
                int64_t spare = c_k[k] * nwarp * waves - nrun[k];
                // (piece length, run) max-heap
                std::vector<std::pair<double, int>> heap;
                for (size_t i = 0; i < runs.size(); i++)
                    if ((runs[i].kind != IT_PLAIN) == (k == 1))
                        heap.push_back({(double)run_rows(runs[i]), (int)i});
                std::make_heap(heap.begin(), heap.end());
                while (spare > 0 && !heap.empty()) {
                    std::pop_heap(heap.begin(), heap.end());
                    auto top = heap.back();
                    heap.pop_back();
                    const int i = top.second, rows = run_rows(runs[i]);
                    if (rows / (m[i] + 1) < BX) {   // pieces shorter than a tile: not worth it
                        heap.push_back({0.0, i});
                        std::push_heap(heap.begin(), heap.end());
                        if (top.first == 0.0) break;
                        continue;
                    }
                    m[i]++;
                    spare--;
                    heap.push_back({(double)rows / m[i], i});
                    std::push_heap(heap.begin(), heap.end());
                }
                for (size_t i = 0; i < runs.size(); i++)
                    if ((runs[i].kind != IT_PLAIN) == (k == 1))
                        cost = std::max(cost, waves * ((run_rows(runs[i]) + m[i] - 1) / m[i] + 2.0 * h +
                                                       item_overhead) *
                                                  runs[i].weight);
            }
            cost *= 1.0 + 0.03 * (waves - 1);   // ties go to fewer, longer items
            if (cost < best) {
                best = cost;
                row_pieces = m;
            }
        }
    }
    std::vector<RbItem> by_kind[2];  // plain strips / wall strips: one copy of the code each
    for (size_t ri = 0; ri < runs.size(); ri++) {
        const Run &r = runs[ri];
        const bool by_rows = !row_pieces.empty();
        const int m = by_rows ? row_pieces[ri] : pieces(r, best_seg);
        const int64_t rx0 = g.own0 + (int64_t)r.ti0 * BX, rrows = run_rows(r);
        // tile-granular cuts: never between a TC_NEED_NEXT tile and its successor (the cut
        // moves one tile down; the row-granular pieces are >= BX rows long and cannot end
        // within the halo of the boundary row they do not own)
        auto cut = [&](int i) {
            int c = (int)((int64_t)r.len * i / m);
            if (c > 0 && c < r.len && (cls[(size_t)(r.ti0 + c - 1) * tiles_y + r.tj] & TC_NEED_NEXT)) c--;
            return c;
        };
        for (int i = 0; i < m; i++) {
            const int a = cut(i), b = cut(i + 1);
            if (!by_rows && a >= b) continue;
            RbItem it;
            if (by_rows) {
                const int64_t ra = i == 0 ? 0 : (rrows * i / m) & ~(int64_t)1;
                const int64_t rb = i == m - 1 ? rrows : (rrows * (i + 1) / m) & ~(int64_t)1;
                it.x0 = (int32_t)(rx0 + ra);
                it.x1 = (int32_t)(rx0 + rb);
            } else {
                it.x0 = (int32_t)(g.own0 + (int64_t)(r.ti0 + a) * BX);
                it.x1 = (int32_t)std::min<int64_t>(g.own0 + (int64_t)(r.ti0 + b) * BX, g.own1);
            }
            int flags = r.kind, st0 = h, st1 = SW - h;
            it.ty0 = r.tj * BY - h;
            if (r.kind == IT_WALL_LO) { it.ty0 = 0; st0 = 0; st1 = BY; }
            if (r.kind == IT_WALL_HI) {
                it.ty0 = (int32_t)(g.NY - SW);
                st0 = r.tj * BY - it.ty0;
                st1 = SW;
            }
            const int ta = by_rows ? (i == 0 ? 0 : -1) : a, tb = by_rows ? (i == m - 1 ? r.len : -1) : b;
            if (ta >= 0 && (cls[(size_t)(r.ti0 + ta) * tiles_y + r.tj] & TC_BC_LO)) flags |= IT_BC_LO;
            if (tb >= 1 && (cls[(size_t)(r.ti0 + tb - 1) * tiles_y + r.tj] & TC_BC_HI)) flags |= IT_BC_HI;
            it.pad = flags | (st0 << 8) | (st1 << 16);
            by_kind[r.kind != IT_PLAIN].push_back(it);
        }
    }
    // kind by kind, each padded to whole CTAs; inside a kind neighbouring strips of the same
    // rows run side by side: their shared halo columns are then fetched from HBM once
    std::vector<RbItem> items;
    for (auto &v : by_kind) {
        std::stable_sort(v.begin(), v.end(), [](const RbItem &a, const RbItem &b) {
            return a.x0 != b.x0 ? a.x0 < b.x0 : a.ty0 < b.ty0;
        });
        items.insert(items.end(), v.begin(), v.end());
        while (items.size() % nwarp) items.push_back(RbItem{0, 0, 0, 0});
    }
    if (items.size() > pl.cap_items) {
        if (pl.d_items) SB_CUDA(cudaFreeAsync(pl.d_items, s->stream));
        pl.d_items = nullptr;
        pl.cap_items = 0;
        SB_CUDA(cudaMallocAsync(&pl.d_items, items.size() * sizeof(RbItem), s->stream));
        pl.cap_items = items.size();
    }
    if (frozen.size() > pl.cap_frozen) {
        if (pl.d_frozen) SB_CUDA(cudaFreeAsync(pl.d_frozen, s->stream));
        if (pl.d_frozen_part) SB_CUDA(cudaFreeAsync(pl.d_frozen_part, s->stream));
        pl.d_frozen = nullptr; pl.d_frozen_part = nullptr;
        pl.cap_frozen = 0;
        SB_CUDA(cudaMallocAsync(&pl.d_frozen, frozen.size() * sizeof(int32_t), s->stream));
        SB_CUDA(cudaMallocAsync(&pl.d_frozen_part, frozen.size() * sizeof(double), s->stream));
        pl.cap_frozen = frozen.size();
    }
    if (!frozen.empty())
        SB_CUDA(cudaMemcpyAsync(pl.d_frozen, frozen.data(), frozen.size() * sizeof(int32_t),
                                cudaMemcpyHostToDevice, s->stream));
    pl.n_frozen = (int)frozen.size();
    pl.frozen_seq = 0;
    if (!slow.empty())
        SB_CUDA(cudaMemcpyAsync(pl.d_slow, slow.data(), slow.size() * sizeof(int32_t),
                                cudaMemcpyHostToDevice, s->stream));
    if (!items.empty())
        SB_CUDA(cudaMemcpyAsync(pl.d_items, items.data(), items.size() * sizeof(RbItem),
                                cudaMemcpyHostToDevice, s->stream));
    SB_CUDA(cudaStreamSynchronize(s->stream));  // the vectors go out of scope
    if (s->dbg.trace_plan) {
        int nk[3] = {0, 0, 0}, nbc = 0, longest = 0;
        for (const RbItem &it : items) {
            if (it.x1 <= it.x0) continue;
            nk[it.pad & 3]++;
            nbc += (it.pad & (IT_BC_LO | IT_BC_HI)) != 0;
            longest = std::max(longest, it.x1 - it.x0);
        }
        fprintf(stderr, "[sb plan] T=%d tiles %dx%d slow %zu frozen %zu items %zu (plain %d, wall %d+%d, "
                "with boundary rows %d) longest %d rows, seg %d, resident CTAs %lld\n", T, tiles_x,
                tiles_y, slow.size(), frozen.size(), items.size(), nk[0], nk[1], nk[2], nbc, longest,
                best_seg, (long long)resident);
    }
    pl.n_slow = (int)slow.size();
    pl.n_items = (int)items.size();
    pl.tiles_x = tiles_x;
    pl.tiles_y = tiles_y;
    pl.T = T;
    pl.epoch = s->flag_epoch;
    return SB_OK;
}

sb_status launch_sor_rb_stream(sb_sim *s, int part_base, int part_stride, int h,
                               const RbFin *fin_in) {
    const Geom &g = s->g;
    const int gpar = (int)(((g.gx0 % 2) + 2) % 2);
    const int TB = s->prm.temporal_block;
    const int nw = stream_cta_warps(TB), ipc = stream_cta_items(TB);
    RbFin fin{};
    if (fin_in) {
        fin = *fin_in;
        if (!s->plan.d_counter) {
            SB_CUDA(cudaMallocAsync(&s->plan.d_counter, sizeof(unsigned), s->stream));
            SB_CUDA(cudaMemsetAsync(s->plan.d_counter, 0, sizeof(unsigned), s->stream));
        }
        fin.counter = s->plan.d_counter;
    }
    StreamPeers pe{};
    if (s->slab) {
        pe.lo_p[0] = s->link.lo_p[0]; pe.lo_p[1] = s->link.lo_p[1];
        pe.hi_p[0] = s->link.hi_p[0]; pe.hi_p[1] = s->link.hi_p[1];
        pe.lo_row0 = s->link.lo_row0; pe.hi_row0 = s->link.hi_row0;
        pe.own0 = (int)g.own0; pe.own1 = (int)g.own1; pe.H = s->link.H;
    }
    pe.pbuf = rb_pbuf_ptr(s);
    pe.ctl = s->d_ctl;
    SlabLink lk{};
    lk.world = 1;
    if (s->slab) lk = s->link;
    stream_kernel(TB)<<<s->plan.n_items / ipc, 32 * nw, stream_smem_bytes(TB), s->stream>>>(
        s->plan.d_items, rb_pbuf_ptr(s), s->rhs, s->d_ctl, s->d_partial, part_base, part_stride,
        g.pitch, gpar, rb_consts(s), fin, pe, lk);
    s->launches++;
    SB_CUDA(cudaGetLastError());
    return SB_OK;
}

sb_status launch_frozen_mirror(sb_sim *s, int BX, int BY) {
    const int tiles_y = (int)((s->g.NY + BY - 1) / BY);
    frozen_mirror_kernel<<<s->plan.n_frozen, 128, 0, s->stream>>>(s->plan.d_frozen, s->g, tiles_y, BX,
                                                                  BY, rb_pbuf_ptr(s), s->d_ctl);
    s->launches++;
    SB_CUDA(cudaGetLastError());
    return SB_OK;
}

sb_status launch_frozen_fill(sb_sim *s, int part_stride, int slot) {
    frozen_fill_kernel<<<1, 1024, 0, s->stream>>>(s->plan.d_frozen_part, s->plan.n_frozen, s->d_partial,
                                                  part_stride, slot, RB_TMAX);
    s->launches++;
    SB_CUDA(cudaGetLastError());
    return SB_OK;
}

void preload_sor_rb_stream() {
    cudaFuncAttributes a;
    for (int tb = 1; tb <= RB_TMAX; tb++) cudaFuncGetAttributes(&a, stream_kernel(tb));
    cudaFuncGetAttributes(&a, rb_class_kernel);
    cudaFuncGetAttributes(&a, frozen_mirror_kernel);
    cudaFuncGetAttributes(&a, frozen_fill_kernel);
    cudaGetLastError();
}

}  // namespace sb
