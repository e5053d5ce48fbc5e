// sor_rb.cu -- K4b: performance-mode SOR pass.  Red-black ordering, pressure BC, residual
// norm and up to 4 sweeps fused into ONE pass over HBM per launch (temporal blocking).
//
// Replaces, per launch, T iterations of the loop body of Simulation::solve_sor
// (/root/reference/src/simulation.rs:250-281): copy_pressure_to_boundaries
// (src/grid/mod.rs:343-412), the sweep (:253-274, here as a red half-sweep over
// (x+y) even then a black half-sweep) and calculate_norm_squared (:216-227).
//
// One CTA owns a TXR x TW tile (x rows, y columns; y contiguous) staged into shared
// memory by two TMA box loads (p, rhs; out-of-grid cells arrive as zeros).  The u8 cell
// flags are read with plain 16-bit loads while the TMA is in flight: a TMA box must start
// on a 16-byte boundary, which an odd-sized halo of 1-byte cells cannot honour; cells
// outside the grid get flag 0 = "not a cell".  The tile carries a halo of h = 2T+1 cells: after sweep k
// the values at distance >= 2k from the tile edge are exact, so T sweeps leave the
// inner (TXR-2h) x (TW-2h-2) region exact, together with the residuals of all T
// sweeps.  That region is written to the other pressure buffer (ping-pong), and one
// partial sum of squared residuals per sweep and tile goes to the finalize kernel.
//
// HBM traffic per launch and cell: 8 (p in) + 8 (rhs) + 1 (flag) + 8 (p out) = 25 B for
// T sweeps (halo re-reads are served by L2), i.e. 25/T B per cell-sweep.
//
// Each thread owns two adjacent columns (one 16-byte shared-memory word) of a run of
// rows and walks them with a 3-row register window, so a half-sweep reads each
// pressure value once per thread plus one foreign column neighbour.
//
// Arithmetic (identical in oracle/stroemung_oracle.c, SO_SOR_RED_BLACK):
//     t     = fma(1/dx^2, pE+pW, fma(1/dy^2, pS+pN, -rhs))
//     p_new = fma(mid, t, (1-w)*p)            mid = w / (2/dx^2 + 2/dy^2)
//     r     = fma(-(2/dx^2 + 2/dy^2), p, t)   residual, all interior cells
#include "sb_internal.cuh"

namespace sb {

namespace {

constexpr int TXR = 48;                 // tile rows
constexpr int TW = 128;                 // tile columns
constexpr int NTHR = 256;
constexpr int TCOLS = TW / 2;           // thread columns (2 cells each)
constexpr int TROWS = NTHR / TCOLS;     // thread rows
constexpr int RPT = TXR / TROWS;        // rows per thread
constexpr int TMAX = 4;
constexpr int TILE = TXR * TW;
constexpr size_t SMEM_BYTES = (size_t)TILE * 17 + 64 + (NTHR / 32) * TMAX * sizeof(double);
static_assert(TXR % TROWS == 0, "rows must split evenly");

struct RbConsts {
    double rdx2, rdy2, diag, mid, omw;
};

// ---- mbarrier / TMA wrappers (inline PTX) -------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1,
                                            uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}

__device__ __forceinline__ double2 lds2(const double *sp, int idx) {
    return *reinterpret_cast<const double2 *>(sp + idx);
}

// pressure BC of one boundary cell from its fluid neighbours (src/grid/mod.rs:351-399)
__device__ __forceinline__ void bc_cell(double *sp, int idx, int edge) {
    switch (edge) {
    case SB_EDGE_N: sp[idx] = sp[idx - 1]; break;
    case SB_EDGE_NE: sp[idx] = (sp[idx - 1] + sp[idx + TW]) / 2.0; break;
    case SB_EDGE_E: sp[idx] = sp[idx + TW]; break;
    case SB_EDGE_SE: sp[idx] = (sp[idx + 1] + sp[idx + TW]) / 2.0; break;
    case SB_EDGE_S: sp[idx] = sp[idx + 1]; break;
    case SB_EDGE_SW: sp[idx] = (sp[idx + 1] + sp[idx - TW]) / 2.0; break;
    case SB_EDGE_W: sp[idx] = sp[idx - TW]; break;
    case SB_EDGE_NW: sp[idx] = (sp[idx - 1] + sp[idx - TW]) / 2.0; break;
    default: break;
    }
}

struct TileCtx {
    int r_begin, col0;      // first tile row / first (even) tile column of this thread
    int64_t gx_base, gy0;   // global x of tile row 0, global y of col0
    int64_t NX, NY;
    int64_t own_gx0, own_gx1;  // owned global rows (norm is counted there only)
    int h, hy;
};

__device__ __forceinline__ bool interior(const TileCtx &c, int64_t gx, int64_t gy) {
    return gx >= 1 && gx <= c.NX - 2 && gy >= 1 && gy <= c.NY - 2;
}

// one colour of one sweep over this thread's cells; colour 1 also folds the residuals of
// the black fluid cells it updates into acc (their neighbours are already final)
template <int COLOUR>
__device__ __forceinline__ void half_sweep(double *sp, const double *sr, const uint8_t *sf,
                                           const TileCtx &c, const RbConsts &k, double &acc) {
    int r0 = max(c.r_begin, 1), r1 = min(c.r_begin + RPT, TXR - 1);
    double2 pm = lds2(sp, (r0 - 1) * TW + c.col0);
    double2 pc = lds2(sp, r0 * TW + c.col0);
#pragma unroll 4
    for (int r = r0; r < r1; r++) {
        double2 pn = lds2(sp, (r + 1) * TW + c.col0);
        const int64_t gx = c.gx_base + r;
        const int sel = (int)((gx + c.gy0 + COLOUR) & 1);  // which cell of the pair has this colour
        const int col = c.col0 + sel;
        const int idx = r * TW + col;
        const bool upd = sf[idx] == CF_FLUID && col >= 1 && col <= TW - 2 &&
                         interior(c, gx, c.gy0 + sel);
        if (upd) {
            const double pE = sel ? pn.y : pn.x, pW = sel ? pm.y : pm.x;
            double pN, pS, pold;
            if (sel == 0) { pS = pc.y; pN = sp[idx - 1]; pold = pc.x; }
            else          { pN = pc.x; pS = sp[idx + 1]; pold = pc.y; }
            const double t = fma(k.rdx2, pE + pW, fma(k.rdy2, pS + pN, -sr[idx]));
            const double pnew = fma(k.mid, t, k.omw * pold);
            sp[idx] = pnew;
            if (sel == 0) pc.x = pnew; else pc.y = pnew;
            if (COLOUR == 1) {
                const bool owned = r >= c.h && r < TXR - c.h && col >= c.hy && col < TW - c.hy &&
                                   gx >= c.own_gx0 && gx < c.own_gx1;
                if (owned) {
                    const double rr = fma(-k.diag, pnew, t);
                    acc = fma(rr, rr, acc);
                }
            }
        }
        pm = pc;
        pc = pn;
    }
}

// residuals of the owned interior cells not covered by the black half-sweep:
// every red cell, and black cells that are not fluid (obstacle cells count in the norm,
// src/simulation.rs:216-227 sums over ALL interior cells)
__device__ __forceinline__ void norm_rest(const double *sp, const double *sr, const uint8_t *sf,
                                          const TileCtx &c, const RbConsts &k, double &acc,
                                          bool all_black) {
    int r0 = max(c.r_begin, c.h), r1 = min(c.r_begin + RPT, TXR - c.h);
    if (c.col0 < c.hy || c.col0 >= TW - c.hy || r0 >= r1) return;
    double2 pm = lds2(sp, (r0 - 1) * TW + c.col0);
    double2 pc = lds2(sp, r0 * TW + c.col0);
#pragma unroll 4
    for (int r = r0; r < r1; r++) {
        double2 pn = lds2(sp, (r + 1) * TW + c.col0);
        const int64_t gx = c.gx_base + r;
        if (gx >= c.own_gx0 && gx < c.own_gx1) {
            const int sel = (int)((gx + c.gy0) & 1);  // the red cell of the pair
            {
                const int idx = r * TW + c.col0 + sel;
                if (interior(c, gx, c.gy0 + sel)) {
                    const double pE = sel ? pn.y : pn.x, pW = sel ? pm.y : pm.x;
                    double pN, pS, pp;
                    if (sel == 0) { pS = pc.y; pN = sp[idx - 1]; pp = pc.x; }
                    else          { pN = pc.x; pS = sp[idx + 1]; pp = pc.y; }
                    const double t = fma(k.rdx2, pE + pW, fma(k.rdy2, pS + pN, -sr[idx]));
                    const double rr = fma(-k.diag, pp, t);
                    acc = fma(rr, rr, acc);
                }
            }
            {
                const int bsel = sel ^ 1;  // the black cell: only if it was not swept
                const int idx = r * TW + c.col0 + bsel;
                if ((all_black || sf[idx] != CF_FLUID) && interior(c, gx, c.gy0 + bsel)) {
                    const double pE = bsel ? pn.y : pn.x, pW = bsel ? pm.y : pm.x;
                    double pN, pS, pp;
                    if (bsel == 0) { pS = pc.y; pN = sp[idx - 1]; pp = pc.x; }
                    else           { pN = pc.x; pS = sp[idx + 1]; pp = pc.y; }
                    const double t = fma(k.rdx2, pE + pW, fma(k.rdy2, pS + pN, -sr[idx]));
                    const double rr = fma(-k.diag, pp, t);
                    acc = fma(rr, rr, acc);
                }
            }
        }
        pm = pc;
        pc = pn;
    }
}

__global__ void __launch_bounds__(NTHR, 2)
sor_rb_kernel(const __grid_constant__ CUtensorMap tm_p0, const __grid_constant__ CUtensorMap tm_p1,
              const __grid_constant__ CUtensorMap tm_rhs, const uint8_t *__restrict__ cflag,
              Geom g, double *const *__restrict__ pbuf, const SorCtl *__restrict__ ctl,
              double *__restrict__ partial, int tiles_y, int ntiles, int h, RbConsts k,
              int norm_only) {
    // norm_only: no sweeps, no write-back; partial[tile] = sum of squared residuals of the
    // current field (calculate_norm_squared on its own, src/simulation.rs:216-227)
    const int T = norm_only ? 1 : ctl->active_T;
    if (T == 0) return;
    const int src = ctl->src;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *sp = reinterpret_cast<double *>(smem_raw);
    double *sr = sp + TILE;
    uint8_t *sf = reinterpret_cast<uint8_t *>(sr + TILE);
    uint64_t *bar = reinterpret_cast<uint64_t *>(sf + TILE);
    double *sred = reinterpret_cast<double *>(sf + TILE + 64);

    const int hy = h;  // h is even, so the owned columns start 16-byte aligned
    const int BX = TXR - 2 * h, BY = TW - 2 * hy;
    const int tile_i = blockIdx.x / tiles_y, tile_j = blockIdx.x - tile_i * tiles_y;
    const int tx0 = tile_i * BX - h;    // local row of tile row 0
    const int ty0 = tile_j * BY - hy;   // column of tile column 0

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

    TileCtx c;
    const int tc = threadIdx.x % TCOLS, tr = threadIdx.x / TCOLS;
    c.r_begin = tr * RPT;
    c.col0 = 2 * tc;
    c.gx_base = g.gx0 + tx0;
    c.gy0 = (int64_t)ty0 + c.col0;
    c.NX = g.NX;
    c.NY = g.NY;
    c.own_gx0 = g.gx0 + g.own0;
    c.own_gx1 = g.gx0 + g.own1;
    c.h = h;
    c.hy = hy;
    double acc[TMAX];
#pragma unroll
    for (int i = 0; i < TMAX; i++) acc[i] = 0.0;

    // cell flags of this thread's own cells (nobody else reads them): global -> smem
    for (int r = c.r_begin; r < c.r_begin + RPT; r++) {
        const int64_t lx = (int64_t)tx0 + r;
        uint16_t ff = 0;
        if (lx >= 0 && lx < g.nxl && c.gy0 >= 0 && c.gy0 + 1 < g.pitch)
            ff = *reinterpret_cast<const uint16_t *>(cflag + lx * g.pitch + c.gy0);
        *reinterpret_cast<uint16_t *>(sf + r * TW + c.col0) = ff;
    }

    mbar_wait(bar, 0);

#pragma unroll
    for (int it = 0; it < TMAX; it++) {
        if (it < T && !norm_only) {
            // pressure BC: boundary cells take the (average of the) fluid neighbour(s)
            {
                int r0 = max(c.r_begin, 1), r1 = min(c.r_begin + RPT, TXR - 1);
                for (int r = r0; r < r1; r++) {
                    const int idx = r * TW + c.col0;
                    const uint16_t ff = *reinterpret_cast<const uint16_t *>(sf + idx);
                    if ((ff & 0x7878) == 0) continue;
                    const int e0 = (ff >> 3) & 15, e1 = (ff >> 11) & 15;
                    if (e0 && c.col0 >= 1) bc_cell(sp, idx, e0);
                    if (e1 && c.col0 + 1 <= TW - 2) bc_cell(sp, idx + 1, e1);
                }
            }
            __syncthreads();
            half_sweep<0>(sp, sr, sf, c, k, acc[it]);
            __syncthreads();
            half_sweep<1>(sp, sr, sf, c, k, acc[it]);
            __syncthreads();
            norm_rest(sp, sr, sf, c, k, acc[it], false);
            __syncthreads();
        }
    }
    if (norm_only) norm_rest(sp, sr, sf, c, k, acc[0], true);

    // write the exact inner region to the other buffer
    if (!norm_only) {
        double *pout = pbuf[src ^ 1];
        int r0 = max(c.r_begin, h), r1 = min(c.r_begin + RPT, TXR - h);
        if (c.col0 >= hy && c.col0 < TW - hy && c.gy0 < g.NY) {
            for (int r = r0; r < r1; r++) {
                const int64_t lx = (int64_t)tx0 + r;
                if (lx < 0 || lx >= g.nxl) continue;
                const double2 val = lds2(sp, r * TW + c.col0);
                double *dst = pout + lx * g.pitch + c.gy0;
                if (c.gy0 + 1 < g.NY) *reinterpret_cast<double2 *>(dst) = val;
                else dst[0] = val.x;
            }
        }
    }

    // block-reduce the per-sweep residual sums (fixed tree => deterministic)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int it = 0; it < TMAX; it++) {
        double v = acc[it];
        for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) sred[warp * TMAX + it] = v;
    }
    __syncthreads();
    if (threadIdx.x < TMAX && (int)threadIdx.x < T) {
        double tsum = 0.0;
        for (int w = 0; w < NTHR / 32; w++) tsum += sred[w * TMAX + threadIdx.x];
        partial[(int64_t)threadIdx.x * ntiles + blockIdx.x] = tsum;
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
    SB_CUDA(cudaFuncSetAttribute(sor_rb_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)SMEM_BYTES));
    s->tmaps_ready = true;
    return SB_OK;
}

}  // namespace

int rb_halo_rows(int T) { return 2 * T + 2; }

static double *const *pbuf_ptr(sb_sim *s) {
    return reinterpret_cast<double *const *>(reinterpret_cast<char *>(s->d_ctl) + 256);
}

// one guarded pass: performs ctl->active_T sweeps from pbuf[ctl->src] into the other buffer
// (norm_only: just the residual partial sums of pbuf[ctl->src])
sb_status launch_sor_rb_pass(sb_sim *s, int *ntiles_out, int norm_only) {
    sb_status st = ensure_tmaps(s);
    if (st) return st;
    const Geom &g = s->g;
    int T = s->prm.temporal_block;
    int h = rb_halo_rows(T);
    int BX = TXR - 2 * h, BY = TW - 2 * h;
    int tiles_x = (int)((g.nxl + BX - 1) / BX), tiles_y = (int)((g.NY + BY - 1) / BY);
    int ntiles = tiles_x * tiles_y;
    size_t need = (size_t)ntiles * TMAX + 64;
    if (need > s->partial_cap) {
        if (s->d_partial) cudaFree(s->d_partial);
        s->d_partial = nullptr;
        SB_CUDA(cudaMalloc(&s->d_partial, need * sizeof(double)));
        s->partial_cap = need;
    }
    RbConsts k;
    double dx2 = s->prm.delx * s->prm.delx, dy2 = s->prm.dely * s->prm.dely;
    k.rdx2 = 1.0 / dx2;
    k.rdy2 = 1.0 / dy2;
    k.diag = (2.0 * k.rdx2) + (2.0 * k.rdy2);
    k.mid = s->prm.omega / ((2.0 / dx2) + (2.0 / dy2));
    k.omw = 1.0 - s->prm.omega;
    sor_rb_kernel<<<ntiles, NTHR, SMEM_BYTES, s->stream>>>(s->tm_p[0], s->tm_p[1], s->tm_rhs,
                                                          s->cflag, g, pbuf_ptr(s), s->d_ctl,
                                                          s->d_partial, tiles_y, ntiles, h, k, norm_only);
    s->launches++;
    SB_CUDA(cudaGetLastError());
    *ntiles_out = ntiles;
    return SB_OK;
}

}  // namespace sb
