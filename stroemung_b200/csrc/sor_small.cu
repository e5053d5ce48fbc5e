// sor_small.cu -- K4d: the whole red-black SOR solve of a SMALL grid in ONE launch.
//
// The reference's own default case is a 100 x 20 grid (BASELINE config 1): 2 000 cells, 100
// sweeps per tick.  Pass by pass that is ~300 dependent launches of kernels that each run for
// a microsecond -- launch latency, 1.4 ms per tick.  A grid of up to 12 288 cells fits the
// shared memory of one SM with room to spare, so one CTA keeps p, rhs and the cell flags
// there and runs solve_sor (/root/reference/src/simulation.rs:239-285) start to finish:
// per iteration the pressure BC (src/grid/mod.rs:343-412), the red and the black half-sweep
// (:253-274 in red-black order), the residual norm over all interior cells (:216-227) and the
// exit test (:279), with block barriers in between and the decision taken on the device.
//
// Arithmetic is that of the tile / streaming kernels and of the oracle's red-black
// restatement (sor_rb.cuh), so p is bit-identical whichever kernel ran; the norm is summed in
// another order (covered by the 1e-12 allowance).
#include <stdlib.h>

#include "sor_rb.cuh"

namespace sb {

namespace {

constexpr int SMALL_THREADS = 1024;

// only the neighbours the edge class names are read: a ring cell has no others
__device__ __forceinline__ double small_bc(const double *sp, int c, int ny, int edge) {
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

// per-cell code in shared memory: bits 0-3 edge class of a boundary cell that takes the BC,
// bit 4 interior cell (counts in the norm), bit 5 fluid interior cell (swept), bit 6 colour
__global__ void __launch_bounds__(SMALL_THREADS, 1)
sor_small_kernel(Geom g, double *const *__restrict__ pbuf, const double *__restrict__ rhs,
                 const uint8_t *__restrict__ cflag, SorCtl *ctl, RbConsts k, double fluid_cells,
                 double initial_norm, double eps2, int test_exit, double *norm_hist) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nx = (int)g.NX, ny = (int)g.NY, n = nx * ny;
    double *sp = reinterpret_cast<double *>(smem_raw);
    double *sr = sp + n;
    uint8_t *sc = reinterpret_cast<uint8_t *>(sr + n);
    __shared__ double s_warp[SMALL_THREADS / 32];
    __shared__ int s_stop;
    double *p = pbuf[ctl->src];
    const uint32_t max_it = ctl->max_iterations;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int c = tid; c < n; c += SMALL_THREADS) {
        const int x = c / ny, y = c - x * ny;
        const int64_t gc = (int64_t)x * g.pitch + y;
        sp[c] = p[gc];
        sr[c] = rhs[gc];
        const uint8_t f = cflag[gc];
        const bool interior = x >= 1 && x <= nx - 2 && y >= 1 && y <= ny - 2;
        uint8_t code = 0;
        if (cf_is_boundary(f)) code = (uint8_t)cf_edge(f);
        if (interior) code |= 16;
        if (interior && cf_is_fluid(f)) code |= 32;
        if ((x + y) & 1) code |= 64;
        sc[c] = code;
    }
    __syncthreads();

    uint32_t it = 0;
    double norm = 0.0;
    int cap = 1;
    while (it < max_it) {
        // pressure BC: reads fluid cells, writes boundary cells
        for (int c = tid; c < n; c += SMALL_THREADS) {
            const int edge = sc[c] & 15;
            if (edge) sp[c] = small_bc(sp, c, ny, edge);
        }
        __syncthreads();
#pragma unroll
        for (int colour = 0; colour < 2; colour++) {
            for (int c = tid; c < n; c += SMALL_THREADS) {
                const uint8_t code = sc[c];
                if ((code & 32) && ((code >> 6) & 1) == colour) {
                    const double t = fma(k.rdx2, sp[c + ny] + sp[c - ny],
                                         fma(k.rdy2, sp[c + 1] + sp[c - 1], -sr[c]));
                    sp[c] = fma(k.mid, t, k.omw * sp[c]);
                }
            }
            __syncthreads();
        }
        // residual norm over ALL interior cells, divided by the fluid cell count
        double acc = 0.0;
        for (int c = tid; c < n; c += SMALL_THREADS) {
            if (sc[c] & 16) {
                const double t = fma(k.rdx2, sp[c + ny] + sp[c - ny],
                                     fma(k.rdy2, sp[c + 1] + sp[c - 1], -sr[c]));
                const double r = fma(-k.diag, sp[c], t);
                acc = fma(r, r, acc);
            }
        }
        acc = warp_sum_down(acc);
        if (lane == 0) s_warp[warp] = acc;
        __syncthreads();
        if (tid == 0) {
            double t = 0.0;
            for (int w = 0; w < SMALL_THREADS / 32; w++) t += s_warp[w];
            const double nrm = t / fluid_cells;
            s_warp[0] = nrm;
            if (norm_hist) norm_hist[it] = nrm;
            s_stop = test_exit && ((nrm < initial_norm) || (nrm < eps2));
        }
        __syncthreads();
        norm = s_warp[0];
        const int stop = s_stop;
        it++;
        __syncthreads();  // s_warp / s_stop are rewritten in the next iteration
        if (stop) { cap = 0; break; }
    }
    for (int c = tid; c < n; c += SMALL_THREADS) {
        const int x = c / ny, y = c - x * ny;
        p[(int64_t)x * g.pitch + y] = sp[c];
    }
    if (tid == 0) {
        ctl->iters_done = it;
        ctl->last_norm = norm;
        ctl->norms[0] = norm;
        ctl->active_T = 0;
        ctl->finished = 1;
        ctl->cap_hit = (cap && max_it > 0) ? 1 : 0;
    }
}

}  // namespace

// grids this path takes: single GPU, everything in one CTA's shared memory
bool sor_small_fits(const sb_sim *s) {
    return s->dbg.sor_small && !s->slab && s->g.NX >= 3 && s->g.NY >= 3 && s->g.NX * s->g.NY <= 12288;
}

// the whole solve; the host has initialised *d_ctl (src, max_iterations) and reads it back
sb_status launch_sor_small(sb_sim *s, double initial_norm, double eps2, int test_exit,
                           double *norm_hist) {
    const size_t n = (size_t)(s->g.NX * s->g.NY);
    const size_t smem = n * 17 + 16;
    static bool attr_set[64] = {false};
    if (!attr_set[s->device & 63]) {
        SB_CUDA(cudaFuncSetAttribute(sor_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     12288 * 17 + 16));
        attr_set[s->device & 63] = true;
    }
    prof_mark(s);
    sor_small_kernel<<<1, SMALL_THREADS, smem, s->stream>>>(
        s->g, rb_pbuf_ptr(s), s->rhs, s->cflag, s->d_ctl, rb_consts(s), s->fluid_cells,
        initial_norm, eps2, test_exit, norm_hist);
    s->launches++;
    prof_mark(s);
    SB_CUDA(cudaGetLastError());
    return SB_OK;
}

void preload_sor_small() {
    cudaFuncAttributes a;
    cudaFuncGetAttributes(&a, sor_small_kernel);
    cudaGetLastError();
}

}  // namespace sb
