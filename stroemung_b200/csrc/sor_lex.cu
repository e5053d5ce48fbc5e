// sor_lex.cu -- K4a: reference-order (lexicographic, in-place) SOR sweep as a wavefront.
//
// Replaces the sweep loop of Simulation::solve_sor
// (/root/reference/src/simulation.rs:253-274):
//     for x in 1..nx-1 { for y in 1..ny-1 { if Fluid {
//         p = (1-w)*p + mid*(((pE+pW)/dx^2) + ((pS+pN)/dy^2) - rhs) }}}      in place
// Cell (x, y) needs the NEW values of (x-1, y) and (x, y-1) and the OLD values of
// (x+1, y) and (x, y+1).  One warp owns a band of 32 consecutive rows and walks it as
// a skewed wavefront: at step t lane l updates (x0+l, 1+t-l).  The new west value comes
// from lane l-1 by shuffle, the new north value is the lane's own previous result, old
// east/south values are still untouched in memory.  Bands hand over through global
// memory: band b may work on column y once band b-1 has published it (progress counter,
// release/acquire by __threadfence).  Bands take their number from a ticket so a band
// only ever waits for bands that already run (no co-residency assumption).
//
// Arithmetic is strict IEEE in the reference's association (-fmad=false, true
// divisions), so the pressure field equals the reference's bit for bit.
#include "sb_internal.cuh"

namespace sb {

namespace {

constexpr int LEX_PUBLISH = 16;   // publish progress every this many columns
constexpr int LEX_PREFETCH = 32;  // columns ahead for L1 prefetch

__device__ __forceinline__ void prefetch_l1(const void *ptr) {
    asm volatile("prefetch.global.L1 [%0];" ::"l"(ptr));
}
__device__ __forceinline__ double ld_cg(const double *ptr) {
    double v;
    asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(ptr));
    return v;
}
__device__ __forceinline__ int ld_volatile(const int *ptr) {
    int v;
    asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(ptr));
    return v;
}

__global__ void __launch_bounds__(32)
sor_lex_kernel(Geom g, double *const *__restrict__ pbuf, const SorCtl *__restrict__ ctl,
               int guarded, const double *__restrict__ rhs, const uint8_t *__restrict__ cflag,
               int *__restrict__ sync, double delx2, double dely2, double one_minus_w,
               double middle) {
    if (guarded && ctl->active_T == 0) return;
    double *p = pbuf[ctl->src];
    const int lane = threadIdx.x;
    int band = 0;
    if (lane == 0) band = atomicAdd(&sync[0], 1);
    band = __shfl_sync(0xffffffffu, band, 0);
    int *progress = sync + 1;  // progress[b]: columns <= value are final in band b's last row

    const int64_t gx = 1 + (int64_t)band * 32 + lane;  // global row of this lane
    const bool row_ok = gx <= g.NX - 2;
    const int64_t lx = gx - g.gx0;
    const int64_t row = lx * g.pitch;
    // the lane that owns the band's last interior row publishes progress
    const int64_t last_row_gx = min((int64_t)1 + (int64_t)band * 32 + 31, g.NX - 2);
    const bool publisher = gx == last_row_gx;
    const int64_t ny2 = g.NY - 2;  // last interior column
    int known = band == 0 ? 0x7fffffff : 0;  // columns of band-1 known to be final (lane 0)

    double p_c = 0.0;      // old value of the cell this lane updates next
    double last_new = 0.0; // value of (x, y-1) after this sweep (new north neighbour)
    if (row_ok) {
        last_new = p[row + 0];  // (x, 0): ring cell, never updated by the sweep
        p_c = p[row + 1];
    }
    const int64_t steps = ny2 + 31;
    for (int64_t t = 0; t < steps; t++) {
        const int64_t y = 1 + t - lane;
        const bool active = row_ok && y >= 1 && y <= ny2;
        // west neighbour, new value: previous step's result of lane-1
        double p_w = __shfl_up_sync(0xffffffffu, last_new, 1);
        if (lane == 0 && active) {
            if (band > 0) {
                if ((int)y > known) {
                    do { known = ld_volatile(&progress[band - 1]); } while (known < (int)y);
                    __threadfence();
                }
                p_w = ld_cg(&p[row - g.pitch + y]);
            } else {
                p_w = p[row - g.pitch + y];  // global row 0: ring, constant during the sweep
            }
        }
        if (active) {
            const int64_t c = row + y;
            if (((y + LEX_PREFETCH) & 15) == 0 && y + LEX_PREFETCH < g.NY) {
                prefetch_l1(&p[c + LEX_PREFETCH]);
                prefetch_l1(&rhs[c + LEX_PREFETCH]);
                if (lane == 31 || gx == g.NX - 2) prefetch_l1(&p[c + g.pitch + LEX_PREFETCH]);
                if (((y + LEX_PREFETCH) & 127) == 0) prefetch_l1(&cflag[c + LEX_PREFETCH]);
            }
            const double p_s = p[c + 1];        // (x, y+1) old
            const double p_e = p[c + g.pitch];  // (x+1, y) old
            const double p_n = last_new;        // (x, y-1) new
            double p_new = p_c;
            if (cf_is_fluid(cflag[c])) {
                p_new = (one_minus_w * p_c) +
                        middle * ((((p_e + p_w) / delx2) + ((p_s + p_n) / dely2)) - rhs[c]);
                p[c] = p_new;
            }
            last_new = p_new;
            p_c = p_s;
            if (publisher && ((y % LEX_PUBLISH) == 0 || y == ny2)) {
                __threadfence();
                *(volatile int *)&progress[band] = (int)y;
            }
        }
    }
}

}  // namespace

static double *const *pbuf_ptr(sb_sim *s) {
    return reinterpret_cast<double *const *>(reinterpret_cast<char *>(s->d_ctl) + 256);
}

sb_status launch_sor_lex_sweep(sb_sim *s, int guarded) {
    const Geom &g = s->g;
    if (g.NX < 3 || g.NY < 3) return SB_OK;
    int nbands = (int)((g.NX - 2 + 31) / 32);
    size_t need = (size_t)nbands + 1;
    if (need > s->lex_sync_cap) {
        if (s->d_lex_sync) cudaFree(s->d_lex_sync);
        s->d_lex_sync = nullptr;
        SB_CUDA(cudaMalloc(&s->d_lex_sync, need * sizeof(int32_t)));
        s->lex_sync_cap = need;
    }
    SB_CUDA(cudaMemsetAsync(s->d_lex_sync, 0, need * sizeof(int32_t), s->stream));
    double delx2 = s->prm.delx * s->prm.delx;
    double dely2 = s->prm.dely * s->prm.dely;
    double one_minus_w = 1.0 - s->prm.omega;
    double middle = s->prm.omega / ((2.0 / delx2) + (2.0 / dely2));
    prof_mark(s);
    sor_lex_kernel<<<nbands, 32, 0, s->stream>>>(g, pbuf_ptr(s), s->d_ctl, guarded, s->rhs,
                                                 s->cflag, s->d_lex_sync, delx2, dely2,
                                                 one_minus_w, middle);
    prof_mark(s);
    s->launches++;
    SB_CUDA(cudaGetLastError());
    return SB_OK;
}

}  // namespace sb
