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
// east/south values are still untouched in memory.
//
// Bands hand over column by column through a flag-in-data buffer (the "LL" scheme of
// collective libraries): the lane that owns a band's last row stores every new value as
// {lo32, seq, hi32, seq} -- two 8-byte words that each carry the sweep's sequence number --
// and lane 0 of the next band polls that element until both words show the sequence number.
// Aligned 8-byte accesses are single-copy atomic, so a value is either complete or not there
// yet: no fence, no release/acquire, no L1 invalidation anywhere in the sweep, and a band runs
// a handful of columns behind the one above instead of a publish interval.  (The first version
// published a progress counter every 16 columns behind __threadfence: 6.0 ms per sweep at
// 2048^2; moving every memory operand off the step's critical path 3.2 ms; this 0.x ms --
// profiles/r2_lex_history.txt.)  Bands take their number from a ticket so a band only ever
// waits for bands that already run (no co-residency assumption).
//
// Arithmetic is strict IEEE in the reference's association (-fmad=false, true
// divisions), so the pressure field equals the reference's bit for bit.
#include "sb_internal.cuh"

namespace sb {

namespace {

constexpr int LEX_AHEAD = 4;      // columns every operand is loaded ahead of its use

__device__ __forceinline__ uint4 ld_ll(const uint4 *ptr) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(ptr) : "memory");
    return v;
}
__device__ __forceinline__ void st_ll(uint4 *ptr, double val, uint32_t seq) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(val);
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(ptr),
                 "r"((uint32_t)b), "r"(seq), "r"((uint32_t)(b >> 32)), "r"(seq) : "memory");
}

// The operands of one cell update that come from memory
struct LexOps {
    double s, e, rhs;   // (x, y+1) old, (x+1, y) old, rhs
    uint4 w;            // (x-1, y) new as a hand-over element [lane 0 of bands > 0]
    double w0;          // (x-1, y) of the ring row [lane 0 of band 0]
    unsigned fl;
};

// A step's critical path is shuffle -> add -> divide -> add -> add -> mul -> add; every memory
// operand of column y is therefore requested LEX_AHEAD steps before its use into a register
// FIFO that rotates by NAME (the step loop is unrolled LEX_AHEAD times), so no load latency
// -- in particular not the L2 round trip of the hand-over element -- sits on that path.  None
// of the p cells can change in between: (x, y+1) and (x+1, y) are written by this sweep only
// after (x, y); a hand-over element that was requested too early fails its tag test and is
// polled again.
__global__ void __launch_bounds__(32)
sor_lex_kernel(Geom g, double *const *__restrict__ pbuf, const SorCtl *__restrict__ ctl,
               int guarded, const double *__restrict__ rhs, const uint8_t *__restrict__ cflag,
               int *__restrict__ sync, uint4 *__restrict__ ll, int64_t ll_pitch, uint32_t seq,
               double delx2, double dely2, double one_minus_w, double middle) {
    if (guarded && ctl->active_T == 0) return;
    double *p = pbuf[ctl->src];
    const int lane = threadIdx.x;
    int band = 0;
    if (lane == 0) band = atomicAdd(&sync[0], 1);
    band = __shfl_sync(0xffffffffu, band, 0);

    const int64_t gx = 1 + (int64_t)band * 32 + lane;  // global row of this lane
    const bool row_ok = gx <= g.NX - 2;
    const int64_t lx = gx - g.gx0;
    const int64_t row = lx * g.pitch;
    // the lane that owns the band's last interior row hands it to the next band
    const int64_t last_row_gx = min((int64_t)1 + (int64_t)band * 32 + 31, g.NX - 2);
    const bool publisher = gx == last_row_gx && gx < g.NX - 2;
    uint4 *ll_out = ll + (int64_t)band * ll_pitch;
    const uint4 *ll_in = ll + (int64_t)(band > 0 ? band - 1 : 0) * ll_pitch;
    const int ny2 = (int)g.NY - 2;  // last interior column
    const bool handover = lane == 0 && band > 0;   // the west value comes from another band

    double p_c = 0.0;      // old value of the cell this lane updates next
    double last_new = 0.0; // value of (x, y-1) after this sweep (new north neighbour)
    if (row_ok) {
        last_new = p[row + 0];  // (x, 0): ring cell, never updated by the sweep
        p_c = p[row + 1];
    }
    auto fetch = [&](LexOps &o, int y) {
        o.s = o.e = o.rhs = o.w0 = 0.0;
        o.w = make_uint4(0u, 0u, 0u, 0u);
        o.fl = 0xffu;   // not a fluid cell
        if (!(row_ok && y >= 1 && y <= ny2)) return;
        const int64_t c = row + y;
        o.s = p[c + 1];
        o.e = p[c + g.pitch];
        o.rhs = rhs[c];
        o.fl = cflag[c];
        if (handover) o.w = ld_ll(ll_in + y);
        else if (lane == 0) o.w0 = p[c - g.pitch];  // global row 0: ring, constant in the sweep
    };
    LexOps F[LEX_AHEAD];
#pragma unroll
    for (int j = 0; j < LEX_AHEAD; j++) fetch(F[j], 1 - lane + j);
    const int steps = ny2 + 31;
    for (int t0 = 0; t0 < steps; t0 += LEX_AHEAD) {
#pragma unroll
        for (int j = 0; j < LEX_AHEAD; j++) {
            const int y = 1 + t0 + j - lane;
            const bool active = row_ok && y >= 1 && y <= ny2;
            // west neighbour, new value: previous step's result of lane-1
            double p_w = __shfl_up_sync(0xffffffffu, last_new, 1);
            if (lane == 0) {
                p_w = F[j].w0;
                if (handover && active) {
                    uint4 w = F[j].w;
                    while (w.y != seq || w.w != seq) w = ld_ll(ll_in + y);   // not there yet
                    p_w = __longlong_as_double((long long)(((unsigned long long)w.z << 32) | w.x));
                }
            }
            if (active) {
                const double p_s = F[j].s;      // (x, y+1) old
                const double p_e = F[j].e;      // (x+1, y) old
                const double p_n = last_new;    // (x, y-1) new
                double p_new = p_c;
                if (cf_is_fluid((uint8_t)F[j].fl)) {
                    p_new = (one_minus_w * p_c) +
                            middle * ((((p_e + p_w) / delx2) + ((p_s + p_n) / dely2)) - F[j].rhs);
                    p[row + y] = p_new;
                }
                last_new = p_new;
                p_c = p_s;
                if (publisher) st_ll(ll_out + y, p_new, seq);
            }
            fetch(F[j], y + LEX_AHEAD);
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
    // hand-over rows: one per band, tagged with the sweep's sequence number (never reset)
    const int64_t ll_pitch = (g.NY + 7) & ~(int64_t)7;
    const size_t ll_need = (size_t)nbands * (size_t)ll_pitch;
    if (ll_need > s->lex_ll_cap) {
        if (s->d_lex_ll) cudaFree(s->d_lex_ll);
        s->d_lex_ll = nullptr;
        s->lex_ll_cap = 0;
        SB_CUDA(cudaMalloc(&s->d_lex_ll, ll_need * sizeof(uint4)));
        SB_CUDA(cudaMemsetAsync(s->d_lex_ll, 0, ll_need * sizeof(uint4), s->stream));
        s->lex_ll_cap = ll_need;
        s->lex_seq = 0;
    }
    if (++s->lex_seq == 0) {   // the 32-bit tag wrapped: start over on a clean buffer
        SB_CUDA(cudaMemsetAsync(s->d_lex_ll, 0, s->lex_ll_cap * sizeof(uint4), s->stream));
        s->lex_seq = 1;
    }
    double delx2 = s->prm.delx * s->prm.delx;
    double dely2 = s->prm.dely * s->prm.dely;
    double one_minus_w = 1.0 - s->prm.omega;
    double middle = s->prm.omega / ((2.0 / delx2) + (2.0 / dely2));
    prof_mark(s);
    sor_lex_kernel<<<nbands, 32, 0, s->stream>>>(g, pbuf_ptr(s), s->d_ctl, guarded, s->rhs,
                                                 s->cflag, s->d_lex_sync, s->d_lex_ll, ll_pitch,
                                                 s->lex_seq, delx2, dely2, one_minus_w, middle);
    prof_mark(s);
    s->launches++;
    SB_CUDA(cudaGetLastError());
    return SB_OK;
}

}  // namespace sb
