// sor_rb.cuh -- shared by the two kernels of the performance-mode SOR pass: the tile kernel
// (sor_rb.cu: walls, obstacles, grid ring, slab edges) and the streaming kernel
// (sor_rb_stream.cu: all-fluid regions).  Both evaluate, per cell,
//     t     = fma(1/dx^2, pE+pW, fma(1/dy^2, pS+pN, -rhs))
//     p_new = fma(mid, t, (1-w)*p)            mid = w / (2/dx^2 + 2/dy^2)
//     r     = fma(-(2/dx^2 + 2/dy^2), p, t)
// (the red-black restatement of /root/reference/src/simulation.rs:253-274 and
// src/math.rs:176-186; identical in oracle/stroemung_oracle.c, SO_SOR_RED_BLACK).
#pragma once

#include "sb_internal.cuh"

namespace sb {

constexpr int RB_TXR = 48;    // tile rows (x)
constexpr int RB_TW = 128;    // tile / strip columns (y)
constexpr int RB_TMAX = 4;    // sweeps fused per pass, at most

struct RbConsts {
    double rdx2, rdy2, diag, mid, omw;
};

inline RbConsts rb_consts(const sb_sim *s) {
    RbConsts k;
    const double dx2 = s->prm.delx * s->prm.delx, dy2 = s->prm.dely * s->prm.dely;
    k.rdx2 = 1.0 / dx2;
    k.rdy2 = 1.0 / dy2;
    k.diag = (2.0 * k.rdx2) + (2.0 * k.rdy2);
    k.mid = s->prm.omega / ((2.0 / dx2) + (2.0 / dy2));
    k.omw = 1.0 - s->prm.omega;
    return k;
}

// the two pressure buffer pointers live behind the control block (capi.cu)
inline double *const *rb_pbuf_ptr(sb_sim *s) {
    return reinterpret_cast<double *const *>(reinterpret_cast<char *>(s->d_ctl) + 256);
}

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
// 1-D bulk copy global -> shared through the TMA engine (16-byte aligned, size % 16 == 0)
__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes,
                                          uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// finalize of a pass done by the streaming kernel's last CTA (single GPU): what the separate
// sor_finalize_kernel (stages.cu) would be launched with
struct RbFin {
    int enabled, test_exit;
    double fluid_cells, initial_norm, eps2;
    double *norm_hist;
    unsigned *counter;   // CTAs done so far (resets itself)
};

// Exit rule of /root/reference/src/simulation.rs:279 on the T norms of a pass and the
// bookkeeping that follows (one thread).  A red-black pass runs T sweeps speculatively; if
// the rule fires at level k < T the pass is repeated from the same source buffer with T = k.
__device__ __forceinline__ void sor_advance_ctl(SorCtl *ctl, const double *level_norm, int T,
                                                double initial_norm, double eps2, int test_exit,
                                                double *norm_hist) {
    int exit_at = 0;
    for (int lvl = 0; lvl < T; lvl++) {
        ctl->norms[lvl] = level_norm[lvl];
        if (norm_hist) norm_hist[ctl->iters_done + lvl] = level_norm[lvl];
        if (test_exit && !exit_at &&
            ((level_norm[lvl] < initial_norm) || (level_norm[lvl] < eps2)))
            exit_at = lvl + 1;
    }
    if (exit_at && exit_at < T) {
        ctl->active_T = exit_at;  // redo this pass with fewer sweeps (same source buffer)
        return;
    }
    ctl->iters_done += T;
    ctl->src ^= (ctl->block_T > 0) ? 1 : 0;  // red-black passes ping-pong; in-place modes don't
    ctl->last_norm = level_norm[T - 1];
    if (exit_at) {
        ctl->active_T = 0;
        ctl->finished = 1;
    } else if (ctl->iters_done >= ctl->max_iterations) {
        ctl->active_T = 0;
        ctl->finished = 1;
        ctl->cap_hit = 1;
    } else {
        uint32_t rem = ctl->max_iterations - ctl->iters_done;
        int next = ctl->block_T > 0 ? ctl->block_T : 1;
        ctl->active_T = (int)(rem < (uint32_t)next ? rem : (uint32_t)next);
    }
}

__device__ __forceinline__ double warp_sum_down(double v) {
    for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace sb
