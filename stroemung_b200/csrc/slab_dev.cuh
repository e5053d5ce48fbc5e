// slab_dev.cuh -- device side of the cross-slab exchange (row slabs along x, one per GPU).
//
// The reference is a single-process solver: the slab decomposition,
// and with it everything in this file, is the build's own.  Ranks talk through memory of
// their peers mapped into this process (CUDA IPC over NVLink 5 / NVSwitch, or plain peer
// access between handles of one process): plain stores for payload, a system-scope
// release store of a round number as the flag, an acquire load on the receiving side.
//
// slab_allgather(): every rank posts up to 8 doubles into slot [round & 1][rank] of EVERY
// rank's mailbox and then waits until all `world` slots of its own mailbox carry the
// current round.  All ranks reduce the gathered values in rank order, so every rank gets
// bit-identical results and takes identical control decisions (SOR exit test) without a
// host round trip.  Because a rank can only reach round r+1 after everyone has posted
// round r, two slot parities are enough (a slot is rewritten two rounds later, after its
// reader has left the round it was read in).
// The round doubles as a barrier: stores a rank made to peer memory in EARLIER kernels of
// its stream (halo rows written by the red-black pass epilogue or the put kernel) are
// complete before the flag store of this kernel (kernel boundary + fence.sys), so a peer
// that has seen the flag may read them in its following kernels.
#pragma once

#include "sb_internal.cuh"

namespace sb {

__device__ __forceinline__ void st_release_sys_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys_f64(double *p, double v) {
    asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ double ld_relaxed_sys_f64(const double *p) {
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

// a wait gives up after this many SM clocks (~20 s): a missing peer must end in an error
// code on the host, never in a hung GPU
constexpr long long SLAB_WAIT_CLOCKS = 40LL * 1000 * 1000 * 1000;

// Block-wide.  my_vals[0..n): this rank's contribution (any memory space).
// gathered[r * 8 + i] (shared memory, world * 8 doubles) receives value i of rank r.
// Needs blockDim.x >= world.  Returns false (block-uniform) if the round failed.
__device__ __forceinline__ bool slab_allgather(const SlabLink &lk, const double *my_vals, int n,
                                               double *gathered) {
    __shared__ unsigned long long s_round;
    __shared__ int s_bad;
    if (threadIdx.x == 0) {
        s_round = *lk.seq + 1;
        s_bad = *reinterpret_cast<volatile int32_t *>(lk.err);
    }
    __syncthreads();
    const unsigned long long round = s_round;
    if (s_bad) return false;
    const int t = threadIdx.x;
    if (t < lk.world) {
        MailSlot *dst = lk.mbox[t] + (round & 1) * SB_MAX_WORLD + lk.rank;
        for (int i = 0; i < n; i++) st_relaxed_sys_f64(&dst->vals[i], my_vals[i]);
        // the release store orders the values (and this thread's earlier stores) before the
        // flag by itself: no separate fence in front of it
        st_release_sys_u64(&dst->seq, round);
        const MailSlot *src = lk.mbox[lk.rank] + (round & 1) * SB_MAX_WORLD + t;
        const long long t0 = clock64();
        bool ok = true;
        // relaxed polls, one acquire fence once the flag is there
        while (ld_relaxed_sys_u64(&src->seq) != round) {
            if (clock64() - t0 > SLAB_WAIT_CLOCKS) {
                ok = false;
                break;
            }
        }
        asm volatile("fence.acq_rel.sys;" ::: "memory");
        if (ok) {
            for (int i = 0; i < n; i++) gathered[t * 8 + i] = ld_relaxed_sys_f64(&src->vals[i]);
        } else {
            atomicExch(&s_bad, 1);
            *lk.err = 1;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) *lk.seq = round;
    return s_bad == 0;
}

}  // namespace sb
