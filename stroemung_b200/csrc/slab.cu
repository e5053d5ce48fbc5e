// slab.cu -- multi-GPU: row slabs along x, one handle per GPU, halos and scalar reductions
// exchanged through peer memory over NVLink.
//
// The reference has no multi-process code (a single-threaded Rust loop,
// /root/reference/src/simulation.rs:324-333); the decomposition is the build's own.  Slab g
// owns the global rows [x_begin, x_end) with full contiguous y-lines plus SB_SLAB_HALO
// halo rows on each side.  What moves between slabs, and when:
//   p    the red-black pass writes its edge rows straight into the neighbours' halo rows
//        (sor_rb.cu epilogue, P2P stores) -- SB_SLAB_HALO = 2T+2 rows for T = 4 fused sweeps;
//   u, v after the velocity update, by the put kernel below;
//   rhs  after calculate_rhs (the red-black tiles recompute their 2T+2 halo rows, so the
//        right-hand side must be valid there too), put kernel;
//   kind at classification time (put kernel, bytes);
//   residual sums, min/max ranges, fluid counts: slab_allgather() (slab_dev.cuh), summed
//        in rank order on every rank -- deterministic and identical everywhere.
// No NCCL call sits on the data path: the scalar all-reduce is one 1-block kernel whose
// latency is a pair of NVLink round trips, and the host never takes part.
//
// Connection: sb_slab_export() describes this rank's allocations (CUDA IPC handles plus, for
// handles living in the same process, the raw pointers); the host all-gathers the blobs
// (torch.distributed in stroemung_b200/multi.py) and hands them to sb_slab_connect(), which
// maps the peers and finishes construction collectively.
#include <string.h>
#include <unistd.h>

#include <algorithm>

#include "sb_internal.cuh"
#include "slab_dev.cuh"

namespace sb {

namespace {

struct SlabBlob {                 // what sb_slab_export writes (<= SB_SLAB_BLOB_BYTES)
    uint32_t magic, version;
    int32_t rank, world, device, pid;
    uint64_t host_tag;            // distinguishes processes beyond the pid (boot-unique enough)
    int64_t nxl, pitch, own0, own1, NX, NY;
    uint64_t raw[7];              // p0, p1, u, v, cflag, mbox, rhs: same-process access
    cudaIpcMemHandle_t ipc[7];    // the same allocations for other processes
};
static_assert(sizeof(SlabBlob) <= SB_SLAB_BLOB_BYTES, "blob must fit the ABI constant");
constexpr uint32_t BLOB_MAGIC = 0x53423230u;  // "SB20"

__global__ void allreduce_kernel(SlabLink lk, double *vals, int n, unsigned ops) {
    __shared__ double gathered[SB_MAX_WORLD * 8];
    __shared__ double mine[8];
    if ((int)threadIdx.x < n) mine[threadIdx.x] = vals[threadIdx.x];
    __syncthreads();
    if (!slab_allgather(lk, mine, n, gathered)) return;
    if ((int)threadIdx.x < n) {
        const int i = threadIdx.x;
        const unsigned op = (ops >> (2 * i)) & 3u;
        double acc = gathered[i];
        for (int r = 1; r < lk.world; r++) {
            const double x = gathered[r * 8 + i];
            acc = op == XR_SUM ? acc + x : op == XR_MIN ? fmin(acc, x) : fmax(acc, x);
        }
        vals[i] = acc;
    }
}

// 16-byte copies of `n16` units per side: my first H owned rows -> lower neighbour, my last
// H owned rows -> upper neighbour.  Rows are pitch-contiguous, so each side is one range.
__global__ void put_rows_kernel(const uint4 *__restrict__ src_lo, uint4 *__restrict__ dst_lo,
                                const uint4 *__restrict__ src_hi, uint4 *__restrict__ dst_hi,
                                size_t n16) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
        if (dst_lo) dst_lo[i] = src_lo[i];
        if (dst_hi) dst_hi[i] = src_hi[i];
    }
}

uint64_t process_tag() {
    // pid alone could collide across containers sharing a GPU box; add the hostname hash
    char host[256] = {0};
    gethostname(host, sizeof(host) - 1);
    uint64_t h = 1469598103934665603ull;
    for (const char *c = host; *c; c++) h = (h ^ (uint64_t)(unsigned char)*c) * 1099511628211ull;
    return h;
}

}  // namespace

sb_status slab_check_error(sb_sim *s) {
    if (!s->slab || !s->connected) return SB_OK;
    int32_t e = 0;
    SB_CUDA(cudaMemcpyAsync(&e, s->d_xerr, sizeof(e), cudaMemcpyDeviceToHost, s->stream));
    SB_CUDA(cudaStreamSynchronize(s->stream));
    if (e) {
        set_error("slab exchange timed out: a peer rank did not reach the same collective call");
        return SB_CUDA_ERROR;
    }
    return SB_OK;
}

sb_status slab_allreduce(sb_sim *s, double *d_vals, int n, unsigned ops) {
    if (!s->slab) return SB_OK;
    if (!s->connected) {
        set_error("slab handle is not connected yet (sb_slab_connect)");
        return SB_INVALID_ARGUMENT;
    }
    allreduce_kernel<<<1, 32, 0, s->stream>>>(s->link, d_vals, n, ops);
    s->launches++;
    SB_CUDA(cudaGetLastError());
    return SB_OK;
}

sb_status slab_put_rows(sb_sim *s, const void *field, void *lo_field, void *hi_field,
                        size_t esize) {
    if (!s->slab) return SB_OK;
    const Geom &g = s->g;
    const size_t row_bytes = (size_t)g.pitch * esize;  // pitch % 16 == 0 -> 16-byte units
    const size_t n16 = (size_t)s->link.H * row_bytes / 16;
    const char *base = static_cast<const char *>(field);
    const uint4 *src_lo = reinterpret_cast<const uint4 *>(base + (size_t)g.own0 * row_bytes);
    const uint4 *src_hi =
        reinterpret_cast<const uint4 *>(base + (size_t)(g.own1 - s->link.H) * row_bytes);
    uint4 *dst_lo = lo_field ? reinterpret_cast<uint4 *>(static_cast<char *>(lo_field) +
                                                         (size_t)s->link.lo_row0 * row_bytes)
                             : nullptr;
    uint4 *dst_hi = hi_field ? reinterpret_cast<uint4 *>(static_cast<char *>(hi_field) +
                                                         (size_t)s->link.hi_row0 * row_bytes)
                             : nullptr;
    if (!dst_lo && !dst_hi) return SB_OK;
    int blocks = (int)std::min<size_t>((n16 + 255) / 256, 296);
    put_rows_kernel<<<blocks, 256, 0, s->stream>>>(src_lo, dst_lo, src_hi, dst_hi, n16);
    s->launches++;
    SB_CUDA(cudaGetLastError());
    return SB_OK;
}

sb_status slab_sync_halos(sb_sim *s, int with_flags) {
    if (!s->slab) return SB_OK;
    sb_status st;
    // nobody may still be reading the halo rows that are about to be overwritten
    if ((st = slab_allreduce(s, s->d_scalars, 0, 0))) return st;
    const int c = s->cur;
    if ((st = slab_put_rows(s, s->p[c], s->link.lo_p[c], s->link.hi_p[c], 8))) return st;
    if ((st = slab_put_rows(s, s->u, s->lo_u, s->hi_u, 8))) return st;
    if ((st = slab_put_rows(s, s->v, s->lo_v, s->hi_v, 8))) return st;
    if (with_flags) {
        s->flag_epoch++;  // the neighbours rewrite my halo rows' flags as well
        if ((st = slab_put_rows(s, s->cflag, s->lo_flag, s->hi_flag, 1))) return st;
    }
    return slab_allreduce(s, s->d_scalars, 0, 0);
}

void slab_release(sb_sim *s) {
    for (void *p : s->ipc_opened) cudaIpcCloseMemHandle(p);
    s->ipc_opened.clear();
    cudaFree(s->d_mbox);
    cudaFree(s->d_xseq);
    cudaFree(s->d_xerr);
    s->d_mbox = nullptr;
    s->d_xseq = nullptr;
    s->d_xerr = nullptr;
}

// allocate the mailbox etc. (called from allocate() in capi.cu for world > 1)
sb_status slab_prepare(sb_sim *s) {
    const size_t mb = 2 * SB_MAX_WORLD * sizeof(MailSlot);
    SB_CUDA(cudaMalloc(&s->d_mbox, mb));
    SB_CUDA(cudaMemsetAsync(s->d_mbox, 0, mb, s->stream));
    SB_CUDA(cudaMalloc(&s->d_xseq, sizeof(unsigned long long)));
    SB_CUDA(cudaMemsetAsync(s->d_xseq, 0, sizeof(unsigned long long), s->stream));
    SB_CUDA(cudaMalloc(&s->d_xerr, sizeof(int32_t)));
    SB_CUDA(cudaMemsetAsync(s->d_xerr, 0, sizeof(int32_t), s->stream));
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, allreduce_kernel);
    cudaFuncGetAttributes(&fa, put_rows_kernel);
    cudaGetLastError();
    preload_classify();
    preload_grid();
    preload_stages();
    preload_sor_rb();
    preload_sor_rb_stream();
    preload_render();
    preload_sor_small();
    preload_sor_mid();
    s->slab = true;
    s->connected = false;
    memset(&s->link, 0, sizeof(s->link));
    s->link.rank = s->prm.rank;
    s->link.world = s->prm.world;
    s->link.H = SB_SLAB_HALO;
    s->link.seq = s->d_xseq;
    s->link.err = s->d_xerr;
    return SB_OK;
}

}  // namespace sb

using namespace sb;

extern "C" {

sb_status sb_slab_export(sb_sim *sim, uint8_t blob[SB_SLAB_BLOB_BYTES]) {
    if (!sim || !blob) return SB_INVALID_ARGUMENT;
    SB_CUDA(cudaSetDevice(sim->device));
    if (!sim->slab) {
        set_error("sb_slab_export: the handle was not created with world > 1");
        return SB_INVALID_ARGUMENT;
    }
    SB_CUDA(cudaStreamSynchronize(sim->stream));  // mailbox zeroed, own rows uploaded
    SlabBlob b;
    memset(&b, 0, sizeof(b));
    b.magic = BLOB_MAGIC;
    b.version = 1;
    b.rank = sim->prm.rank;
    b.world = sim->prm.world;
    b.device = sim->device;
    b.pid = (int32_t)getpid();
    b.host_tag = process_tag();
    b.nxl = sim->g.nxl; b.pitch = sim->g.pitch; b.own0 = sim->g.own0; b.own1 = sim->g.own1;
    b.NX = sim->g.NX; b.NY = sim->g.NY;
    void *ptrs[7] = {sim->p[0], sim->p[1], sim->u, sim->v, sim->cflag, sim->d_mbox, sim->rhs};
    for (int i = 0; i < 7; i++) {
        b.raw[i] = (uint64_t)(uintptr_t)ptrs[i];
        SB_CUDA(cudaIpcGetMemHandle(&b.ipc[i], ptrs[i]));
    }
    memset(blob, 0, SB_SLAB_BLOB_BYTES);
    memcpy(blob, &b, sizeof(b));
    return SB_OK;
}

sb_status sb_slab_connect(sb_sim *sim, const uint8_t *blobs, size_t n_blobs) {
    if (!sim || !blobs) return SB_INVALID_ARGUMENT;
    SB_CUDA(cudaSetDevice(sim->device));
    if (!sim->slab || sim->connected || (int)n_blobs != sim->prm.world) {
        set_error("sb_slab_connect: needs an unconnected slab handle and `world` blobs");
        return SB_INVALID_ARGUMENT;
    }
    const int me = sim->prm.rank, world = sim->prm.world;
    const int32_t my_pid = (int32_t)getpid();
    const uint64_t my_tag = process_tag();
    std::vector<SlabBlob> bs(world);
    for (int r = 0; r < world; r++) {
        memcpy(&bs[r], blobs + (size_t)r * SB_SLAB_BLOB_BYTES, sizeof(SlabBlob));
        const SlabBlob &b = bs[r];
        if (b.magic != BLOB_MAGIC || b.rank != r || b.world != world || b.NX != sim->g.NX ||
            b.NY != sim->g.NY || b.pitch != sim->g.pitch) {
            set_error("sb_slab_connect: blob " + std::to_string(r) + " does not match this run");
            return SB_INVALID_ARGUMENT;
        }
    }
    // map allocation `which` of rank r into this process
    auto map = [&](int r, int which, void **out) -> sb_status {
        const SlabBlob &b = bs[r];
        if (r == me) {
            *out = (void *)(uintptr_t)b.raw[which];
            return SB_OK;
        }
        if (b.pid == my_pid && b.host_tag == my_tag) {  // another handle of this process
            if (b.device != sim->device) {
                int can = 0;
                SB_CUDA(cudaDeviceCanAccessPeer(&can, sim->device, b.device));
                if (!can) {
                    set_error("no peer access between the devices of two slabs");
                    return SB_CUDA_ERROR;
                }
                cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) SB_CUDA(e);
                cudaGetLastError();
            }
            *out = (void *)(uintptr_t)b.raw[which];
            return SB_OK;
        }
        void *p = nullptr;
        SB_CUDA(cudaIpcOpenMemHandle(&p, b.ipc[which], cudaIpcMemLazyEnablePeerAccess));
        sim->ipc_opened.push_back(p);
        *out = p;
        return SB_OK;
    };
    sb_status st;
    SlabLink &lk = sim->link;
    for (int r = 0; r < world; r++) {
        void *p = nullptr;
        if ((st = map(r, 5, &p))) return st;
        lk.mbox[r] = static_cast<MailSlot *>(p);
    }
    if (me > 0) {
        void *p[7];
        for (int i = 0; i < 7; i++)
            if (i != 5 && (st = map(me - 1, i, &p[i]))) return st;
        lk.lo_p[0] = (double *)p[0]; lk.lo_p[1] = (double *)p[1];
        sim->lo_u = (double *)p[2]; sim->lo_v = (double *)p[3]; sim->lo_flag = (uint8_t *)p[4];
        sim->lo_rhs = (double *)p[6];
        lk.lo_row0 = bs[me - 1].own1;  // its upper halo rows
    }
    if (me + 1 < world) {
        void *p[7];
        for (int i = 0; i < 7; i++)
            if (i != 5 && (st = map(me + 1, i, &p[i]))) return st;
        lk.hi_p[0] = (double *)p[0]; lk.hi_p[1] = (double *)p[1];
        sim->hi_u = (double *)p[2]; sim->hi_v = (double *)p[3]; sim->hi_flag = (uint8_t *)p[4];
        sim->hi_rhs = (double *)p[6];
        lk.hi_row0 = bs[me + 1].own0 - lk.H;  // its lower halo rows (= 0)
    }
    sim->connected = true;
    // collective part of construction: halos of the uploaded state, then try_from's work
    if ((st = slab_sync_halos(sim, 1))) return st;
    if ((st = finish_create(sim))) return st;
    return slab_check_error(sim);
}

sb_status sb_slab_sync_halos(sb_sim *sim) {
    if (!sim) return SB_INVALID_ARGUMENT;
    SB_CUDA(cudaSetDevice(sim->device));
    sb_status st = slab_sync_halos(sim, 0);
    if (st) return st;
    return slab_check_error(sim);
}

}  // extern "C"
