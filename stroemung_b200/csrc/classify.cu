// classify.cu -- K0: boundary classification (integer work, bit-exact).
//
// Replaces SimulationGrid::rebuild_boundary_list / neighbors / calculate_edges
// (/root/reference/src/grid/mod.rs:167-235, 270-332):
//   * every non-Fluid cell goes on the boundary list, sorted x-major
//     (BTreeSet<BoundaryIndex>, src/types.rs:18-19) == increasing linear index,
//   * fluid_cells counts Fluid cells anywhere in the array,
//   * the 4-bit "neighbour is Fluid" mask (W, E, N, S; out of grid = not fluid)
//     selects the edge class; any other mask is BoundaryTooThinError and the
//     FIRST offender in x-major order is the one reported.
#include "sb_internal.cuh"

namespace sb {

namespace {

constexpr int CL_THREADS = 256;
constexpr int CL_ITEMS = 4;
constexpr int CL_CHUNK = CL_THREADS * CL_ITEMS;

// mask bits: W=8, E=4, N=2, S=1 -> edge class (0xFF = too thin); src/grid/mod.rs:297-331
__constant__ uint8_t c_edge_of_mask[16] = {
    SB_EDGE_NONE, SB_EDGE_S, SB_EDGE_N, 0xFF,        // 0000 0001 0010 0011
    SB_EDGE_E,    SB_EDGE_SE, SB_EDGE_NE, 0xFF,      // 0100 0101 0110 0111
    SB_EDGE_W,    SB_EDGE_SW, SB_EDGE_NW, 0xFF,      // 1000 1001 1010 1011
    0xFF,         0xFF,       0xFF,       0xFF};     // 11xx

__global__ void edge_kernel(uint8_t *__restrict__ cflag, Geom g,
                            unsigned long long *__restrict__ err,
                            unsigned long long *__restrict__ fluid_owned) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t total = g.nxl * g.NY;
    unsigned long long fluid = 0;
    if (i < total) {
        int64_t lx = i / g.NY, y = i - lx * g.NY;
        int64_t c = lx * g.pitch + y;
        uint8_t fl = cflag[c];
        if (fl & CF_VALID) {
            int kind = cf_kind(fl);
            if (kind == SB_KIND_FLUID) {
                if (lx >= g.own0 && lx < g.own1) fluid = 1;
                // CF_NEAR: some 4-neighbour inside the grid is not fluid (neighbours this
                // slab cannot see count as "not fluid": the flag only has to be conservative)
                bool near = false;
                if (g.gx0 + lx > 0) near |= !(lx > 0 && cf_is_fluid(cflag[c - g.pitch]));
                if (g.gx0 + lx < g.NX - 1)
                    near |= !(lx + 1 < g.nxl && cf_is_fluid(cflag[c + g.pitch]));
                if (y > 0) near |= !cf_is_fluid(cflag[c - 1]);
                if (y + 1 < g.NY) near |= !cf_is_fluid(cflag[c + 1]);
                cflag[c] = (uint8_t)(CF_FLUID | (near ? CF_NEAR : 0));
            } else {
                // neighbour bytes may be rewritten concurrently, but only their edge
                // bits change; kind and valid bits are stable
                bool w = lx > 0 && cf_is_fluid(cflag[c - g.pitch]);
                bool e = lx + 1 < g.nxl && cf_is_fluid(cflag[c + g.pitch]);
                bool n = y > 0 && cf_is_fluid(cflag[c - 1]);
                bool s = y + 1 < g.NY && cf_is_fluid(cflag[c + 1]);
                int m = (w << 3) | (e << 2) | (n << 1) | (int)s;
                uint8_t edge = c_edge_of_mask[m];
                if (edge == 0xFF) {
                    // rows whose x-neighbours lie outside this slab cannot be judged here
                    bool judged = (lx > 0 || g.gx0 + lx == 0) &&
                                  (lx + 1 < g.nxl || g.gx0 + lx == g.NX - 1);
                    if (judged && lx >= g.own0 && lx < g.own1)
                        atomicMin(err, (unsigned long long)((g.gx0 + lx) * g.NY + y));
                    edge = 0;
                }
                cflag[c] = (uint8_t)(CF_VALID | kind | (edge << 3));
            }
        }
    }
    // block-level count of owned fluid cells -> one atomic per block
    __shared__ unsigned long long sh[CL_THREADS / 32];
    for (int o = 16; o; o >>= 1) fluid += __shfl_down_sync(0xffffffffu, fluid, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = fluid;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int k = 0; k < CL_THREADS / 32; k++) t += sh[k];
        if (t) atomicAdd(fluid_owned, t);
    }
}

__device__ __forceinline__ bool is_listed(const uint8_t *cflag, const Geom &g, int64_t i) {
    if (i >= g.nxl * g.NY) return false;
    int64_t lx = i / g.NY, y = i - lx * g.NY;
    return cf_is_boundary(cflag[lx * g.pitch + y]);
}

// per-chunk count of boundary cells
__global__ void count_kernel(const uint8_t *__restrict__ cflag, Geom g,
                             int64_t *__restrict__ counts) {
    int64_t base = (int64_t)blockIdx.x * CL_CHUNK + (int64_t)threadIdx.x * CL_ITEMS;
    int c = 0;
#pragma unroll
    for (int k = 0; k < CL_ITEMS; k++) c += is_listed(cflag, g, base + k);
    __shared__ int sh[CL_THREADS / 32];
    for (int o = 16; o; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int k = 0; k < CL_THREADS / 32; k++) t += sh[k];
        counts[blockIdx.x] = t;
    }
}

// single-block exclusive scan of counts[0..n); counts[n] receives the total
__global__ void scan_kernel(int64_t *__restrict__ counts, int64_t n) {
    __shared__ int64_t warp_sums[32];
    __shared__ int64_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < n; base += blockDim.x) {
        int64_t i = base + threadIdx.x;
        int64_t v = i < n ? counts[i] : 0;
        int64_t incl = v;
        for (int o = 1; o < 32; o <<= 1) {
            int64_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((threadIdx.x & 31) >= o) incl += t;
        }
        if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            int nw = blockDim.x >> 5;
            int64_t w = threadIdx.x < nw ? warp_sums[threadIdx.x] : 0;
            int64_t wi = w;
            for (int o = 1; o < 32; o <<= 1) {
                int64_t t = __shfl_up_sync(0xffffffffu, wi, o);
                if ((int)threadIdx.x >= o) wi += t;
            }
            warp_sums[threadIdx.x] = wi - w;  // exclusive prefix of the warp sums
        }
        __syncthreads();
        int64_t excl = carry + warp_sums[threadIdx.x >> 5] + incl - v;
        if (i < n) counts[i] = excl;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) counts[n] = carry;
}

// write list entries in index order
__global__ void fill_kernel(const uint8_t *__restrict__ cflag, Geom g,
                            const int64_t *__restrict__ offsets, int64_t *__restrict__ lin,
                            uint8_t *__restrict__ ke, double *__restrict__ bu,
                            double *__restrict__ bv) {
    int64_t base = (int64_t)blockIdx.x * CL_CHUNK + (int64_t)threadIdx.x * CL_ITEMS;
    bool flag[CL_ITEMS];
    int c = 0;
#pragma unroll
    for (int k = 0; k < CL_ITEMS; k++) {
        flag[k] = is_listed(cflag, g, base + k);
        c += flag[k];
    }
    int incl = c;
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, o);
        if ((threadIdx.x & 31) >= o) incl += t;
    }
    __shared__ int warp_sums[CL_THREADS / 32];
    if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = incl;
    __syncthreads();
    int woff = 0;
    for (int k = 0; k < (int)(threadIdx.x >> 5); k++) woff += warp_sums[k];
    int64_t pos = offsets[blockIdx.x] + woff + incl - c;
#pragma unroll
    for (int k = 0; k < CL_ITEMS; k++) {
        if (flag[k]) {
            int64_t i = base + k;
            int64_t lx = i / g.NY, y = i - lx * g.NY;
            int64_t cidx = lx * g.pitch + y;
            uint8_t fl = cflag[cidx];
            lin[pos] = cidx;
            ke[pos] = (uint8_t)(cf_kind(fl) | (cf_edge(fl) << 3));  // listed cells are not fluid
            bu[pos] = 0.0;
            bv[pos] = 0.0;
            pos++;
        }
    }
}

// scatter the host's sparse velocity table into the list (binary search by index)
__global__ void velocity_kernel(const sb_boundary_velocity *__restrict__ tab, size_t ntab,
                                Geom g, const int64_t *__restrict__ lin,
                                const uint8_t *__restrict__ ke, uint64_t n,
                                double *__restrict__ bu, double *__restrict__ bv) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntab) return;
    int64_t lx = (int64_t)tab[t].x - g.gx0;
    if (lx < 0 || lx >= g.nxl || (int64_t)tab[t].y >= g.NY) return;
    int64_t key = lx * g.pitch + (int64_t)tab[t].y;
    uint64_t lo = 0, hi = n;
    while (lo < hi) {
        uint64_t mid = (lo + hi) >> 1;
        if (lin[mid] < key) lo = mid + 1;
        else hi = mid;
    }
    if (lo < n && lin[lo] == key) {
        int kind = ke[lo] & 7;
        if (kind == SB_KIND_INFLOW || kind == SB_KIND_MOVING_WALL) {
            bu[lo] = tab[t].u;
            bv[lo] = tab[t].v;
        }
    }
}

// All transient device memory is stream-ordered (cudaMallocAsync / cudaFreeAsync): a plain
// cudaFree synchronises the whole device, which would stall on -- and with two slabs of one
// process on one GPU deadlock against -- a peer's waiting all-gather kernel.
template <typename T>
sb_status ensure(cudaStream_t st, T *&ptr, size_t &cap, size_t need) {
    if (need <= cap) return SB_OK;
    if (ptr) SB_CUDA(cudaFreeAsync(ptr, st));
    ptr = nullptr;
    cap = 0;
    size_t ncap = need + need / 4 + 64;
    SB_CUDA(cudaMallocAsync(&ptr, ncap * sizeof(T), st));
    cap = ncap;
    return SB_OK;
}

void free_list(cudaStream_t st, BList &b) {
    void *ptrs[] = {b.lin, b.ke, b.bu, b.bv, b.ru, b.rv, b.nu, b.nv, b.wu, b.wv};
    for (void *p : ptrs)
        if (p) cudaFreeAsync(p, st);
    b = BList();
}

sb_status alloc_list(cudaStream_t st, BList &b, uint64_t n) {
    uint64_t cap = n + 16;
    SB_CUDA(cudaMallocAsync(&b.lin, cap * sizeof(int64_t), st));
    SB_CUDA(cudaMallocAsync(&b.ke, cap, st));
    double **arrs[] = {&b.bu, &b.bv, &b.ru, &b.rv, &b.nu, &b.nv, &b.wu, &b.wv};
    for (double **a : arrs) SB_CUDA(cudaMallocAsync(a, cap * sizeof(double), st));
    b.n = n;
    b.cap = cap;
    return SB_OK;
}

}  // namespace

// Classify all local cells, rebuild the list.  On SB_BOUNDARY_TOO_THIN the previous
// list and fluid count stay in force (src/grid/mod.rs:232-233); the edge bits of
// cflag are then those of the FAILED scan, exactly like the reference's Display-only
// BTreeSet, and get rewritten by the caller's rollback + re-classification.
sb_status classify(sb_sim *s) {
    const Geom &g = s->g;
    s->flag_epoch++;
    // u_v_restore is reset before the scan, so also when it fails (src/grid/mod.rs:205)
    s->restore_valid = false;
    s->uvmax_valid = false;   // the fluid set may change
    int64_t total = g.nxl * g.NY;
    unsigned long long init[2] = {~0ull, 0ull};
    SB_CUDA(cudaMemcpyAsync(s->d_err, init, sizeof(init), cudaMemcpyHostToDevice, s->stream));
    int nb_edge = (int)((total + CL_THREADS - 1) / CL_THREADS);
    edge_kernel<<<nb_edge, CL_THREADS, 0, s->stream>>>(s->cflag, g, s->d_err, s->d_err + 1);
    s->launches++;
    unsigned long long res[2];
    SB_CUDA(cudaMemcpyAsync(res, s->d_err, sizeof(res), cudaMemcpyDeviceToHost, s->stream));
    SB_CUDA(cudaStreamSynchronize(s->stream));
    uint8_t err_kind = 0;
    if (res[0] != ~0ull) {
        uint8_t fl = 0;
        int64_t ex = (int64_t)(res[0] / (unsigned long long)g.NY);
        int64_t ey = (int64_t)(res[0] % (unsigned long long)g.NY);
        SB_CUDA(cudaMemcpyAsync(&fl, s->cflag + (ex - g.gx0) * g.pitch + ey, 1,
                                cudaMemcpyDeviceToHost, s->stream));
        SB_CUDA(cudaStreamSynchronize(s->stream));
        err_kind = (uint8_t)cf_kind(fl);
    }
    if (s->slab) {
        // first offender in x-major order over ALL slabs (min of index * 8 + kind; indices
        // stay far below 2^50, exact in a double) and the global fluid count
        s->h_scalars[0] = res[0] != ~0ull ? (double)res[0] * 8.0 + (double)err_kind : 1e300;
        s->h_scalars[1] = (double)res[1];
        SB_CUDA(cudaMemcpyAsync(s->d_scalars, s->h_scalars, 2 * sizeof(double),
                                cudaMemcpyHostToDevice, s->stream));
        sb_status st2 = slab_allreduce(s, s->d_scalars, 2, XR_MIN | (XR_SUM << 2));
        if (st2) return st2;
        SB_CUDA(cudaMemcpyAsync(s->h_scalars, s->d_scalars, 2 * sizeof(double),
                                cudaMemcpyDeviceToHost, s->stream));
        SB_CUDA(cudaStreamSynchronize(s->stream));
        if ((st2 = slab_check_error(s))) return st2;
        if (s->h_scalars[0] < 1e299) {
            unsigned long long idx = (unsigned long long)(s->h_scalars[0] / 8.0);
            err_kind = (uint8_t)(s->h_scalars[0] - (double)idx * 8.0);
            res[0] = idx;
        } else {
            res[0] = ~0ull;
        }
        res[1] = (unsigned long long)s->h_scalars[1];
    }
    if (res[0] != ~0ull) {
        s->err_xy[0] = res[0] / (unsigned long long)g.NY;
        s->err_xy[1] = res[0] % (unsigned long long)g.NY;
        s->err_kind = err_kind;
        set_error("BoundaryTooThinError: cell has fluid on opposing sides");
        return SB_BOUNDARY_TOO_THIN;
    }
    int64_t nchunks = (total + CL_CHUNK - 1) / CL_CHUNK;
    sb_status st = ensure(s->stream, s->d_scan, s->scan_cap, (size_t)nchunks + 1);
    if (st) return st;
    count_kernel<<<(int)nchunks, CL_THREADS, 0, s->stream>>>(s->cflag, g, s->d_scan);
    scan_kernel<<<1, 1024, 0, s->stream>>>(s->d_scan, nchunks);
    s->launches += 2;
    int64_t nlist = 0;
    SB_CUDA(cudaMemcpyAsync(&nlist, s->d_scan + nchunks, sizeof(int64_t), cudaMemcpyDeviceToHost,
                            s->stream));
    SB_CUDA(cudaStreamSynchronize(s->stream));
    free_list(s->stream, s->bl);
    st = alloc_list(s->stream, s->bl, (uint64_t)nlist);
    if (st) return st;
    fill_kernel<<<(int)nchunks, CL_THREADS, 0, s->stream>>>(s->cflag, g, s->d_scan, s->bl.lin,
                                                            s->bl.ke, s->bl.bu, s->bl.bv);
    s->launches++;
    SB_CUDA(cudaGetLastError());
    s->fluid_cells = (double)res[1];  // owned fluid cells; slab mode sums over ranks later
    sb_status st3 = apply_velocity_table(s);
    if (st3) return st3;
    // the list is complete when classify() returns (sb_boundary_list may read it right away)
    SB_CUDA(cudaStreamSynchronize(s->stream));
    return SB_OK;
}

sb_status apply_velocity_table(sb_sim *s) {
    size_t n = s->velocities.size();
    if (s->bl.n == 0) return SB_OK;
    // the table REPLACES the velocities: cells it no longer names go back to (0, 0)
    SB_CUDA(cudaMemsetAsync(s->bl.bu, 0, s->bl.n * sizeof(double), s->stream));
    SB_CUDA(cudaMemsetAsync(s->bl.bv, 0, s->bl.n * sizeof(double), s->stream));
    if (n == 0) return SB_OK;
    sb_boundary_velocity *d_tab = nullptr;
    SB_CUDA(cudaMallocAsync(&d_tab, n * sizeof(sb_boundary_velocity), s->stream));
    SB_CUDA(cudaMemcpyAsync(d_tab, s->velocities.data(), n * sizeof(sb_boundary_velocity),
                            cudaMemcpyHostToDevice, s->stream));
    velocity_kernel<<<(int)((n + 255) / 256), 256, 0, s->stream>>>(
        d_tab, n, s->g, s->bl.lin, s->bl.ke, s->bl.n, s->bl.bu, s->bl.bv);
    s->launches++;
    SB_CUDA(cudaFreeAsync(d_tab, s->stream));
    SB_CUDA(cudaStreamSynchronize(s->stream));
    return SB_OK;
}

// force-load this file's kernels (CUDA loads lazily by default, and a first launch that has
// to load code synchronises the context -- fatal while a peer slab of the same process spins
// in an all-gather on the same GPU)
void preload_classify() {
    cudaFuncAttributes a;
    cudaFuncGetAttributes(&a, edge_kernel);
    cudaFuncGetAttributes(&a, count_kernel);
    cudaFuncGetAttributes(&a, scan_kernel);
    cudaFuncGetAttributes(&a, fill_kernel);
    cudaFuncGetAttributes(&a, velocity_kernel);
    cudaGetLastError();
}

}  // namespace sb
