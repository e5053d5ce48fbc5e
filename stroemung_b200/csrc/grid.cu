// grid.cu -- cell-mask helpers: kind upload, device-side presets, 2x2 cell edits.
//
//   presets      /root/reference/src/grid/presets.rs:8-87 (empty, simple_inflow, obstacle
//                with its integer circle rasteriser) plus the parameterised masks the
//                BASELINE configs need (channel + circle, backward-facing step, cavity)
//   edit block   draw_cells, /root/reference/src/lib.rs:38-78
#include "sb_internal.cuh"

namespace sb {

namespace {

// cflag = VALID | kind (| previous edge bits) for every cell of the grid present in this
// slab; everything else (padding columns, rows outside the grid) becomes 0 = "not a cell"
__global__ void mark_valid_kernel(Geom g, uint8_t *__restrict__ cflag,
                                  const uint8_t *__restrict__ kind_rows, int keep_edges) {
    int64_t y = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
    int64_t lx = blockIdx.x;
    if (y >= g.pitch || lx >= g.nxl) return;
    int64_t c = lx * g.pitch + y;
    int64_t gx = g.gx0 + lx;
    uint8_t out = 0;
    if (y < g.NY && gx >= 0 && gx < g.NX) {
        uint8_t old = cflag[c];
        uint8_t kind = kind_rows ? kind_rows[c] : (uint8_t)(old & 7);
        uint8_t bits = 0;
        if (keep_edges) {
            // like the reference, BCs keep using the stale classification until the list is
            // rebuilt; a cell that turns fluid gets the conservative CF_NEAR
            if ((kind & 7) == SB_KIND_FLUID) bits = CF_NEAR;
            else if (!cf_is_fluid(old)) bits = (uint8_t)(old & 0x78);
        }
        out = (uint8_t)(CF_VALID | (kind & 7) | bits);
    }
    cflag[c] = out;
}

// edge bits of boundary cells are cleared (the list kernel puts the old ones back); fluid
// cells get a conservative CF_NEAR until the next successful classification
__global__ void clear_edges_kernel(Geom g, uint8_t *__restrict__ cflag) {
    int64_t y = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
    int64_t lx = blockIdx.x;
    if (y >= g.pitch || lx >= g.nxl) return;
    uint8_t f = cflag[lx * g.pitch + y] & 0x87;
    if (f == CF_FLUID) f |= CF_NEAR;
    cflag[lx * g.pitch + y] = f;
}

__global__ void list_edges_kernel(uint8_t *__restrict__ cflag, BList bl) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= bl.n) return;
    int64_t b = bl.lin[k];
    cflag[b] = (uint8_t)((cflag[b] & 0x87) | ((bl.ke[k] >> 3) << 3));
}

// the reference's draw_circle (src/grid/presets.rs:42-62): cells with
// cx - floor(r) <= x < cx + floor(r) (saturating at 0), same in y, and sqrt(dx^2+dy^2) < r
__device__ __forceinline__ bool in_circle(int64_t x, int64_t y, int64_t cx, int64_t cy,
                                          double radius) {
    int64_t ri = (int64_t)radius;
    int64_t x_lo = cx >= ri ? cx - ri : 0, y_lo = cy >= ri ? cy - ri : 0;
    if (x < x_lo || x >= cx + ri || y < y_lo || y >= cy + ri) return false;
    int64_t dx = x - cx, dy = y - cy;
    return sqrt((double)(dx * dx + dy * dy)) < radius;
}

struct PresetArgs {
    int preset;
    int64_t a0, a1;
    double r;
};

__global__ void preset_kernel(Geom g, uint8_t *__restrict__ cflag, PresetArgs pa) {
    int64_t y = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
    int64_t lx = blockIdx.x;
    if (y >= g.pitch || lx >= g.nxl) return;
    int64_t c = lx * g.pitch + y;
    int64_t x = g.gx0 + lx;
    if (y >= g.NY || x < 0 || x >= g.NX) {
        cflag[c] = 0;
        return;
    }
    int kind = SB_KIND_FLUID;
    const int64_t NX = g.NX, NY = g.NY;
    switch (pa.preset) {
    case 0: break;  // empty: all Fluid (src/grid/presets.rs:8-17)
    case 1:         // simple_inflow (:19-40)
    case 2:         // obstacle (:64-87)
    case 3:         // channel + circle(cx, cy, r)
        if (y == 0 || y == NY - 1) kind = SB_KIND_NOSLIP;
        else if (x == 0) kind = SB_KIND_INFLOW;
        else if (x == NX - 1) kind = SB_KIND_OUTFLOW;
        if (pa.preset == 2 && in_circle(x, y, 20, NY / 2, 5.0)) kind = SB_KIND_NOSLIP;
        if (pa.preset == 3 && in_circle(x, y, pa.a0, pa.a1, pa.r)) kind = SB_KIND_NOSLIP;
        break;
    case 4:  // backward-facing step: block x < a0, y >= a1; inflow above it
        if (y == 0 || y == NY - 1) kind = SB_KIND_NOSLIP;
        else if (x < pa.a0 && y >= pa.a1) kind = SB_KIND_NOSLIP;
        else if (x == 0) kind = SB_KIND_INFLOW;
        else if (x == NX - 1) kind = SB_KIND_OUTFLOW;
        break;
    case 5:  // lid-driven cavity: NoSlip ring, moving lid on y == 0
        if (y == 0 && x >= 1 && x <= NX - 2) kind = SB_KIND_MOVING_WALL;
        else if (y == 0 || y == NY - 1 || x == 0 || x == NX - 1) kind = SB_KIND_NOSLIP;
        break;
    }
    cflag[c] = (uint8_t)(CF_VALID | kind);
}

// draw_cells (src/lib.rs:38-78): the 2x2 block at (gx, gy); interior cells whose kind
// differs get u = v = p = 0 and the new kind.  backup: 4 x {u, v, p, kind, touched}.
__global__ void edit_block_kernel(Geom g, uint8_t *cflag, double *u, double *v, double *p,
                                  int64_t gx, int64_t gy, uint8_t kind, double *backup,
                                  int restore, int32_t *modified) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int mod = 0;
    for (int k = 0; k < 4; k++) {
        // order of src/lib.rs:42-47: (x,y), (x+1,y), (x,y+1), (x+1,y+1)
        int64_t x = gx + (k & 1), y = gy + (k >> 1);
        double *bk = backup + 5 * k;
        if (restore) {
            if (bk[4] != 0.0) {
                int64_t lx = x - g.gx0;
                if (lx < 0 || lx >= g.nxl) continue;
                int64_t c = lx * g.pitch + y;
                u[c] = bk[0]; v[c] = bk[1]; p[c] = bk[2];
                cflag[c] = (uint8_t)((cflag[c] & 0xF8) | ((int)bk[3] & 7));
            }
            continue;
        }
        bk[4] = 0.0;
        if (!(x > 0 && x < g.NX - 1 && y > 0 && y < g.NY - 1)) continue;
        int64_t lx = x - g.gx0;
        if (lx < 0 || lx >= g.nxl) continue;
        int64_t c = lx * g.pitch + y;
        if (cf_kind(cflag[c]) != kind) {
            bk[0] = u[c]; bk[1] = v[c]; bk[2] = p[c]; bk[3] = (double)cf_kind(cflag[c]);
            bk[4] = 1.0;
            u[c] = 0.0; v[c] = 0.0; p[c] = 0.0;
            cflag[c] = (uint8_t)((cflag[c] & 0xF8) | kind);
            mod = 1;
        }
    }
    if (!restore) *modified = mod;
}

dim3 cell_grid(const Geom &g) { return dim3((unsigned)g.nxl, (unsigned)((g.pitch + 255) / 256)); }

}  // namespace

sb_status launch_mark_valid(sb_sim *s, const uint8_t *d_kind_rows, int keep_edges) {
    s->flag_epoch++;
    mark_valid_kernel<<<cell_grid(s->g), 256, 0, s->stream>>>(s->g, s->cflag, d_kind_rows,
                                                              keep_edges);
    s->launches++;
    SB_CUDA(cudaGetLastError());
    return SB_OK;
}

sb_status restore_edges_from_list(sb_sim *s) {
    s->flag_epoch++;
    clear_edges_kernel<<<cell_grid(s->g), 256, 0, s->stream>>>(s->g, s->cflag);
    s->launches++;
    if (s->bl.n) {
        list_edges_kernel<<<(int)((s->bl.n + 255) / 256), 256, 0, s->stream>>>(s->cflag, s->bl);
        s->launches++;
    }
    SB_CUDA(cudaGetLastError());
    return SB_OK;
}

sb_status launch_preset(sb_sim *s, int preset, const double *args) {
    s->flag_epoch++;
    PresetArgs pa;
    pa.preset = preset;
    pa.a0 = (int64_t)args[0];
    pa.a1 = (int64_t)args[1];
    pa.r = args[2];
    preset_kernel<<<cell_grid(s->g), 256, 0, s->stream>>>(s->g, s->cflag, pa);
    s->launches++;
    SB_CUDA(cudaGetLastError());
    return SB_OK;
}

sb_status launch_edit_block(sb_sim *s, int64_t gx, int64_t gy, uint8_t kind, double *backup,
                            int restore, int32_t *modified) {
    s->flag_epoch++;
    edit_block_kernel<<<1, 32, 0, s->stream>>>(s->g, s->cflag, s->u, s->v, s->p[s->cur], gx, gy,
                                               kind, backup, restore, modified);
    s->launches++;
    SB_CUDA(cudaGetLastError());
    return SB_OK;
}

// force-load this file's kernels (CUDA loads lazily by default, and a first launch that has
// to load code synchronises the context -- fatal while a peer slab of the same process spins
// in an all-gather on the same GPU)
void preload_grid() {
    cudaFuncAttributes a;
    cudaFuncGetAttributes(&a, mark_valid_kernel);
    cudaFuncGetAttributes(&a, clear_edges_kernel);
    cudaFuncGetAttributes(&a, list_edges_kernel);
    cudaFuncGetAttributes(&a, preset_kernel);
    cudaFuncGetAttributes(&a, edit_block_kernel);
    cudaGetLastError();
}

}  // namespace sb
