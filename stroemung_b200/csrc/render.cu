// render.cu -- N4: the colour mapping that follows the tick in the reference's frame loop.
//
//   render_simulation   /root/reference/src/visualization.rs:79-105
//   color_pressure      :49-70      color_speed  :29-47      hsl_to_rgb  :7-27
//
// The reference walks the grid on the host every frame and sets one macroquad pixel per
// cell, which forces a download of p (or u and v) and the cell types.  Here one kernel
// reads the fields where they live and writes the RGBA8 image, so a frame costs 4 bytes per
// cell over PCIe instead of 9 or 17.
//
// Arithmetic follows the Rust source operation for operation: the hue in f64, cast to f32,
// hsl_to_rgb in f32 (`%` is fmodf), then macroquad 0.4.13's `From<Color> for [u8; 4]`
// (`(c * 255.) as u8`, a saturating cast, NaN -> 0; the crate is not vendored in the
// reference tree -- Cargo.lock:284-285).  Image layout is macroquad's `Image`: row-major
// with width nx, pixel (x, y) at byte 4 (y nx + x) -- the transpose of the field arrays,
// done through a shared-memory tile so that both sides are coalesced.
#include "sb_internal.cuh"

namespace sb {

namespace {

// Rust `as u8` on an f32
__device__ __forceinline__ uint8_t sat_u8(float v) {
    if (!(v > 0.0f)) return 0;  // negatives, -0, NaN
    if (v >= 255.0f) return 255;
    return (uint8_t)v;
}

__device__ __forceinline__ uchar4 hue_to_rgba(float hue) {
    const float saturation = 1.0f, lightness = 0.5f;
    const float c = (1.0f - fabsf(2.0f * lightness - 1.0f)) * saturation;
    const float x = c * (1.0f - fabsf(fmodf(hue / 60.0f, 2.0f) - 1.0f));
    const float m = lightness - c / 2.0f;
    float r, g, b;
    if (hue < 60.0f) { r = c; g = x; b = 0.0f; }
    else if (hue < 120.0f) { r = x; g = c; b = 0.0f; }
    else if (hue < 180.0f) { r = 0.0f; g = c; b = x; }
    else if (hue < 240.0f) { r = 0.0f; g = x; b = c; }
    else if (hue < 300.0f) { r = x; g = 0.0f; b = c; }
    else { r = c; g = 0.0f; b = x; }
    return make_uchar4(sat_u8((r + m) * 255.0f), sat_u8((g + m) * 255.0f),
                       sat_u8((b + m) * 255.0f), sat_u8(1.0f * 255.0f));
}

// one 32 x 32 tile of cells per block of 32 x 8 threads
template <int SPEED>
__global__ void __launch_bounds__(256)
render_kernel(Geom g, const uint8_t *__restrict__ cflag, const double *__restrict__ a,
              const double *__restrict__ b, double lo, double hi, uchar4 *__restrict__ img,
              int64_t width) {
    __shared__ uchar4 tile[32][33];
    const int64_t y0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
    const int64_t rows = g.own1 - g.own0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int64_t r = r0 + threadIdx.y + 8 * k, y = y0 + threadIdx.x;
        uchar4 px = make_uchar4(0, 0, 0, 0);
        if (r < rows && y < g.NY) {
            const int64_t c = (g.own0 + r) * g.pitch + y;
            if (cf_is_fluid(cflag[c])) {
                double q;
                if (SPEED) {
                    const double uu = a[c], vv = b[c];
                    q = sqrt((uu * uu) + (vv * vv));
                } else {
                    q = a[c];
                }
                // 240 offset: blue to red instead of the whole hue circle
                const float hue = (float)(240.0 - (((q - lo) * 240.0) / (hi - lo)));
                px = hue_to_rgba(hue);
            } else {
                px = SPEED ? make_uchar4(127, 127, 127, 255) : make_uchar4(127, 0, 0, 255);
            }
        }
        tile[threadIdx.y + 8 * k][threadIdx.x] = px;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int64_t y = y0 + threadIdx.y + 8 * k, r = r0 + threadIdx.x;
        if (r < rows && y < g.NY) img[y * width + r] = tile[threadIdx.x][threadIdx.y + 8 * k];
    }
}

}  // namespace

// d_img: device buffer of (owned rows) x NY pixels
sb_status launch_render(sb_sim *s, int speed, uchar4 *d_img) {
    const Geom &g = s->g;
    const int64_t rows = g.own1 - g.own0;
    dim3 grid((unsigned)((g.NY + 31) / 32), (unsigned)((rows + 31) / 32)), block(32, 8);
    if (speed)
        render_kernel<1><<<grid, block, 0, s->stream>>>(g, s->cflag, s->u, s->v, s->speed_range[0],
                                                        s->speed_range[1], d_img, rows);
    else
        render_kernel<0><<<grid, block, 0, s->stream>>>(g, s->cflag, s->p[s->cur], nullptr,
                                                        s->pressure_range[0], s->pressure_range[1],
                                                        d_img, rows);
    s->launches++;
    SB_CUDA(cudaGetLastError());
    return SB_OK;
}

void preload_render() {
    cudaFuncAttributes a;
    cudaFuncGetAttributes(&a, render_kernel<0>);
    cudaFuncGetAttributes(&a, render_kernel<1>);
    cudaGetLastError();
}

}  // namespace sb
