// cellops.cuh -- per-cell stencil operators, strict IEEE f64 (compiled with -fmad=false).
//
// Device restatement of /root/reference/src/math.rs:19-186 and
// src/simulation.rs:349-392.  Every expression keeps the association of the Rust
// source (Rust never contracts a*b+c; powi(2) is x*x; `/` is a correctly rounded
// IEEE division), so results equal the reference's bit for bit.  The kernels and
// the C-ABI known-answer entry points (sb_du2dx, ...) share these functions.
//
// Naming: c = (i, j); e/w = (i+1, j)/(i-1, j) [x is the slow axis]; s/n = (i, j+1)/(i, j-1)
// ["north" is j-1, src/grid/mod.rs:168-176].
#pragma once

namespace sb {

// src/math.rs:19-33
__device__ __forceinline__ double du2dx(double u_w, double u_c, double u_e, double delx,
                                        double gamma) {
    double a = u_c + u_e, b = u_w + u_c;
    double left_side = (a * a) - (b * b);
    double inner_left2 = fabs(a) * (u_c - u_e);
    double inner_right2 = fabs(b) * (u_w - u_c);
    return (left_side + (gamma * (inner_left2 - inner_right2))) / (4.0 * delx);
}

// src/math.rs:53-77
__device__ __forceinline__ double duvdx(double u_c, double u_s, double u_w, double u_sw,
                                        double v_c, double v_e, double v_w, double delx,
                                        double gamma) {
    double a = u_c + u_s, b = u_w + u_sw;
    double left_side = (a * (v_c + v_e)) - (b * (v_w + v_c));
    double inner_left2 = fabs(a) * (v_c - v_e);
    double inner_right2 = fabs(b) * (v_w - v_c);
    return (left_side + (gamma * (inner_left2 - inner_right2))) / (4.0 * delx);
}

// src/math.rs:97-120
__device__ __forceinline__ double duvdy(double u_c, double u_n, double u_s, double v_c,
                                        double v_n, double v_e, double v_ne, double dely,
                                        double gamma) {
    double a = v_c + v_e, b = v_n + v_ne;
    double left_side = (a * (u_c + u_s)) - (b * (u_n + u_c));
    double inner_left2 = fabs(a) * (u_c - u_s);
    double inner_right2 = fabs(b) * (u_n - u_c);
    return (left_side + (gamma * (inner_left2 - inner_right2))) / (4.0 * dely);
}

// src/math.rs:136-150
__device__ __forceinline__ double dv2dy(double v_n, double v_c, double v_s, double dely,
                                        double gamma) {
    double a = v_c + v_s, b = v_n + v_c;
    double left_side = (a * a) - (b * b);
    double inner_left2 = fabs(a) * (v_c - v_s);
    double inner_right2 = fabs(b) * (v_n - v_c);
    return (left_side + (gamma * (inner_left2 - inner_right2))) / (4.0 * dely);
}

// src/math.rs:162-174
__device__ __forceinline__ double laplacian(double e_c, double e_n, double e_s, double e_w,
                                            double e_e, double delx, double dely) {
    double d2edx2 = ((e_e - (2. * e_c)) + e_w) / (delx * delx);
    double d2edy2 = ((e_s - (2. * e_c)) + e_n) / (dely * dely);
    return d2edx2 + d2edy2;
}

// src/math.rs:176-186
__device__ __forceinline__ double residual(double p_c, double p_n, double p_s, double p_w,
                                           double p_e, double delx, double dely, double rhs) {
    double part1 = ((p_e - p_c) - (p_c - p_w)) / (delx * delx);
    double part2 = ((p_s - p_c) - (p_c - p_n)) / (dely * dely);
    return (part1 + part2) - rhs;
}

// 3x3 neighbourhoods of u and v around (i, j); names as above (ne = (i+1, j-1) etc.)
struct Stencil9 {
    double c, n, s, w, e, nw, ne, sw, se;
};

// src/simulation.rs:349-363
__device__ __forceinline__ double calculate_f(const Stencil9 &u, const Stencil9 &v, double delx,
                                              double dely, double delt, double gamma,
                                              double reynolds) {
    return u.c + (delt * (((laplacian(u.c, u.n, u.s, u.w, u.e, delx, dely) / reynolds) -
                           du2dx(u.w, u.c, u.e, delx, gamma)) -
                          duvdy(u.c, u.n, u.s, v.c, v.n, v.e, v.ne, dely, gamma)));
}

// src/simulation.rs:378-392
__device__ __forceinline__ double calculate_g(const Stencil9 &u, const Stencil9 &v, double delx,
                                              double dely, double delt, double gamma,
                                              double reynolds) {
    return v.c + (delt * (((laplacian(v.c, v.n, v.s, v.w, v.e, delx, dely) / reynolds) -
                           duvdx(u.c, u.s, u.w, u.sw, v.c, v.e, v.w, delx, gamma)) -
                          dv2dy(v.n, v.c, v.s, dely, gamma)));
}

// view[(a, b)] == blk[3*a + b]: a indexes x (w, c, e), b indexes y (n, c, s)
__device__ __forceinline__ Stencil9 stencil_from_block(const double *blk) {
    Stencil9 s;
    s.nw = blk[0]; s.w = blk[1]; s.sw = blk[2];
    s.n = blk[3];  s.c = blk[4]; s.s = blk[5];
    s.ne = blk[6]; s.e = blk[7]; s.se = blk[8];
    return s;
}

}  // namespace sb
