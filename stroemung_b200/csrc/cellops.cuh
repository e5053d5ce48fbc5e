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

#include <math.h>

namespace sb {

// ---- division by a per-launch constant ----------------------------------------------------
// Every `/` of the reference on this path divides by a run constant (dx*dx, 4*dx, Re, dt ...).
// nvcc's correctly rounded f64 division is a ~25-instruction routine, most of it spent on the
// reciprocal; with the divisor fixed the reciprocal is hoisted to the host and the quotient
// is finished with Markstein's correction steps:
//     q0 = a*r;  e0 = fma(-d, q0, a);  q1 = fma(e0, r, q0)        (q1 is a faithful quotient)
//                e1 = fma(-d, q1, a);  q  = fma(e1, r, q1)        (= RN(a/d))
// With r = RN(1/d) and q1 faithful the last step yields the correctly rounded quotient
// (Markstein 1990; Muller et al., Handbook of Floating-Point Arithmetic, thm. 4.8), i.e. the
// SAME BITS as the IEEE division of the Rust source.  The proof needs: no over/underflow in
// the intermediates and the significand of d not all ones.  make_divc() checks d on the host,
// div() checks a per call and falls back to the true division outside the safe exponent
// window (and returns signed zeros directly).  oracle/fastdiv_check.c brute-forces the
// equality on the CPU; the GPU parity tests compare every field bit for bit with the oracle,
// which uses the plain `/`.
struct DivC {
    double d;   // the divisor
    double r;   // RN(1/d)
    int fast;   // correction-step path allowed for this divisor
    __device__ __forceinline__ double operator()(double a) const {
        if (fast) {
            const double aa = fabs(a);
            if (aa >= 0x1p-450 && aa <= 0x1p450) {
                double q = a * r;
                double e = fma(-d, q, a);
                q = fma(e, r, q);
                e = fma(-d, q, a);
                return fma(e, r, q);
            }
            if (a == 0.0) return d > 0.0 ? a : -a;
        }
        return a / d;
    }
};

// The same correction steps without a branch per division: the operand check (exponent
// inside the safe window, or an exact zero) is accumulated into *bad with integer
// instructions and the caller redoes the whole cell with plain divisions if it ever fires.
struct DivF {
    double d, r;
    unsigned *bad;
    __device__ __forceinline__ double operator()(double a) const {
        // operand check on the high word: exponent field in [573, 1473] (2^-450 .. 2^450),
        // or an exact zero
        const unsigned hi = (unsigned)__double2hiint(a);
        const unsigned ha = hi & 0x7fffffffu;
        const bool inwin = (ha - (573u << 20)) < (901u << 20);
        const bool zero = (ha | (unsigned)__double2loint(a)) == 0u;
        *bad |= (inwin || zero) ? 0u : 1u;
        const double q0 = a * r;
        double er = fma(-d, q0, a);
        double q = fma(er, r, q0);
        er = fma(-d, q, a);
        q = fma(er, r, q);
        // sign from q0 = a*r: right for every quotient, and the fma chain would turn -0 into +0
        return __hiloint2double((__double2hiint(q) & 0x7fffffff) | (__double2hiint(q0) & 0x80000000),
                                __double2loint(q));
    }
};

// plain IEEE division (known-answer entry points, one-off uses)
struct DivT {
    double d;
    __device__ __forceinline__ double operator()(double a) const { return a / d; }
};

inline DivC make_divc(double d) {
    DivC c;
    c.d = d;
    c.r = 1.0 / d;
    int ex = 0;
    const double m = frexp(fabs(d), &ex);  // m in [0.5, 1)
    c.fast = isfinite(d) && d != 0.0 && ex > -400 && ex < 400 && m != 0x1.fffffffffffffp-1;
    return c;
}

// src/math.rs:19-33
template <class D4>
__device__ __forceinline__ double du2dx(double u_w, double u_c, double u_e, const D4 &four_delx,
                                        double gamma) {
    double a = u_c + u_e, b = u_w + u_c;
    double left_side = (a * a) - (b * b);
    double inner_left2 = fabs(a) * (u_c - u_e);
    double inner_right2 = fabs(b) * (u_w - u_c);
    return four_delx(left_side + (gamma * (inner_left2 - inner_right2)));
}

// src/math.rs:53-77
template <class D4>
__device__ __forceinline__ double duvdx(double u_c, double u_s, double u_w, double u_sw,
                                        double v_c, double v_e, double v_w, const D4 &four_delx,
                                        double gamma) {
    double a = u_c + u_s, b = u_w + u_sw;
    double left_side = (a * (v_c + v_e)) - (b * (v_w + v_c));
    double inner_left2 = fabs(a) * (v_c - v_e);
    double inner_right2 = fabs(b) * (v_w - v_c);
    return four_delx(left_side + (gamma * (inner_left2 - inner_right2)));
}

// src/math.rs:97-120
template <class D4>
__device__ __forceinline__ double duvdy(double u_c, double u_n, double u_s, double v_c,
                                        double v_n, double v_e, double v_ne, const D4 &four_dely,
                                        double gamma) {
    double a = v_c + v_e, b = v_n + v_ne;
    double left_side = (a * (u_c + u_s)) - (b * (u_n + u_c));
    double inner_left2 = fabs(a) * (u_c - u_s);
    double inner_right2 = fabs(b) * (u_n - u_c);
    return four_dely(left_side + (gamma * (inner_left2 - inner_right2)));
}

// src/math.rs:136-150
template <class D4>
__device__ __forceinline__ double dv2dy(double v_n, double v_c, double v_s, const D4 &four_dely,
                                        double gamma) {
    double a = v_c + v_s, b = v_n + v_c;
    double left_side = (a * a) - (b * b);
    double inner_left2 = fabs(a) * (v_c - v_s);
    double inner_right2 = fabs(b) * (v_n - v_c);
    return four_dely(left_side + (gamma * (inner_left2 - inner_right2)));
}

// src/math.rs:162-174; dx2 / dy2 divide by (delx * delx) / (dely * dely)
template <class D2>
__device__ __forceinline__ double laplacian(double e_c, double e_n, double e_s, double e_w,
                                            double e_e, const D2 &dx2, const D2 &dy2) {
    double d2edx2 = dx2((e_e - (2. * e_c)) + e_w);
    double d2edy2 = dy2((e_s - (2. * e_c)) + e_n);
    return d2edx2 + d2edy2;
}

// src/math.rs:176-186
template <class D2>
__device__ __forceinline__ double residual(double p_c, double p_n, double p_s, double p_w,
                                           double p_e, const D2 &dx2, const D2 &dy2, double rhs) {
    double part1 = dx2((p_e - p_c) - (p_c - p_w));
    double part2 = dy2((p_s - p_c) - (p_c - p_n));
    return (part1 + part2) - rhs;
}

// 3x3 neighbourhoods of u and v around (i, j); names as above (ne = (i+1, j-1) etc.)
struct Stencil9 {
    double c, n, s, w, e, nw, ne, sw, se;
};

// the run constants every F / G evaluation divides by
template <class D>
struct FgDiv {
    D dx2, dy2;          // delx * delx, dely * dely   (laplacian)
    D four_dx, four_dy;  // 4.0 * delx, 4.0 * dely     (du2dx, duvdx / duvdy, dv2dy)
    D re;                // reynolds
};

template <class D>
__host__ __device__ inline FgDiv<D> fg_div_plain(double delx, double dely, double reynolds) {
    FgDiv<D> k;
    k.dx2.d = delx * delx; k.dy2.d = dely * dely;
    k.four_dx.d = 4.0 * delx; k.four_dy.d = 4.0 * dely;
    k.re.d = reynolds;
    return k;
}

// src/simulation.rs:349-363
template <class D>
__device__ __forceinline__ double calculate_f(const Stencil9 &u, const Stencil9 &v,
                                              const FgDiv<D> &k, double delt, double gamma) {
    return u.c + (delt * ((k.re(laplacian(u.c, u.n, u.s, u.w, u.e, k.dx2, k.dy2)) -
                           du2dx(u.w, u.c, u.e, k.four_dx, gamma)) -
                          duvdy(u.c, u.n, u.s, v.c, v.n, v.e, v.ne, k.four_dy, gamma)));
}

// src/simulation.rs:378-392
template <class D>
__device__ __forceinline__ double calculate_g(const Stencil9 &u, const Stencil9 &v,
                                              const FgDiv<D> &k, double delt, double gamma) {
    return v.c + (delt * ((k.re(laplacian(v.c, v.n, v.s, v.w, v.e, k.dx2, k.dy2)) -
                           duvdx(u.c, u.s, u.w, u.sw, v.c, v.e, v.w, k.four_dx, gamma)) -
                          dv2dy(v.n, v.c, v.s, k.four_dy, gamma)));
}

// ---- performance mode (SB_SOR_RED_BLACK; extension, not in the reference) -----------------------
// The same formulas with every division replaced by a multiplication with a reciprocal
// computed once per run and a*b+c taken fused.  The tree below is restated operand for
// operand in oracle/stroemung_oracle.c (fg_cell_fast, so_calculate_rhs): bit-identical to it,
// last-bit differences against the strict operators above (covered, like the red-black
// ordering, by the performance-mode tolerance).  ~70 FP64 instructions per cell for F, G and
// rhs instead of ~630 with the 13 exact divisions.
struct FgFast {
    double rdx2, rdy2, r4dx, r4dy, rre, rdx, rdy, rdt, gamma, delt;
};

inline FgFast make_fg_fast(double delx, double dely, double delt, double gamma, double reynolds) {
    FgFast k;
    k.rdx2 = 1.0 / (delx * delx);
    k.rdy2 = 1.0 / (dely * dely);
    k.r4dx = 1.0 / (4.0 * delx);
    k.r4dy = 1.0 / (4.0 * dely);
    k.rre = 1.0 / reynolds;
    k.rdx = 1.0 / delx;
    k.rdy = 1.0 / dely;
    k.rdt = 1.0 / delt;
    k.gamma = gamma;
    k.delt = delt;
    return k;
}

__device__ __forceinline__ double lap_fast(const FgFast &k, double c, double n, double s, double w,
                                           double e) {
    const double tc = 2.0 * c;
    return fma(k.rdx2, (e - tc) + w, k.rdy2 * ((s - tc) + n));
}

// numerator of a donor-cell term: (a*pa - b*pb) + gamma*(|a|*da - |b|*db)
__device__ __forceinline__ double donor_fast(const FgFast &k, double a, double pa, double b,
                                             double pb, double da, double db) {
    const double left = fma(a, pa, -(b * pb));
    const double d = fma(fabs(a), da, -(fabs(b) * db));
    return fma(k.gamma, d, left);
}

__device__ __forceinline__ double calculate_f_fast(const Stencil9 &u, const Stencil9 &v,
                                                   const FgFast &k) {
    const double a1 = u.c + u.e, b1 = u.w + u.c;
    const double x_f = donor_fast(k, a1, a1, b1, b1, u.c - u.e, u.w - u.c);            // du2dx
    const double a2 = v.c + v.e, b2 = v.n + v.ne;
    const double y_f = donor_fast(k, a2, u.c + u.s, b2, u.n + u.c, u.c - u.s, u.n - u.c);  // duvdy
    const double lu = lap_fast(k, u.c, u.n, u.s, u.w, u.e);
    return fma(k.delt, fma(-k.r4dy, y_f, fma(-k.r4dx, x_f, k.rre * lu)), u.c);
}

__device__ __forceinline__ double calculate_g_fast(const Stencil9 &u, const Stencil9 &v,
                                                   const FgFast &k) {
    const double a3 = u.c + u.s, b3 = u.w + u.sw;
    const double x_g = donor_fast(k, a3, v.c + v.e, b3, v.w + v.c, v.c - v.e, v.w - v.c);  // duvdx
    const double a4 = v.c + v.s, b4 = v.n + v.c;
    const double y_g = donor_fast(k, a4, a4, b4, b4, v.c - v.s, v.n - v.c);            // dv2dy
    const double lv = lap_fast(k, v.c, v.n, v.s, v.w, v.e);
    return fma(k.delt, fma(-k.r4dy, y_g, fma(-k.r4dx, x_g, k.rre * lv)), v.c);
}

__device__ __forceinline__ double rhs_fast(const FgFast &k, double f_c, double f_w, double g_c,
                                           double g_n) {
    return k.rdt * fma(k.rdx, f_c - f_w, k.rdy * (g_c - g_n));
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
