/*
 * fastdiv_check.c -- TEST INFRASTRUCTURE: brute-force check of the constant-divisor division
 * used by the strict CUDA kernels (stroemung_b200/csrc/cellops.cuh, DivC / DivF).
 *
 * The reference divides with IEEE `/` (src/math.rs:32,76,119,149,170-171,
 * src/simulation.rs:210-212,360,389).  The kernels compute, for a run-constant divisor d
 * with r = RN(1/d):   q0 = a*r; e = fma(-d,q0,a); q1 = fma(e,r,q0); e = fma(-d,q1,a);
 *                     q  = fma(e,r,q1)        (Markstein's correction steps)
 * and claim q == a/d bit for bit whenever 2^-450 <= |a| <= 2^450.  This program compares the
 * two on the divisors of every BASELINE configuration plus random ones, over random and
 * adversarial dividends (near-multiples of d, i.e. quotients close to rounding midpoints are
 * generated from random quotients multiplied back).  Prints the mismatch count; exit 1 if any.
 *
 *     gcc -O2 -ffp-contract=off -o fastdiv_check fastdiv_check.c -lm && ./fastdiv_check [iters]
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static uint64_t s = 0x5EED5EEDull;
static uint64_t rnd(void) {
    s += 0x9E3779B97F4A7C15ull;
    uint64_t z = s;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static double fastdiv(double a, double d, double r) {
    double q = a * r;
    double e = fma(-d, q, a);
    q = fma(e, r, q);
    e = fma(-d, q, a);
    q = fma(e, r, q);
    return q == 0.0 ? a * r : q;
}
static double from_bits(uint64_t b) { double x; memcpy(&x, &b, 8); return x; }

int main(int argc, char **argv) {
    long per = argc > 1 ? atol(argv[1]) : 4000;
    const double dx[] = {0.1, 0.2, 1.0 / 1022, 4.1 / 2046, 7.5 / 4094, 10.0 / 8192, 10.0 / 32768};
    const double misc[] = {100.0, 1000.0, 400.0, 200.0, 0.005, 2e-4, 1e-4, 3.0, 7.0,
                           1.0000000000000002, 1.9999999999999996, 0.3333333333333333};
    double ds[64];
    int nd = 0;
    for (unsigned i = 0; i < sizeof(dx) / sizeof(dx[0]); i++) {
        ds[nd++] = dx[i]; ds[nd++] = dx[i] * dx[i]; ds[nd++] = 4.0 * dx[i];
    }
    for (unsigned i = 0; i < sizeof(misc) / sizeof(misc[0]); i++) ds[nd++] = misc[i];
    long bad = 0, n = 0;
    for (int k = 0; k < nd + 1500; k++) {
        double d;
        if (k < nd) d = ds[k];
        else {
            d = from_bits((rnd() & 0xFFFFFFFFFFFFFull) | ((uint64_t)(1023 - 60 + (rnd() % 120)) << 52));
            if (rnd() & 1) d = -d;
        }
        const double r = 1.0 / d;
        for (long i = 0; i < per; i++) {
            double a;
            switch (i % 4) {
            case 0:  /* random dividend anywhere in the safe window */
                a = from_bits((rnd() & 0xFFFFFFFFFFFFFull) |
                              ((uint64_t)(1023 - 450 + (rnd() % 900)) << 52) | ((rnd() & 1) << 63));
                break;
            case 1:  /* small integer multiples of d: exact or nearly exact quotients */
                a = d * (double)(1 + rnd() % 1000000);
                break;
            case 2: { /* random quotient multiplied back: quotients near representable values */
                double q = from_bits((rnd() & 0xFFFFFFFFFFFFFull) | ((uint64_t)(1023 - 30 + (rnd() % 60)) << 52));
                a = q * d;
                break;
            }
            default: { /* quotient near a rounding midpoint: (q + ulp/2) * d */
                double q = from_bits((rnd() & 0xFFFFFFFFFFFFFull) | ((uint64_t)1023 << 52));
                a = fma(q, d, d * 0x1p-53);
                break;
            }
            }
            if (a == 0.0) continue;
            n++;
            if (fastdiv(a, d, r) != a / d) bad++;
        }
    }
    /* signed zeros */
    if (signbit(fastdiv(-0.0, 0.1, 1.0 / 0.1)) != signbit(-0.0 / 0.1)) bad++;
    if (signbit(fastdiv(0.0, -0.1, 1.0 / -0.1)) != signbit(0.0 / -0.1)) bad++;
    printf("checked %ld quotients, %ld mismatches\n", n, bad);
    return bad ? 1 : 0;
}
