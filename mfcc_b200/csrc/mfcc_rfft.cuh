// mfcc_rfft.cuh — register-resident REAL-input DFT codelets (16 and 32 points, with the
// zero-padded tail pruned at compile time) and the complex 4/8/16-point DFTs that the
// two-pass real FFT of the fused kernel is made of (mfcc_fused_ct.cu):
//
//   N = RB * RA real points, n = a + RA b, k = k1 + RB k2
//   pass 1: for each a, real DFT-RB over b          -> Y[k1][a], k1 = 0 .. RB/2 (Hermitian half)
//           times W_N^(a k1)
//   pass 2: for each k1, complex DFT-RA over a      -> X[k1 + RB k2]
//
// No packing of two real points into one complex point, hence no split step after the
// transform.  Everything is statically indexed so that arrays stay in registers.
// The functions are host+device so that tests/codelets (g++) can check them against a
// direct double-precision DFT without a GPU.
// No reference code corresponds to this (SURVEY.md §8a "Ref file:line = none").
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define MFCC_HD __host__ __device__ __forceinline__
#else
#define MFCC_HD inline
#endif

namespace mfcc {
namespace rf {

struct cplx { float re, im; };

constexpr float kH = 0.70710678118654752f;    // cos(pi/4)
constexpr float kC1 = 0.92387953251128674f;   // cos(pi/8)
constexpr float kS1 = 0.38268343236508977f;   // sin(pi/8)

// a * (cr + i ci): 2 FMUL + 2 FFMA
MFCC_HD cplx cmulc(cplx a, float cr, float ci)
{
    cplx r;
    r.re = fmaf(-a.im, ci, a.re * cr);
    r.im = fmaf(a.im, cr, a.re * ci);
    return r;
}
MFCC_HD cplx conj(cplx a) { return cplx{a.re, -a.im}; }

// Forward 4-point DFT (W4 = -i), natural order, in place.
MFCC_HD void cdft4(cplx &x0, cplx &x1, cplx &x2, cplx &x3)
{
    const cplx t0{x0.re + x2.re, x0.im + x2.im}, t1{x0.re - x2.re, x0.im - x2.im};
    const cplx t2{x1.re + x3.re, x1.im + x3.im}, t3{x1.re - x3.re, x1.im - x3.im};
    x0 = cplx{t0.re + t2.re, t0.im + t2.im};
    x2 = cplx{t0.re - t2.re, t0.im - t2.im};
    x1 = cplx{t1.re + t3.im, t1.im - t3.re};
    x3 = cplx{t1.re - t3.im, t1.im + t3.re};
}

// Outputs 0 and 1 only of the 4-point DFT.
MFCC_HD void cdft4_01(cplx a0, cplx a1, cplx a2, cplx a3, cplx &y0, cplx &y1)
{
    const cplx s02{a0.re + a2.re, a0.im + a2.im}, d02{a0.re - a2.re, a0.im - a2.im};
    const cplx s13{a1.re + a3.re, a1.im + a3.im}, d13{a1.re - a3.re, a1.im - a3.im};
    y0 = cplx{s02.re + s13.re, s02.im + s13.im};
    y1 = cplx{d02.re + d13.im, d02.im - d13.re};
}

// ---- FMA-fused twiddled butterflies (MFCC_RFFT_FUSED) ----
// (a + w b, a - w b) for a COMPILE-TIME twiddle w = c + i s.  Multiplying first (2 FMUL + 2 FFMA) and adding after
// (4 FADD) takes 8 instructions; factoring the larger of c, s out of w b leaves u = b (1 + i s/c) (2 FFMA) and four
// FFMAs a +- c u: 6 instructions.  The ratio is folded by the compiler (the arguments are literals at every call
// site); rounding is that of the product form to within one ulp of the factored-out component.
#ifndef MFCC_RFFT_FUSED
#define MFCC_RFFT_FUSED 0
#endif
MFCC_HD void bfly_tw(cplx a, cplx b, float c, float s, cplx &sum, cplx &dif)
{
    if (c == 0.0f && s == -1.0f) {             // w = -i: w b = (b.im, -b.re), no multiply at all
        sum = cplx{a.re + b.im, a.im - b.re};
        dif = cplx{a.re - b.im, a.im + b.re};
    } else if (fabsf(c) >= fabsf(s)) {
        const float r = s / c;                 // w b = c (b.re - r b.im, b.im + r b.re)
        const float ur = fmaf(-r, b.im, b.re), ui = fmaf(r, b.re, b.im);
        sum = cplx{fmaf(c, ur, a.re), fmaf(c, ui, a.im)};
        dif = cplx{fmaf(-c, ur, a.re), fmaf(-c, ui, a.im)};
    } else {
        const float r = c / s;                 // w b = s (r b.re - b.im, r b.im + b.re)
        const float ur = fmaf(r, b.re, -b.im), ui = fmaf(r, b.im, b.re);
        sum = cplx{fmaf(s, ur, a.re), fmaf(s, ui, a.im)};
        dif = cplx{fmaf(-s, ur, a.re), fmaf(-s, ui, a.im)};
    }
}
// 4-point DFT of (x0, w1 x1, w2 x2, w3 x3), in place: 4 + 6 + 6 + 8 = 24 instructions against 12 + 16.
MFCC_HD void cdft4_tw(cplx &x0, cplx &x1, cplx &x2, cplx &x3, float c1, float s1, float c2, float s2, float c3, float s3)
{
    cplx t0, t1, t2, t3;
    bfly_tw(x0, x2, c2, s2, t0, t1);
    bfly_tw(cmulc(x1, c1, s1), x3, c3, s3, t2, t3);
    x0 = cplx{t0.re + t2.re, t0.im + t2.im};
    x2 = cplx{t0.re - t2.re, t0.im - t2.im};
    x1 = cplx{t1.re + t3.im, t1.im - t3.re};
    x3 = cplx{t1.re - t3.im, t1.im + t3.re};
}
// Outputs 0 and 1 only of the 4-point DFT of (a0, w1 a1, w2 a2, w3 a3).
MFCC_HD void cdft4_01_tw(cplx a0, cplx a1, cplx a2, cplx a3, float c1, float s1, float c2, float s2, float c3, float s3,
                         cplx &y0, cplx &y1)
{
    cplx s02, d02, s13, d13;
    bfly_tw(a0, a2, c2, s2, s02, d02);
    bfly_tw(cmulc(a1, c1, s1), a3, c3, s3, s13, d13);
    y0 = cplx{s02.re + s13.re, s02.im + s13.im};
    y1 = cplx{d02.re + d13.im, d02.im - d13.re};
}

// Last stage of the 8-point DFT: x[2 ka + 1] *= W8^ka, then radix-2 over the pairs; natural order out.
MFCC_HD void cdft8_last(cplx (&x)[8])
{
    cplx r[8];
#if MFCC_RFFT_FUSED
    r[0] = cplx{x[0].re + x[1].re, x[0].im + x[1].im};
    r[4] = cplx{x[0].re - x[1].re, x[0].im - x[1].im};
    bfly_tw(x[2], x[3], kH, -kH, r[1], r[5]);          // W8^1
    bfly_tw(x[4], x[5], 0.0f, -1.0f, r[2], r[6]);      // W8^2 = -i
    bfly_tw(x[6], x[7], -kH, -kH, r[3], r[7]);         // W8^3
#else
    x[3] = cmulc(x[3], kH, -kH);             // W8^1
    x[5] = cplx{x[5].im, -x[5].re};          // W8^2 = -i
    x[7] = cmulc(x[7], -kH, -kH);            // W8^3
#pragma unroll
    for (int ka = 0; ka < 4; ++ka) {
        r[ka] = cplx{x[2 * ka].re + x[2 * ka + 1].re, x[2 * ka].im + x[2 * ka + 1].im};
        r[ka + 4] = cplx{x[2 * ka].re - x[2 * ka + 1].re, x[2 * ka].im - x[2 * ka + 1].im};
    }
#endif
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = r[i];
}

// Forward 8-point DFT, natural order in and out: n = nb + 2 na, k = ka + 4 kb.
MFCC_HD void cdft8(cplx (&x)[8])
{
    cdft4(x[0], x[2], x[4], x[6]);
    cdft4(x[1], x[3], x[5], x[7]);
    cdft8_last(x);
}

// Forward 16-point DFT, natural order in and out: n = nb + 4 na, k = ka + 4 kb.
MFCC_HD void cdft16(cplx (&x)[16])
{
#pragma unroll
    for (int nb = 0; nb < 4; ++nb) cdft4(x[nb], x[nb + 4], x[nb + 8], x[nb + 12]);
    // y[nb][ka] sits in x[nb + 4 ka]; multiply by W16^(nb ka)
#if MFCC_RFFT_FUSED
    cdft4(x[0], x[1], x[2], x[3]);
    cdft4_tw(x[4], x[5], x[6], x[7], kC1, -kS1, kH, -kH, kS1, -kC1);
    cdft4_tw(x[8], x[9], x[10], x[11], kH, -kH, 0.0f, -1.0f, -kH, -kH);
    cdft4_tw(x[12], x[13], x[14], x[15], kS1, -kC1, -kH, -kH, -kC1, kS1);
#else
    x[1 + 4 * 1] = cmulc(x[1 + 4 * 1], kC1, -kS1);
    x[1 + 4 * 2] = cmulc(x[1 + 4 * 2], kH, -kH);
    x[1 + 4 * 3] = cmulc(x[1 + 4 * 3], kS1, -kC1);
    x[2 + 4 * 1] = cmulc(x[2 + 4 * 1], kH, -kH);
    x[2 + 4 * 2] = cplx{x[2 + 4 * 2].im, -x[2 + 4 * 2].re};
    x[2 + 4 * 3] = cmulc(x[2 + 4 * 3], -kH, -kH);
    x[3 + 4 * 1] = cmulc(x[3 + 4 * 1], kS1, -kC1);
    x[3 + 4 * 2] = cmulc(x[3 + 4 * 2], -kH, -kH);
    x[3 + 4 * 3] = cmulc(x[3 + 4 * 3], -kC1, kS1);
#pragma unroll
    for (int ka = 0; ka < 4; ++ka) cdft4(x[4 * ka], x[4 * ka + 1], x[4 * ka + 2], x[4 * ka + 3]);
#endif
    // X[ka + 4 kb] sits in x[4 ka + kb]: transpose to natural order (register renaming)
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = a + 1; b < 4; ++b) {
            const cplx t = x[4 * a + b];
            x[4 * a + b] = x[4 * b + a];
            x[4 * b + a] = t;
        }
}

// First stage of the real transforms: 4-point DFT over m of the REAL values x_m with the
// last 4 - C of them known to be zero.  t0 = sum, t2 = alternating sum, t1 = bin 1 (bin 3 = conj).
template <int C>
MFCC_HD void rstage(float x0, float x1, float x2, float x3, float &t0, float &t2, cplx &t1)
{
    if constexpr (C >= 4) {
        const float s02 = x0 + x2, d02 = x0 - x2, s13 = x1 + x3, d13 = x1 - x3;
        t0 = s02 + s13; t2 = s02 - s13; t1 = cplx{d02, -d13};
    } else if constexpr (C == 3) {
        const float s02 = x0 + x2, d02 = x0 - x2;
        t0 = s02 + x1; t2 = s02 - x1; t1 = cplx{d02, -x1};
    } else if constexpr (C == 2) {
        t0 = x0 + x1; t2 = x0 - x1; t1 = cplx{x0, -x1};
    } else if constexpr (C == 1) {
        t0 = x0; t2 = x0; t1 = cplx{x0, 0.0f};
    } else {
        t0 = 0.0f; t2 = 0.0f; t1 = cplx{0.0f, 0.0f};
    }
}
// how many of x[j], x[j + M], x[j + 2M], x[j + 3M] lie below NZ
constexpr int live4(int j, int M, int NZ) { return NZ <= j ? 0 : (NZ - j + M - 1) / M > 4 ? 4 : (NZ - j + M - 1) / M; }

// Real 8-point DFT: X[0], X[4] real, X[1..3] complex.  20 operations.
MFCC_HD void rdft8(const float (&t)[8], float &x0, cplx &x1, cplx &x2, cplx &x3, float &x4)
{
    const float a0 = t[0] + t[4], a1 = t[0] - t[4], b0 = t[2] + t[6], b1 = t[2] - t[6];
    const float c0 = t[1] + t[5], c1 = t[1] - t[5], d0 = t[3] + t[7], d1 = t[3] - t[7];
    const float e0 = a0 + b0, e1 = a0 - b0, f0 = c0 + d0, f1 = c0 - d0;
    x0 = e0 + f0;
    x4 = e0 - f0;
    x2 = cplx{e1, -f1};
    const float g = c1 - d1, s = c1 + d1;
    x1 = cplx{fmaf(kH, g, a1), -fmaf(kH, s, b1)};
    x3 = cplx{fmaf(-kH, g, a1), fmaf(-kH, s, b1)};
}

// Real 16-point DFT of x[0 .. NZ) (x[NZ .. 16) taken as zero): X[0 .. 8], X[0].im = X[8].im = 0.
template <int NZ>
MFCC_HD void rdft16(const float (&x)[16], cplx (&X)[9])
{
    float t0[4], t2[4];
    cplx u[4];
    rstage<live4(0, 4, NZ)>(x[0], x[4], x[8], x[12], t0[0], t2[0], u[0]);
    rstage<live4(1, 4, NZ)>(x[1], x[5], x[9], x[13], t0[1], t2[1], u[1]);
    rstage<live4(2, 4, NZ)>(x[2], x[6], x[10], x[14], t0[2], t2[2], u[2]);
    rstage<live4(3, 4, NZ)>(x[3], x[7], x[11], x[15], t0[3], t2[3], u[3]);
    {   // q = 0: real 4-point DFT of t0 -> X[0], X[4], X[8]
        const float s0 = t0[0] + t0[2], d0 = t0[0] - t0[2], s1 = t0[1] + t0[3], d1 = t0[1] - t0[3];
        X[0] = cplx{s0 + s1, 0.0f};
        X[8] = cplx{s0 - s1, 0.0f};
        X[4] = cplx{d0, -d1};
    }
    {   // q = 2: sum_j t2[j] W8^j W4^(j r) -> X[2], X[6]
        const float g = t2[1] - t2[3], s = t2[1] + t2[3];
        X[2] = cplx{fmaf(kH, g, t2[0]), -fmaf(kH, s, t2[2])};
        X[6] = cplx{fmaf(-kH, g, t2[0]), fmaf(-kH, s, t2[2])};
    }
    // q = 1: (u[j] W16^j) through a complex 4-point DFT -> X[1], X[5], X[9] = conj X[7], X[13] = conj X[3]
#if MFCC_RFFT_FUSED
    cdft4_tw(u[0], u[1], u[2], u[3], kC1, -kS1, kH, -kH, kS1, -kC1);
#else
    u[1] = cmulc(u[1], kC1, -kS1);
    u[2] = cmulc(u[2], kH, -kH);
    u[3] = cmulc(u[3], kS1, -kC1);
    cdft4(u[0], u[1], u[2], u[3]);
#endif
    X[1] = u[0];
    X[5] = u[1];
    X[7] = conj(u[2]);
    X[3] = conj(u[3]);
}

// Real 32-point DFT of x[0 .. NZ) (the rest zero): X[0 .. 16], X[0].im = X[16].im = 0.
// tw32[j] = W32^j = exp(-2 pi i j / 32) for j = 1 .. 7 are passed in (compile-time literals at the call site).
template <int NZ>
MFCC_HD void rdft32(const float (&x)[32], cplx (&X)[17])
{
    float t0[8], t2[8];
    cplx u[8];
    rstage<live4(0, 8, NZ)>(x[0], x[8], x[16], x[24], t0[0], t2[0], u[0]);
    rstage<live4(1, 8, NZ)>(x[1], x[9], x[17], x[25], t0[1], t2[1], u[1]);
    rstage<live4(2, 8, NZ)>(x[2], x[10], x[18], x[26], t0[2], t2[2], u[2]);
    rstage<live4(3, 8, NZ)>(x[3], x[11], x[19], x[27], t0[3], t2[3], u[3]);
    rstage<live4(4, 8, NZ)>(x[4], x[12], x[20], x[28], t0[4], t2[4], u[4]);
    rstage<live4(5, 8, NZ)>(x[5], x[13], x[21], x[29], t0[5], t2[5], u[5]);
    rstage<live4(6, 8, NZ)>(x[6], x[14], x[22], x[30], t0[6], t2[6], u[6]);
    rstage<live4(7, 8, NZ)>(x[7], x[15], x[23], x[31], t0[7], t2[7], u[7]);
    {   // q = 0: real 8-point DFT of t0 -> X[0], X[4], X[8], X[12], X[16]
        float x0, x4;
        rdft8(t0, x0, X[4], X[8], X[12], x4);
        X[0] = cplx{x0, 0.0f};
        X[16] = cplx{x4, 0.0f};
    }
    {   // q = 2: X[2 + 4 r] = sum_{j<4} W16^(j (1 + 2 r)) (t2[j] - i (-1)^r t2[j + 4])
        // r even: outputs 0, 1 of the 4-point DFT of W16^j (t2[j], -t2[j+4])  -> X[2], X[10]
        const cplx a0{t2[0], -t2[4]};
        // r odd: outputs 0, 1 of the 4-point DFT of W16^(3 j) (t2[j], +t2[j+4]) -> X[6], X[14]
        const cplx b0{t2[0], t2[4]};
#if MFCC_RFFT_FUSED
        cdft4_01_tw(a0, cplx{t2[1], -t2[5]}, cplx{t2[2], -t2[6]}, cplx{t2[3], -t2[7]}, kC1, -kS1, kH, -kH, kS1, -kC1, X[2], X[10]);
        cdft4_01_tw(b0, cplx{t2[1], t2[5]}, cplx{t2[2], t2[6]}, cplx{t2[3], t2[7]}, kS1, -kC1, -kH, -kH, -kC1, kS1, X[6], X[14]);
#else
        const cplx a1 = cmulc(cplx{t2[1], -t2[5]}, kC1, -kS1);
        const cplx a2 = cmulc(cplx{t2[2], -t2[6]}, kH, -kH);
        const cplx a3 = cmulc(cplx{t2[3], -t2[7]}, kS1, -kC1);
        cdft4_01(a0, a1, a2, a3, X[2], X[10]);
        const cplx b1 = cmulc(cplx{t2[1], t2[5]}, kS1, -kC1);    // W16^3
        const cplx b2 = cmulc(cplx{t2[2], t2[6]}, -kH, -kH);     // W16^6
        const cplx b3 = cmulc(cplx{t2[3], t2[7]}, -kC1, kS1);    // W16^9
        cdft4_01(b0, b1, b2, b3, X[6], X[14]);
#endif
    }
    // q = 1: (u[j] W32^j) through a complex 8-point DFT -> X[1 + 4 r]; r >= 4 gives the conjugates of X[15], X[11], X[7], X[3]
#if MFCC_RFFT_FUSED
    // the W32^j twiddles ride on the first radix-4 stage of the 8-point DFT (even j on u0, u2, u4, u6; odd j on
    // u1 W32^1 first, then u3, u5, u7)
    cdft4_tw(u[0], u[2], u[4], u[6], kC1, -kS1, kH, -kH, kS1, -kC1);
    u[1] = cmulc(u[1], 0.98078528040323044f, -0.19509032201612827f);
    cdft4_tw(u[1], u[3], u[5], u[7], 0.83146961230254524f, -0.55557023301960222f, 0.55557023301960222f, -0.83146961230254524f,
             0.19509032201612827f, -0.98078528040323044f);
    cdft8_last(u);
#else
    u[1] = cmulc(u[1], 0.98078528040323044f, -0.19509032201612827f);
    u[2] = cmulc(u[2], kC1, -kS1);
    u[3] = cmulc(u[3], 0.83146961230254524f, -0.55557023301960222f);
    u[4] = cmulc(u[4], kH, -kH);
    u[5] = cmulc(u[5], 0.55557023301960222f, -0.83146961230254524f);
    u[6] = cmulc(u[6], kS1, -kC1);
    u[7] = cmulc(u[7], 0.19509032201612827f, -0.98078528040323044f);
    cdft8(u);
#endif
    X[1] = u[0];
    X[5] = u[1];
    X[9] = u[2];
    X[13] = u[3];
    X[15] = conj(u[4]);
    X[11] = conj(u[5]);
    X[7] = conj(u[6]);
    X[3] = conj(u[7]);
}


// W32^j = exp(-2 pi i j / 32), j = 0 .. 21 (largest product nb * ka in cdft32), and W64^j, j = 0 .. 15.
struct w2 { float re, im; };
MFCC_HD w2 w32(int j)
{
    constexpr w2 t[22] = {
    {1.0f, 0.0f},
    {0.98078528040323043f, -0.19509032201612825f},
    {0.92387953251128674f, -0.38268343236508978f},
    {0.83146961230254524f, -0.55557023301960218f},
    {0.70710678118654757f, -0.70710678118654746f},
    {0.55557023301960229f, -0.83146961230254524f},
    {0.38268343236508984f, -0.92387953251128674f},
    {0.19509032201612833f, -0.98078528040323043f},
    {0.0f, -1.0f},
    {-0.19509032201612819f, -0.98078528040323043f},
    {-0.38268343236508973f, -0.92387953251128674f},
    {-0.55557023301960196f, -0.83146961230254546f},
    {-0.70710678118654746f, -0.70710678118654757f},
    {-0.83146961230254535f, -0.55557023301960218f},
    {-0.92387953251128674f, -0.38268343236508989f},
    {-0.98078528040323043f, -0.19509032201612861f},
    {-1.0f, 0.0f},
    {-0.98078528040323043f, 0.19509032201612836f},
    {-0.92387953251128685f, 0.38268343236508967f},
    {-0.83146961230254546f, 0.55557023301960196f},
    {-0.70710678118654768f, 0.70710678118654746f},
    {-0.55557023301960218f, 0.83146961230254524f}};
    return t[j];
}
MFCC_HD w2 w64(int j)
{
    constexpr w2 t[16] = {
    {1.0f, 0.0f},
    {0.99518472667219693f, -0.098017140329560604f},
    {0.98078528040323043f, -0.19509032201612825f},
    {0.95694033573220882f, -0.29028467725446233f},
    {0.92387953251128674f, -0.38268343236508978f},
    {0.88192126434835505f, -0.47139673682599764f},
    {0.83146961230254524f, -0.55557023301960218f},
    {0.77301045336273699f, -0.63439328416364549f},
    {0.70710678118654757f, -0.70710678118654746f},
    {0.63439328416364549f, -0.77301045336273699f},
    {0.55557023301960229f, -0.83146961230254524f},
    {0.47139673682599781f, -0.88192126434835494f},
    {0.38268343236508984f, -0.92387953251128674f},
    {0.29028467725446233f, -0.95694033573220894f},
    {0.19509032201612833f, -0.98078528040323043f},
    {0.09801714032956077f, -0.99518472667219682f}};
    return t[j];
}

// Forward 32-point DFT, natural order in and out: n = nb + 4 na (na < 8), k = ka + 8 kb (kb < 4).
MFCC_HD void cdft32(cplx (&x)[32])
{
    cplx t[4][8];
#pragma unroll
    for (int nb = 0; nb < 4; ++nb) {
#pragma unroll
        for (int na = 0; na < 8; ++na) t[nb][na] = x[nb + 4 * na];
        cdft8(t[nb]);
    }
#pragma unroll
    for (int nb = 1; nb < 4; ++nb)
#pragma unroll
        for (int ka = 1; ka < 8; ++ka) {
            const w2 w = w32(nb * ka);
            t[nb][ka] = cmulc(t[nb][ka], w.re, w.im);
        }
#pragma unroll
    for (int ka = 0; ka < 8; ++ka) {
        cdft4(t[0][ka], t[1][ka], t[2][ka], t[3][ka]);
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) x[ka + 8 * kb] = t[kb][ka];
    }
}

// Real 64-point DFT of x[0 .. NZ) (the rest zero): X[0 .. 32], X[0].im = X[32].im = 0.
// k = q + 4 r: bin q of the 4-point DFT over m of x[j + 16 m], times W64^(j q), through a 16-point DFT over j.
template <int NZ>
MFCC_HD void rdft64(const float (&x)[64], cplx (&X)[33])
{
    float t0[16], t2[16];
    cplx u[16];
#define MFCC_RS(j) rstage<live4(j, 16, NZ)>(x[j], x[j + 16], x[j + 32], x[j + 48], t0[j], t2[j], u[j])
    MFCC_RS(0); MFCC_RS(1); MFCC_RS(2); MFCC_RS(3); MFCC_RS(4); MFCC_RS(5); MFCC_RS(6); MFCC_RS(7);
    MFCC_RS(8); MFCC_RS(9); MFCC_RS(10); MFCC_RS(11); MFCC_RS(12); MFCC_RS(13); MFCC_RS(14); MFCC_RS(15);
#undef MFCC_RS
    {   // q = 0: real 16-point DFT of t0 -> X[4 r], r = 0 .. 8
        cplx X0[9];
        rdft16<16>(t0, X0);
#pragma unroll
        for (int r = 0; r <= 8; ++r) X[4 * r] = X0[r];
        X[0].im = 0.0f;
        X[32].im = 0.0f;
    }
    {   // q = 2: (t2[j] W32^j) through a complex 16-point DFT -> X[2 + 4 r], r = 0 .. 7
        cplx z[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const w2 w = w32(j);
            z[j] = cplx{t2[j] * w.re, t2[j] * w.im};
        }
        cdft16(z);
#pragma unroll
        for (int r = 0; r < 8; ++r) X[2 + 4 * r] = z[r];
    }
    {   // q = 1: (u[j] W64^j) through a complex 16-point DFT -> X[1 + 4 r]; r >= 8 gives the conjugates of X[3 + 4 r']
#pragma unroll
        for (int j = 1; j < 16; ++j) {
            const w2 w = w64(j);
            u[j] = cmulc(u[j], w.re, w.im);
        }
        cdft16(u);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            X[1 + 4 * r] = u[r];
            X[3 + 4 * r] = conj(u[15 - r]);
        }
    }
}

template <int RB> struct RDft;
template <> struct RDft<16> {
    template <int NZ> static MFCC_HD void run(const float (&x)[16], cplx (&X)[9]) { rdft16<NZ>(x, X); }
};
template <> struct RDft<32> {
    template <int NZ> static MFCC_HD void run(const float (&x)[32], cplx (&X)[17]) { rdft32<NZ>(x, X); }
};
template <> struct RDft<64> {
    template <int NZ> static MFCC_HD void run(const float (&x)[64], cplx (&X)[33]) { rdft64<NZ>(x, X); }
};

}  // namespace rf
}  // namespace mfcc
