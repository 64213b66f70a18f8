// mfcc_fft.cuh — register-resident small DFTs and the real-FFT split used by the
// fused tile kernels.  Everything is statically indexed so arrays stay in registers.
// No reference code corresponds to this (SURVEY.md §8a "Ref file:line = none").
#pragma once
#include <cuda_runtime.h>

namespace mfcc {
namespace {

struct cplx { float re, im; };

__device__ __forceinline__ cplx cmulc(cplx a, float cr, float ci)
{
    cplx r;
    r.re = fmaf(-a.im, ci, a.re * cr);
    r.im = fmaf(a.im, cr, a.re * ci);
    return r;
}

// Forward 4-point DFT (W4 = -i), in place on four named values.
__device__ __forceinline__ void dft4(cplx &x0, cplx &x1, cplx &x2, cplx &x3)
{
    const cplx t0{x0.re + x2.re, x0.im + x2.im}, t1{x0.re - x2.re, x0.im - x2.im};
    const cplx t2{x1.re + x3.re, x1.im + x3.im}, t3{x1.re - x3.re, x1.im - x3.im};
    x0 = cplx{t0.re + t2.re, t0.im + t2.im};
    x2 = cplx{t0.re - t2.re, t0.im - t2.im};
    x1 = cplx{t1.re + t3.im, t1.im - t3.re};
    x3 = cplx{t1.re - t3.im, t1.im + t3.re};
}

// dft4 with x3 == 0 (zero-padded tail of the frame): two complex adds fewer.
__device__ __forceinline__ void dft4_z3(cplx &x0, cplx &x1, cplx &x2, cplx &x3)
{
    const cplx t0{x0.re + x2.re, x0.im + x2.im}, t1{x0.re - x2.re, x0.im - x2.im};
    const cplx a = x1;
    x0 = cplx{t0.re + a.re, t0.im + a.im};
    x2 = cplx{t0.re - a.re, t0.im - a.im};
    x1 = cplx{t1.re + a.im, t1.im - a.re};
    x3 = cplx{t1.re - a.im, t1.im + a.re};
}

// Forward 16-point DFT, natural order in and out, everything statically indexed.
// n = nb + 4 na, k = ka + 4 kb:  4-point DFTs over na, twiddle W16^(nb ka), 4-point DFTs over nb.
// NZ = number of leading inputs that can be non-zero (13 when x[13..15] are zero padding).
template <int NZ = 16>
__device__ __forceinline__ void dft16(cplx (&x)[16])
{
    constexpr float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, h = 0.70710678118654752f;
#pragma unroll
    for (int nb = 0; nb < 4; ++nb) {
        if (nb + 12 >= NZ) dft4_z3(x[nb], x[nb + 4], x[nb + 8], x[nb + 12]);
        else               dft4(x[nb], x[nb + 4], x[nb + 8], x[nb + 12]);
    }
    // after this, x[nb + 4 ka] holds y[nb][ka]
    x[1 + 4 * 1] = cmulc(x[1 + 4 * 1], c1, -s1);   // W^1
    x[1 + 4 * 2] = cmulc(x[1 + 4 * 2], h, -h);     // W^2
    x[1 + 4 * 3] = cmulc(x[1 + 4 * 3], s1, -c1);   // W^3
    x[2 + 4 * 1] = cmulc(x[2 + 4 * 1], h, -h);     // W^2
    x[2 + 4 * 2] = cplx{x[2 + 4 * 2].im, -x[2 + 4 * 2].re};  // W^4 = -i
    x[2 + 4 * 3] = cmulc(x[2 + 4 * 3], -h, -h);    // W^6
    x[3 + 4 * 1] = cmulc(x[3 + 4 * 1], s1, -c1);   // W^3
    x[3 + 4 * 2] = cmulc(x[3 + 4 * 2], -h, -h);    // W^6
    x[3 + 4 * 3] = cmulc(x[3 + 4 * 3], -c1, s1);   // W^9
#pragma unroll
    for (int ka = 0; ka < 4; ++ka) dft4(x[4 * ka], x[4 * ka + 1], x[4 * ka + 2], x[4 * ka + 3]);
    // now x[4 ka + kb] holds X[ka + 4 kb]; transpose the 4x4 index to natural order
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = a + 1; b < 4; ++b) {
            const cplx t = x[4 * a + b];
            x[4 * a + b] = x[4 * b + a];
            x[4 * b + a] = t;
        }
}

// Forward 8-point DFT, natural order: n = nb + 2 na (na<4), k = ka + 4 kb (ka<4, kb<2).
__device__ __forceinline__ void dft8(cplx (&x)[8])
{
    constexpr float h = 0.70710678118654752f;
    dft4(x[0], x[2], x[4], x[6]);  // nb = 0: y[0][ka] in x[2 ka]
    dft4(x[1], x[3], x[5], x[7]);  // nb = 1: y[1][ka] in x[2 ka + 1]
    x[3] = cmulc(x[3], h, -h);               // W8^1
    x[5] = cplx{x[5].im, -x[5].re};          // W8^2 = -i
    x[7] = cmulc(x[7], -h, -h);              // W8^3
    cplx r[8];
#pragma unroll
    for (int ka = 0; ka < 4; ++ka) {
        r[ka] = cplx{x[2 * ka].re + x[2 * ka + 1].re, x[2 * ka].im + x[2 * ka + 1].im};
        r[ka + 4] = cplx{x[2 * ka].re - x[2 * ka + 1].re, x[2 * ka].im - x[2 * ka + 1].im};
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = r[i];
}

template <int R> struct Dft;
template <> struct Dft<16> { static __device__ __forceinline__ void run(cplx (&x)[16]) { dft16<16>(x); } };
template <> struct Dft<8>  { static __device__ __forceinline__ void run(cplx (&x)[8]) { dft8(x); } };

// Real-FFT split of one (k, N/2-k) pair of the packed transform followed by the
// power spectrum.  zk = Z[k], zm = Z[N/2-k], w = exp(-2 pi i k / N).
// Returns |2 X[k]|^2 and |2 X[N/2-k]|^2 (caller scales by 1/(4N)).
__device__ __forceinline__ void split_power(cplx zk, cplx zm, float2 w, float &pk, float &pm)
{
    const float er = zk.re + zm.re, ei = zk.im - zm.im;   // Z[k] + conj(Z[m])
    const float p = zk.im + zm.im, q = zm.re - zk.re;     // -i (Z[k] - conj(Z[m]))
    const float tr = fmaf(-w.y, q, w.x * p), ti = fmaf(w.y, p, w.x * q);
    const float ar = er + tr, ai = ei + ti, br = er - tr, bi = ei - ti;
    pk = fmaf(ar, ar, ai * ai);
    pm = fmaf(br, br, bi * bi);
}


}  // namespace
}  // namespace mfcc
