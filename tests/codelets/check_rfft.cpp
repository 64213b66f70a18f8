// check_rfft.cpp — host check of mfcc_b200/csrc/mfcc_rfft.cuh (built and run by
// tests/test_codelets.py, no GPU): every codelet against a direct double DFT, and the
// kernel's two-pass index algebra (N = RB * RA, n = a + RA b, k = k1 + RB k2) end to end.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../mfcc_b200/csrc/mfcc_rfft.cuh"

using namespace mfcc::rf;

static double frand() { return (double)rand() / RAND_MAX * 2.0 - 1.0; }
static double g_worst = 0.0;
static int g_fail = 0;

static void expect(const char *what, double got_re, double got_im, double ref_re, double ref_im, double scale)
{
    const double e = std::hypot(got_re - ref_re, got_im - ref_im) / scale;
    if (e > g_worst) g_worst = e;
    if (!(e < 2e-6)) { ++g_fail; if (g_fail < 10) std::printf("FAIL %s: got (%g,%g) ref (%g,%g)\n", what, got_re, got_im, ref_re, ref_im); }
}

template <int N, int NZ, typename F>
static void check_real(const char *name, F run)
{
    for (int rep = 0; rep < 20; ++rep) {
        float x[N];
        for (int i = 0; i < N; ++i) x[i] = i < NZ ? (float)frand() : 1e30f;   // poison the pruned tail
        cplx X[N / 2 + 1];
        run(x, X);
        for (int k = 0; k <= N / 2; ++k) {
            double re = 0, im = 0;
            for (int n = 0; n < NZ; ++n) { re += x[n] * std::cos(2 * M_PI * k * n / N); im -= x[n] * std::sin(2 * M_PI * k * n / N); }
            expect(name, X[k].re, X[k].im, re, im, std::sqrt((double)N));
        }
    }
}

template <int N, typename F>
static void check_cplx(const char *name, F run)
{
    for (int rep = 0; rep < 20; ++rep) {
        cplx x[N], y[N];
        for (int i = 0; i < N; ++i) { x[i] = cplx{(float)frand(), (float)frand()}; y[i] = x[i]; }
        run(y);
        for (int k = 0; k < N; ++k) {
            double re = 0, im = 0;
            for (int n = 0; n < N; ++n) {
                const double c = std::cos(2 * M_PI * k * n / N), s = -std::sin(2 * M_PI * k * n / N);
                re += x[n].re * c - x[n].im * s;
                im += x[n].re * s + x[n].im * c;
            }
            expect(name, y[k].re, y[k].im, re, im, std::sqrt((double)N));
        }
    }
}

// The fused kernel's data flow in plain loops (same index maps, same row packing).
template <int RB, int RA, int L>
static void check_two_pass(const char *name)
{
    constexpr int N = RB * RA, NZ = (L + RA - 1) / RA, H = RB / 2;
    for (int rep = 0; rep < 5; ++rep) {
        std::vector<float> y(N, 0.0f);
        for (int i = 0; i < L; ++i) y[i] = (float)frand();
        // pass 1
        std::vector<cplx> ws((H - 1) * RA);      // rows k1 = 1 .. H-1
        std::vector<cplx> sp(RA);                // (row 0, row H) both real before the twiddle
        for (int a = 0; a < RA; ++a) {
            float x[RB];
            for (int b = 0; b < RB; ++b) x[b] = b < NZ ? y[a + RA * b] : 1e30f;
            cplx X[H + 1];
            RDft<RB>::template run<NZ>(x, X);
            sp[a] = cplx{X[0].re, X[H].re};
            for (int k1 = 1; k1 < H; ++k1) {
                const double ang = -2 * M_PI * a * k1 / N;
                ws[(k1 - 1) * RA + a] = cmulc(X[k1], (float)std::cos(ang), (float)std::sin(ang));
            }
        }
        // pass 2
        std::vector<double> P(N / 2 + 1, -1.0);
        std::vector<cplx> Xo(N / 2 + 1);
        for (int k1 = 1; k1 < H; ++k1) {
            cplx z[16];
            for (int a = 0; a < RA; ++a) z[a] = ws[(k1 - 1) * RA + a];
            cdft16(z);
            for (int k2 = 0; k2 < RA; ++k2) {
                const int k = k1 + RB * k2;
                if (k2 < RA / 2) Xo[k] = z[k2]; else Xo[N - k] = conj(z[k2]);
            }
        }
        {
            float r0[16];
            cplx zh[16];
            for (int a = 0; a < RA; ++a) {
                r0[a] = sp[a].re;
                const double ang = -2 * M_PI * a / (2 * RA);
                zh[a] = cplx{sp[a].im * (float)std::cos(ang), sp[a].im * (float)std::sin(ang)};
            }
            cplx X0[9];
            rdft16<16>(r0, X0);
            for (int k2 = 0; k2 <= RA / 2; ++k2) Xo[RB * k2] = X0[k2];
            cdft16(zh);
            for (int k2 = 0; k2 < RA / 2; ++k2) Xo[H + RB * k2] = zh[k2];
        }
        for (int k = 0; k <= N / 2; ++k) {
            double re = 0, im = 0;
            for (int n = 0; n < L; ++n) { re += y[n] * std::cos(2 * M_PI * k * n / N); im -= y[n] * std::sin(2 * M_PI * k * n / N); }
            expect(name, Xo[k].re, Xo[k].im, re, im, std::sqrt((double)N));
        }
    }
}

// The wide kernel's data flow (mfcc_fused_wide.cu): N = 64 * 32, rows k1 = 1 .. 32 complex (row 32 is real
// before its twiddle), row 0 real through a real DFT-32.
template <int L>
static void check_two_pass_wide(const char *name)
{
    constexpr int RB = 64, RA = 32, N = RB * RA, NZ = (L + RA - 1) / RA, H = RB / 2;
    for (int rep = 0; rep < 3; ++rep) {
        std::vector<float> y(N, 0.0f);
        for (int i = 0; i < L; ++i) y[i] = (float)frand();
        std::vector<cplx> ws(H * RA);            // rows k1 = 1 .. H at index k1 - 1
        std::vector<float> r0(RA);
        for (int a = 0; a < RA; ++a) {
            float x[RB];
            for (int b = 0; b < RB; ++b) x[b] = b < NZ ? y[a + RA * b] : 1e30f;
            cplx X[H + 1];
            rdft64<NZ>(x, X);
            r0[a] = X[0].re;
            for (int k1 = 1; k1 <= H; ++k1) {
                const double ang = -2 * M_PI * a * k1 / N;
                ws[(k1 - 1) * RA + a] = cmulc(X[k1], (float)std::cos(ang), (float)std::sin(ang));
            }
        }
        std::vector<cplx> Xo(N / 2 + 1);
        for (int k1 = 1; k1 <= H; ++k1) {
            cplx z[32];
            for (int a = 0; a < RA; ++a) z[a] = ws[(k1 - 1) * RA + a];
            cdft32(z);
            for (int k2 = 0; k2 < RA; ++k2) {
                const int k = k1 + RB * k2;
                if (k2 < RA / 2) Xo[k] = z[k2];
                else if (k1 < H) Xo[N - k] = conj(z[k2]);
            }
        }
        {
            float x[32];
            for (int a = 0; a < RA; ++a) x[a] = r0[a];
            cplx X0[17];
            rdft32<32>(x, X0);
            for (int k2 = 0; k2 <= RA / 2; ++k2) Xo[RB * k2] = X0[k2];
        }
        for (int k = 0; k <= N / 2; ++k) {
            double re = 0, im = 0;
            for (int n = 0; n < L; ++n) { re += y[n] * std::cos(2 * M_PI * k * n / N); im -= y[n] * std::sin(2 * M_PI * k * n / N); }
            expect(name, Xo[k].re, Xo[k].im, re, im, std::sqrt((double)N));
        }
    }
}

int main()
{
    srand(12345);
    check_cplx<8>("cdft8", [](cplx (&v)[8]) { cdft8(v); });
    check_cplx<16>("cdft16", [](cplx (&v)[16]) { cdft16(v); });
    check_real<16, 16>("rdft16<16>", [](const float (&x)[16], cplx (&X)[9]) { rdft16<16>(x, X); });
    check_real<16, 13>("rdft16<13>", [](const float (&x)[16], cplx (&X)[9]) { rdft16<13>(x, X); });
    check_real<16, 7>("rdft16<7>", [](const float (&x)[16], cplx (&X)[9]) { rdft16<7>(x, X); });
    check_real<32, 32>("rdft32<32>", [](const float (&x)[32], cplx (&X)[17]) { rdft32<32>(x, X); });
    check_real<32, 25>("rdft32<25>", [](const float (&x)[32], cplx (&X)[17]) { rdft32<25>(x, X); });
    check_real<32, 19>("rdft32<19>", [](const float (&x)[32], cplx (&X)[17]) { rdft32<19>(x, X); });
    check_cplx<32>("cdft32", [](cplx (&v)[32]) { cdft32(v); });
    check_real<64, 64>("rdft64<64>", [](const float (&x)[64], cplx (&X)[33]) { rdft64<64>(x, X); });
    check_real<64, 38>("rdft64<38>", [](const float (&x)[64], cplx (&X)[33]) { rdft64<38>(x, X); });
    check_real<64, 17>("rdft64<17>", [](const float (&x)[64], cplx (&X)[33]) { rdft64<17>(x, X); });
    check_two_pass_wide<1200>("two-pass 2048 (L=1200)");
    check_two_pass_wide<2048>("two-pass 2048 (L=2048)");
    check_two_pass<32, 16, 400>("two-pass 512 (L=400)");
    check_two_pass<16, 16, 200>("two-pass 256 (L=200)");
    check_two_pass<32, 16, 512>("two-pass 512 (L=512)");
    std::printf("worst normalised error %.3g, failures %d\n", g_worst, g_fail);
    return g_fail ? 1 : 0;
}
