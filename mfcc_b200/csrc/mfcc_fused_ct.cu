// mfcc_fused_ct.cu — fused tile kernel with COMPILE-TIME geometry (frame length,
// hop, radix split), persistent CTAs and all tables in shared memory.
//
// Why this shape (numbers from tools/microbench*.cu on the B200):
//   * hot code must stay under ~32 KB: at 16 warps/SM the FFMA rate falls from
//     107 to 50 (77 KB) and 23 lanes/clk/SM (154 KB) once the instruction stream
//     outgrows the instruction caches -> no per-warp role specialisation, every
//     warp runs the same compact code and picks its butterfly by index;
//   * register-indexed constant-bank loads (LDC) manage only 0.37 warp-loads/clk/SM
//     -> per-butterfly constants are read from a shared-memory copy of the tables
//     with uniform-address LDS.128 (0.58 warp-loads/clk/SM, 4 constants each);
//   * I2F runs on the 16-lane XU pipe -> int16 samples are converted with the
//     2^23 mantissa trick (one LOP3 + one FADD);
//   * integer division by a runtime hop costs ~20 instructions -> frame length
//     and hop are template parameters.
//
// Phases per tile (lane = frame, 8 warps, CTA barriers between phases):
//   S0 stage, S1 pass 1 (window, REAL DFT-RB over b for a pair of columns a, inter-pass
//   twiddle), S2 pass 2 (complex DFT-RA over a for one row k1) + power, S3 sparse mel
//   (segment form, 4-bin chunks), S4 log + symmetric/antisymmetric halves, S5 DCT on the
//   halves, coalesced store.
//
// The transform is a two-pass REAL FFT, N = RB * RA, n = a + RA b, k = k1 + RB k2
// (mfcc_rfft.cuh): pass 1 keeps only the Hermitian half k1 = 0 .. RB/2 of each column,
// so the workspace holds N/2 complex words per frame and pass 2 delivers bins directly —
// there is no "two reals in one complex" packing and therefore no split step.
//
// No reference code corresponds to this (SURVEY.md §8a "Ref file:line = none").
#include <cuda_runtime.h>

#include <cmath>
#include <cstring>
#include <vector>

#include "mfcc_rfft.cuh"
#include "mfcc_host.h"

namespace mfcc {

namespace {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int kPad = 2;

// Offsets (in floats) of the tables inside the shared-memory blob; set by the host.
struct CtLayout {
    int win;      // [RA/2][NZP] float2: window of samples (a, a + 1) + RA b, a = 2 pair
    int tw;       // [RA][RB/2] float2: W_N^(a k1) at slot k1 - 1, k1 = 1 .. RB/2 - 1 (last slot unused)
    int twh;      // [RA] float2: W_(2 RA)^a, the twiddle of row k1 = RB/2
    int seg;      // int4 per segment: {first bin, chunks, weight offset (floats from melw), 0}
    int seg_lo;   // int[kWarps + 1]: segment range of each warp
    int melw;     // per chunk: 4 rise weights then 4 fall weights (scaled by 1/(4 NFFT))
    int dct;      // [n_cep][halfp] floats: DCT row restricted to m < ceil(M/2), zero padded to halfp
    int total;    // floats, multiple of 4
    int n_seg, half, halfp;
};

struct CtArgs {
    const Tile *tiles;
    int64_t n_tiles;
    int64_t pcm_len;
    float *out;
    const float *tab;   // global copy of the table blob
    CtLayout lay;
    int n_mel, n_cep, logmel;
    float preemph, log_floor;
};

template <int L, int HOP, int RB, int RA>
struct Geo {
    static constexpr int NFFT = RB * RA, NB = NFFT / 2 + 1, H = RB / 2;
    static constexpr int NZ = (L + RA - 1) / RA;           // rows b of a column that carry samples
    static constexpr int NZP = (NZ + 1) / 2 * 2;
    static constexpr int STRIDE = HOP + kPad;              // staged words per hop block
    static constexpr int padded(int i) { return i + kPad * (i / HOP); }
    static constexpr int SLACK = RA * NZ - L;              // words past the last frame read with zero window
    static constexpr int MAXS = (L + 1 > RA * NZ ? L + 1 : RA * NZ);
    static constexpr int STAGED = 31 * STRIDE + padded(MAXS) + 2;
    static constexpr int PW = (NB + 3) * 32;               // 3 slack rows for 4-bin chunks
    static constexpr int UNION = ((STAGED > PW ? STAGED : PW) + 3) / 4 * 4;
    static constexpr int WS = H * RA * 32 * 2;             // floats: rows 1 .. H-1 complex, rows 0 and H (real) share the last
    static_assert(HOP % 2 == 0, "float2 pairs must not straddle a hop block");
    static_assert(HOP % RA == 0, "a column pair must not straddle a hop block");
    static_assert(RA == 2 * kWarps && RA == 16, "one column pair per warp, 16-point second pass");
    static_assert(L <= NFFT && NZ <= RB, "frame does not fit the transform");
};

__device__ __forceinline__ float2 lds_f2(const float *p) { return *reinterpret_cast<const float2 *>(p); }
__device__ __forceinline__ float4 lds_f4(const float *p) { return *reinterpret_cast<const float4 *>(p); }

// int16 pair -> two exact floats without the XU pipe: (v ^ 0x8000) in the mantissa of 2^23.
__device__ __forceinline__ float2 s16x2_to_f32(uint32_t w)
{
    const uint32_t lo = ((w & 0xFFFFu) ^ 0x4B008000u);
    const uint32_t hi = ((w >> 16) ^ 0x4B008000u);
    return make_float2(__uint_as_float(lo) - 8421376.0f, __uint_as_float(hi) - 8421376.0f);
}

__device__ __forceinline__ float to_f32(int16_t v) { return static_cast<float>(v); }
__device__ __forceinline__ float to_f32(float v) { return v; }

template <typename PcmT, int L, int HOP, int RB, int RA>
__global__ void __launch_bounds__(kThreads, 2) fused_ct_kernel(const PcmT *__restrict__ pcm, const CtArgs a)
{
    using G = Geo<L, HOP, RB, RA>;
    constexpr int STRIDE = G::STRIDE, H = G::H, NZ = G::NZ;
    extern __shared__ __align__(16) float smem[];
    float *tab = smem;
    float *staged = smem + a.lay.total;   // S0-S1
    float *pw = staged;                   // S2-S3 (aliases staged)
    float2 *ws = reinterpret_cast<float2 *>(staged + G::UNION);
    float *scr = reinterpret_cast<float *>(ws);   // S3-S5 scratch (aliases ws)

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    // one-time: tables -> shared memory, zero the slack rows of P
    for (int i = threadIdx.x * 4; i < a.lay.total; i += kThreads * 4)
        *reinterpret_cast<float4 *>(tab + i) = __ldg(reinterpret_cast<const float4 *>(a.tab + i));
    for (int i = G::NB * 32 + threadIdx.x; i < G::PW; i += kThreads) pw[i] = 0.0f;

    const float *t_win = tab + a.lay.win, *t_tw = tab + a.lay.tw, *t_twh = tab + a.lay.twh;
    const int4 *t_seg = reinterpret_cast<const int4 *>(tab + a.lay.seg);
    const int *t_seglo = reinterpret_cast<const int *>(tab + a.lay.seg_lo);
    const float *t_melw = tab + a.lay.melw, *t_dct = tab + a.lay.dct;
    const int n_seg = a.lay.n_seg, half = a.lay.half, halfp = a.lay.halfp;
    float *er = scr, *ef = scr + n_seg * 32;
    float *ls = scr + 2 * n_seg * 32, *ld = ls + halfp * 32;
    float *ostage = ld + halfp * 32;

    // ---- S0 machinery.  Fast path (int16, 4-byte aligned, even tile start, no utterance end inside
    // the tile): every warp owns CH consecutive 32-bit words (2 samples each) of the tile and
    // PREFETCHES them into registers one tile ahead, so the HBM latency hides behind S1..S5.
    constexpr int QWORDS = (31 * HOP + L) / 2 + 1;            // words covering i = 0 .. T_max
    constexpr int CH = (QWORDS + kWarps - 1) / kWarps;        // words per warp
    constexpr int NQ = (CH + 31) / 32;                        // steps of 32 words
    uint32_t wreg[NQ];
    float edge = 0.0f;                                        // lane 31: the sample before the warp's first word
    auto tile_fast = [&](const Tile &tl) -> bool {
        if constexpr (sizeof(PcmT) != 2) return false;
        const int T = (tl.n_frames - 1) * HOP + L;
        return ((reinterpret_cast<uintptr_t>(pcm) & 3) == 0) && ((tl.first_sample & 1) == 0) &&
               (tl.utt_end - tl.first_sample >= T + 2) && (a.pcm_len - tl.first_sample >= T + 2);
    };
    auto prefetch = [&](const Tile &tl) {
        const int T = (tl.n_frames - 1) * HOP + L;
        const uint32_t *xw = reinterpret_cast<const uint32_t *>(pcm) + (tl.first_sample >> 1) + warp * CH + lane;
#pragma unroll
        for (int j = 0; j < NQ; ++j) {
            const int ql = 32 * j + lane;
            wreg[j] = (ql < CH && 2 * (warp * CH + ql) <= T) ? __ldg(xw + 32 * j) : 0u;
        }
        edge = 0.0f;
        const int ie = 2 * warp * CH - 1;                     // tile-relative index of the edge sample
        if (lane == 31 && ie <= T && (warp > 0 || tl.first_sample > tl.utt_begin))
            edge = to_f32(pcm[tl.first_sample + ie]);
    };

    Tile tile{};
    bool fast = false;
    if (static_cast<int64_t>(blockIdx.x) < a.n_tiles) {
        tile = a.tiles[blockIdx.x];
        fast = tile_fast(tile);
        if (fast) prefetch(tile);
    }
    for (int64_t t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
        const int n_frames = tile.n_frames;
        const int64_t out_row = tile.out_row;
        __syncthreads();   // previous tile's readers of ws/scr and pw are done; tables are visible

        // ---- S0: stage y[s] = x[s] - a x[s-1] once per sample, padded by kPad words per hop ----
        {
            const int T = (n_frames - 1) * HOP + L;
            if (fast) {
                float vp = edge;
                const int ib = 2 * (warp * CH + lane);
#pragma unroll
                for (int j = 0; j < NQ; ++j) {
                    const int i0 = ib + 64 * j;
                    const float2 x = s16x2_to_f32(wreg[j]);
                    // previous sample = lane-1's second sample; lane 0 takes lane 31's from the step before
                    const float u = lane == 31 ? vp : x.y;
                    const float xp = __shfl_sync(0xffffffffu, u, (lane + 31) & 31);
                    vp = x.y;
                    const float y0 = fmaf(-a.preemph, xp, x.x), y1 = fmaf(-a.preemph, x.x, x.y);
                    if (32 * j + lane < CH && i0 <= T)
                        *reinterpret_cast<float2 *>(staged + G::padded(i0)) = make_float2(y0, y1);
                }
            } else {
                const int64_t room_lo = tile.first_sample - tile.utt_begin;
                const int64_t room_hi = tile.utt_end - tile.first_sample;
                const PcmT *x = pcm + tile.first_sample;
                for (int i = threadIdx.x; i <= T; i += kThreads) {
                    float y = 0.0f;
                    if (i < room_hi) {
                        const float x0 = to_f32(x[i]);
                        const float x1 = (i > -room_lo) ? to_f32(x[i - 1]) : 0.0f;
                        y = fmaf(-a.preemph, x1, x0);
                    }
                    staged[G::padded(i)] = y;
                }
            }
        }
        // the registers are free again: fetch the next tile while this one is transformed
        if (t + gridDim.x < a.n_tiles) {
            tile = a.tiles[t + gridDim.x];
            fast = tile_fast(tile);
            if (fast) prefetch(tile);
        }
        if constexpr (G::SLACK > 0) {
            // the last frame's columns read SLACK words past its end (zero window): keep them finite
            const int T = (n_frames - 1) * HOP + L;
            if (threadIdx.x <= G::SLACK) staged[G::padded(T + 1 + threadIdx.x)] = 0.0f;
        }
        __syncthreads();

        // ---- S1: pass 1.  Warp = column pair (a, a + 1): windowed real DFT-RB over b, inter-pass twiddle ----
        {
            const int pr = warp;
            const float *base = staged + lane * STRIDE + 2 * pr;
            const float *wrow = t_win + pr * (2 * G::NZP);
            float2 in[NZ];
#pragma unroll
            for (int b = 0; b < NZ; b += 2) {
                const float4 w = lds_f4(wrow + 2 * b);
                const float2 y0 = lds_f2(base + G::padded(RA * b));
                in[b] = make_float2(y0.x * w.x, y0.y * w.y);
                if (b + 1 < NZ) {
                    const float2 y1 = lds_f2(base + G::padded(RA * (b + 1)));
                    in[b + 1] = make_float2(y1.x * w.z, y1.y * w.w);
                }
            }
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int col = 2 * pr + half;
                float x[RB];
#pragma unroll
                for (int b = 0; b < RB; ++b) x[b] = b < NZ ? (half ? in[b < NZ ? b : 0].y : in[b < NZ ? b : 0].x) : 0.0f;
                rf::cplx X[H + 1];
                rf::RDft<RB>::template run<NZ>(x, X);
                float2 *wsa = ws + col * 32 + lane;
                wsa[(H - 1) * RA * 32] = make_float2(X[0].re, X[H].re);   // rows 0 and H are real here
                const float *trow = t_tw + col * (2 * H);
#pragma unroll
                for (int k1 = 1; k1 < H; k1 += 2) {
                    const float4 tw = lds_f4(trow + 2 * (k1 - 1));       // twiddles of k1, k1 + 1
                    const rf::cplx v = rf::cmulc(X[k1], tw.x, tw.y);
                    wsa[(k1 - 1) * RA * 32] = make_float2(v.re, v.im);
                    if (k1 + 1 < H) {
                        const rf::cplx u = rf::cmulc(X[k1 + 1], tw.z, tw.w);
                        wsa[k1 * RA * 32] = make_float2(u.re, u.im);
                    }
                }
            }
        }
        __syncthreads();

        // ---- S2: pass 2.  Item = one row k1: complex DFT-RA over a gives bins k1 + RB k2; power ----
        {
#pragma unroll 1
            for (int it = warp; it < H; it += kWarps) {
                if (it < H - 1) {
                    const int k1 = it + 1;
                    const float2 *row = ws + (k1 - 1) * RA * 32 + lane;
                    rf::cplx z[RA];
#pragma unroll
                    for (int c = 0; c < RA; ++c) {
                        const float2 p = row[c * 32];
                        z[c] = rf::cplx{p.x, p.y};
                    }
                    rf::cdft16(z);
                    float *p_lo = pw + k1 * 32 + lane;            // bins k1 + RB k2, k2 < RA/2
                    float *p_hi = pw + (RB - k1) * 32 + lane;     // mirrored: N - k = (RB - k1) + RB (RA - 1 - k2)
#pragma unroll
                    for (int k2 = 0; k2 < RA / 2; ++k2)
                        p_lo[RB * k2 * 32] = fmaf(z[k2].re, z[k2].re, z[k2].im * z[k2].im);
#pragma unroll
                    for (int k2 = RA / 2; k2 < RA; ++k2)
                        p_hi[RB * (RA - 1 - k2) * 32] = fmaf(z[k2].re, z[k2].re, z[k2].im * z[k2].im);
                } else {
                    // rows 0 and H: both real after pass 1.  Row 0 -> real DFT-16 -> bins RB k2;
                    // row H times W_(2 RA)^a -> complex DFT-16 -> bins H + RB k2, k2 < RA/2
                    const float2 *row = ws + (H - 1) * RA * 32 + lane;
                    float r0[RA];
                    rf::cplx zh[RA];
#pragma unroll
                    for (int c = 0; c < RA; c += 2) {
                        const float4 tw = lds_f4(t_twh + 2 * c);
                        const float2 p = row[c * 32], q = row[(c + 1) * 32];
                        r0[c] = p.x;
                        r0[c + 1] = q.x;
                        zh[c] = rf::cplx{p.y * tw.x, p.y * tw.y};
                        zh[c + 1] = rf::cplx{q.y * tw.z, q.y * tw.w};
                    }
                    rf::cplx X0[RA / 2 + 1];
                    rf::rdft16<16>(r0, X0);
                    pw[lane] = X0[0].re * X0[0].re;
                    pw[(RB * (RA / 2)) * 32 + lane] = X0[RA / 2].re * X0[RA / 2].re;
#pragma unroll
                    for (int k2 = 1; k2 < RA / 2; ++k2)
                        pw[(RB * k2) * 32 + lane] = fmaf(X0[k2].re, X0[k2].re, X0[k2].im * X0[k2].im);
                    rf::cdft16(zh);
#pragma unroll
                    for (int k2 = 0; k2 < RA / 2; ++k2)
                        pw[(H + RB * k2) * 32 + lane] = fmaf(zh[k2].re, zh[k2].re, zh[k2].im * zh[k2].im);
                }
            }
        }
        __syncthreads();

        // ---- S3: sparse mel in segment form, 4 bins per step (weights zero past the segment) ----
        {
            const int j0 = t_seglo[warp], j1 = t_seglo[warp + 1];
#pragma unroll 1
            for (int j = j0; j < j1; ++j) {
                const int4 sg = t_seg[j];
                const float *p = pw + sg.x * 32 + lane;
                const float *w = t_melw + sg.z;
                float r0 = 0.0f, r1 = 0.0f, f0 = 0.0f, f1 = 0.0f;
#pragma unroll 1
                for (int c = 0; c < sg.y; ++c) {
                    const float4 wr = lds_f4(w), wf = lds_f4(w + 4);
                    const float p0 = p[0], p1 = p[32], p2 = p[64], p3 = p[96];
                    r0 = fmaf(wr.x, p0, r0); f0 = fmaf(wf.x, p0, f0);
                    r1 = fmaf(wr.y, p1, r1); f1 = fmaf(wf.y, p1, f1);
                    r0 = fmaf(wr.z, p2, r0); f0 = fmaf(wf.z, p2, f0);
                    r1 = fmaf(wr.w, p3, r1); f1 = fmaf(wf.w, p3, f1);
                    p += 128;
                    w += 8;
                }
                er[j * 32 + lane] = r0 + r1;
                ef[j * 32 + lane] = f0 + f1;
            }
        }
        __syncthreads();

        // ---- S4: log; symmetric / antisymmetric halves for the DCT (cos(pi k (M-1-m+1/2)/M) = (-1)^k cos(...)) ----
        const bool live = lane < n_frames;
        {
            const int M = a.n_mel;
#pragma unroll 1
            for (int mh = warp; mh < half; mh += kWarps) {
                const int m2 = M - 1 - mh;
                const float l1 = __logf(fmaxf(er[mh * 32 + lane] + ef[(mh + 1) * 32 + lane], a.log_floor));
                float l2 = 0.0f;
                if (m2 != mh) l2 = __logf(fmaxf(er[m2 * 32 + lane] + ef[(m2 + 1) * 32 + lane], a.log_floor));
                if (a.logmel) {
                    if (live) {
                        float *o = a.out + (out_row + lane) * M;
                        o[mh] = l1;
                        if (m2 != mh) o[m2] = l2;
                    }
                } else {
                    ls[mh * 32 + lane] = m2 != mh ? l1 + l2 : l1;
                    ld[mh * 32 + lane] = m2 != mh ? l1 - l2 : 0.0f;
                }
            }
            // rows [half, halfp) are read by S5 with zero weights; the scratch aliases ws, whose
            // float2 layout puts OTHER lanes' data there, so they must be finite: zero them
            for (int r = half + warp; r < halfp; r += kWarps) {
                ls[r * 32 + lane] = 0.0f;
                ld[r * 32 + lane] = 0.0f;
            }
        }
        if (a.logmel) continue;   // next tile starts with a barrier
        __syncthreads();

        // ---- S5: DCT-II on the halves; even rows read ls, odd rows ld ----
        {
#pragma unroll 1
            for (int k = warp; k < a.n_cep; k += kWarps) {
                const float *src = ((k & 1) ? ld : ls) + lane;
                const float *d = t_dct + k * halfp;
                float c0 = 0.0f, c1 = 0.0f;
#pragma unroll 1
                for (int m = 0; m < halfp; m += 4) {
                    const float4 dv = lds_f4(d + m);
                    c0 = fmaf(dv.x, src[(m + 0) * 32], c0);
                    c1 = fmaf(dv.y, src[(m + 1) * 32], c1);
                    c0 = fmaf(dv.z, src[(m + 2) * 32], c0);
                    c1 = fmaf(dv.w, src[(m + 3) * 32], c1);
                }
                ostage[lane * a.n_cep + k] = c0 + c1;
            }
        }
        __syncthreads();
        {
            const int total = n_frames * a.n_cep;
            float *o = a.out + out_row * a.n_cep;
            for (int i = threadIdx.x; i < total; i += kThreads) o[i] = ostage[i];
        }
    }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
struct CtVariant {
    int L, hop, rb, ra;
    const char *name;
};
constexpr CtVariant kVariants[] = {
    {400, 160, 32, 16, "fused_ct_tile32_L400_H160_real32x16"},   // BASELINE.json configs 1, 2, 5 (16 kHz)
    {200, 80, 16, 16, "fused_ct_tile32_L200_H80_real16x16"},     // BASELINE.json config 3 (8 kHz telephony)
};

struct CtState {
    const CtVariant *v = nullptr;
    CtLayout lay{};
    float *d_tab = nullptr;
    size_t smem = 0;
    int sm_count = 0;
};

template <int L, int HOP, int RB, int RA>
size_t smem_floats_fixed() { return Geo<L, HOP, RB, RA>::UNION + Geo<L, HOP, RB, RA>::WS; }

size_t smem_fixed(const CtVariant &v)
{
    if (v.L == 400) return smem_floats_fixed<400, 160, 32, 16>();
    return smem_floats_fixed<200, 80, 16, 16>();
}

}  // namespace

const char *ct_match(const mfcc_params &p)
{
    for (const auto &v : kVariants) {
        if (p.frame_len != v.L || p.hop_len != v.hop || p.nfft != v.rb * v.ra) continue;
        // scratch (er | ef | ls | ld | ostage) must fit in the workspace, tables in the smem budget
        const int n_seg = p.n_mel + 1, halfp = ((p.n_mel + 1) / 2 + 3) / 4 * 4;
        const size_t scratch = static_cast<size_t>(2 * n_seg + 2 * halfp) * 32 + 32ull * p.n_cep;
        if (scratch > static_cast<size_t>(v.rb) * v.ra * 32) return nullptr;
        return v.name;
    }
    return nullptr;
}

int ct_prepare(mfcc_plan *plan)
{
    const mfcc_params &p = plan->p;
    const CtVariant *var = nullptr;
    for (const auto &v : kVariants)
        if (p.frame_len == v.L && p.hop_len == v.hop && p.nfft == v.rb * v.ra) var = &v;
    if (var == nullptr) return MFCC_ENOTSUP;
    const int RB = var->rb, RA = var->ra, N = RB * RA, H = RB / 2, M = p.n_mel;
    const int NZ = (p.frame_len + RA - 1) / RA, NZP = (NZ + 1) / 2 * 2;
    const HostTables &h = plan->host;
    std::vector<float> tab;
    CtLayout lay{};
    auto align4 = [&]() { while (tab.size() % 4) tab.push_back(0.0f); };
    auto push_int = [&](int v) { float f; std::memcpy(&f, &v, 4); tab.push_back(f); };

    // window of the sample pairs (a, a + 1) + RA b per column pair, zero past the frame
    lay.win = static_cast<int>(tab.size());
    for (int pr = 0; pr < RA / 2; ++pr)
        for (int b = 0; b < NZP; ++b)
            for (int e = 0; e < 2; ++e) {
                const int i = 2 * pr + e + RA * b;
                tab.push_back(b < NZ && i < p.frame_len ? h.window[i] : 0.0f);
            }
    // inter-pass twiddles W_N^(a k1), k1 = 1 .. H-1 at slot k1 - 1
    lay.tw = static_cast<int>(tab.size());
    for (int col = 0; col < RA; ++col)
        for (int sl = 0; sl < H; ++sl) {
            const double ang = sl < H - 1 ? -2.0 * M_PI * static_cast<double>(col) * (sl + 1) / N : 0.0;
            tab.push_back(static_cast<float>(std::cos(ang)));
            tab.push_back(static_cast<float>(std::sin(ang)));
        }
    // twiddle of row H: W_N^(a H) = W_(2 RA)^a
    lay.twh = static_cast<int>(tab.size());
    for (int col = 0; col < RA; ++col) {
        const double ang = -2.0 * M_PI * col / (2.0 * RA);
        tab.push_back(static_cast<float>(std::cos(ang)));
        tab.push_back(static_cast<float>(std::sin(ang)));
    }
    // mel segments in 4-bin chunks, weights pre-scaled by 1/N (pass 2 leaves |X|^2)
    const int n_seg = M + 1;
    std::vector<float> melw;
    std::vector<int> seg;   // 4 ints per segment
    const double scale = 1.0 / N;
    for (int j = 0; j < n_seg; ++j) {
        const int k0 = h.mel_bins[j], k1 = h.mel_bins[j + 1];
        const int chunks = (k1 - k0 + 3) / 4;
        seg.push_back(k0);
        seg.push_back(chunks);
        seg.push_back(static_cast<int>(melw.size()));
        seg.push_back(0);
        for (int c = 0; c < chunks; ++c) {
            for (int e = 0; e < 4; ++e) {
                const int k = k0 + 4 * c + e;
                melw.push_back(k < k1 ? static_cast<float>(static_cast<double>(h.rise[k]) * scale) : 0.0f);
            }
            for (int e = 0; e < 4; ++e) {
                const int k = k0 + 4 * c + e;
                melw.push_back(k < k1 ? static_cast<float>(static_cast<double>(h.fall[k]) * scale) : 0.0f);
            }
        }
    }
    align4();
    lay.seg = static_cast<int>(tab.size());
    for (int v : seg) push_int(v);
    lay.seg_lo = static_cast<int>(tab.size());
    {
        // contiguous split of the segments over the warps, balanced by cost (3 per chunk + 2 per segment)
        std::vector<int> cost(n_seg);
        int total = 0;
        for (int j = 0; j < n_seg; ++j) { cost[j] = seg[4 * j + 1] * 3 + 2; total += cost[j]; }
        int j = 0, acc = 0;
        for (int w = 0; w < kWarps; ++w) {
            push_int(j);
            const double target = static_cast<double>(total) * (w + 1) / kWarps;
            while (j < n_seg && acc + cost[j] * 0.5 <= target) acc += cost[j++];
        }
        push_int(n_seg);   // the last warp ends at n_seg (its target is the total)
    }
    align4();
    lay.melw = static_cast<int>(tab.size());
    tab.insert(tab.end(), melw.begin(), melw.end());
    align4();
    // DCT rows on the first half (symmetry), zero padded to a multiple of 4
    const int half = (M + 1) / 2, halfp = (half + 3) / 4 * 4;
    lay.dct = static_cast<int>(tab.size());
    for (int k = 0; k < p.n_cep; ++k)
        for (int m = 0; m < halfp; ++m) tab.push_back(m < half ? h.dct[static_cast<size_t>(k) * M + m] : 0.0f);
    align4();
    lay.total = static_cast<int>(tab.size());
    lay.n_seg = n_seg;
    lay.half = half;
    lay.halfp = halfp;
    if (static_cast<size_t>(lay.total) * 4 > 12 * 1024) return MFCC_ENOTSUP;   // keeps 2 CTAs/SM at N = 512

    CtState *st = new CtState();
    st->v = var;
    st->lay = lay;
    st->sm_count = plan->sm_count;
    st->smem = sizeof(float) * (static_cast<size_t>(lay.total) + smem_fixed(*var));
    if (cudaMalloc(&st->d_tab, sizeof(float) * tab.size()) != cudaSuccess) {
        cudaGetLastError();
        delete st;
        return MFCC_ENOMEM;
    }
    if (cudaMemcpy(st->d_tab, tab.data(), sizeof(float) * tab.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaFree(st->d_tab);
        delete st;
        return MFCC_ECUDA;
    }
    plan->ct_state = st;
    return MFCC_OK;
}

void ct_release(mfcc_plan *plan)
{
    CtState *st = static_cast<CtState *>(plan->ct_state);
    if (st == nullptr) return;
    if (st->d_tab) cudaFree(st->d_tab);
    delete st;
    plan->ct_state = nullptr;
}

template <typename PcmT, int L, int HOP, int RB, int RA>
static int launch_variant(const mfcc_plan *plan, const CtState *st, const Tile *d_tiles, int64_t n_tiles,
                          const PcmT *d_pcm, int64_t pcm_len, float *d_out, cudaStream_t stream)
{
    auto kern = fused_ct_kernel<PcmT, L, HOP, RB, RA>;
    static thread_local const void *configured = nullptr;
    if (configured != reinterpret_cast<const void *>(kern)) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
            cudaGetLastError();
            return MFCC_ECUDA;
        }
        configured = reinterpret_cast<const void *>(kern);
    }
    const mfcc_params &p = plan->p;
    CtArgs a;
    a.tiles = d_tiles;
    a.n_tiles = n_tiles;
    a.pcm_len = pcm_len;
    a.out = d_out;
    a.tab = st->d_tab;
    a.lay = st->lay;
    a.n_mel = p.n_mel;
    a.n_cep = p.n_cep;
    a.logmel = p.output == MFCC_OUT_LOGMEL;
    a.preemph = p.preemph;
    a.log_floor = p.log_floor;
    const int per_sm = st->smem * 2 + 2048 <= 227 * 1024 ? 2 : 1;
    const int64_t grid = std::min<int64_t>(n_tiles, static_cast<int64_t>(st->sm_count) * per_sm);
    kern<<<static_cast<unsigned>(grid), kThreads, st->smem, stream>>>(d_pcm, a);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError() == cudaSuccess ? MFCC_OK : MFCC_ECUDA;
}

template <typename PcmT>
int ct_launch(const mfcc_plan *plan, const Tile *d_tiles, int64_t n_tiles, const PcmT *d_pcm, int64_t pcm_len,
              float *d_out, cudaStream_t stream)
{
    const CtState *st = static_cast<const CtState *>(plan->ct_state);
    if (st == nullptr) return MFCC_ENOTSUP;
    if (st->v->L == 400)
        return launch_variant<PcmT, 400, 160, 32, 16>(plan, st, d_tiles, n_tiles, d_pcm, pcm_len, d_out, stream);
    return launch_variant<PcmT, 200, 80, 16, 16>(plan, st, d_tiles, n_tiles, d_pcm, pcm_len, d_out, stream);
}

template int ct_launch<int16_t>(const mfcc_plan *, const Tile *, int64_t, const int16_t *, int64_t, float *,
                                cudaStream_t);
template int ct_launch<float>(const mfcc_plan *, const Tile *, int64_t, const float *, int64_t, float *,
                              cudaStream_t);

}  // namespace mfcc
