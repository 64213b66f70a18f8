// mfcc_fused_sp.cu — fully SPECIALISED fused tile kernel: frame geometry AND the mel
// filterbank's bin edges are compile-time, so the sparse filterbank, the log and the DCT
// unroll into straight-line code with immediate shared-memory offsets and constant-bank
// (kernel-parameter) weights.  Runtime values — window, twiddles, mel weights, DCT rows
// (lifter folded in), pre-emphasis, log floor — stay data, so a plan only needs its
// *structure* (frame/hop/NFFT/n_mel/n_cep and the integer bin edges) to match a variant.
//
// What changed against mfcc_fused_ct.cu and why (profiles/r1_v4_real32x16_A.md: the two FFT
// passes were 56 % of the instructions but 37 % of the time; staging, mel, log, DCT and store
// took the rest in seven barrier-separated, latency-bound phases):
//   * S0 input: the next tile's PCM arrives by ONE bulk async copy (cp.async.bulk, TMA engine)
//     into a raw int16 buffer while the current tile is transformed — no prefetch registers,
//     no LDG/LDL instructions, completion on an mbarrier.  Staging then converts 8 samples per
//     thread (LDS.128) instead of 2.
//   * tail: each warp OWNS a contiguous group of mel filters (balanced at compile time), sums
//     their bins from P, takes the log and accumulates its share of every cepstrum in
//     registers; one barrier later the 8 partial cepstra per frame are added and stored
//     coalesced.  4 CTA barriers per tile instead of 7.
// Phases S1 (windowed real DFT-RB per column pair + inter-pass twiddle) and S2 (complex DFT-RA per
// row + power) are those of mfcc_fused_ct.cu.
//
// No reference code corresponds to this (SURVEY.md §8a "Ref file:line = none").
#include <cuda_runtime.h>

#include <cmath>
#include <cstring>
#include <utility>
#include <vector>

#include "mfcc_rfft.cuh"
#include "mfcc_host.h"

namespace mfcc {

namespace {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int kPad = 2;

// ---- compile-time loops ----
template <class F, int... I>
__device__ __forceinline__ void static_for_impl(F &&f, std::integer_sequence<int, I...>)
{
    (f(std::integral_constant<int, I>{}), ...);
}
template <int N, class F>
__device__ __forceinline__ void static_for(F &&f)
{
    static_for_impl(f, std::make_integer_sequence<int, N>{});
}

// ---- variants: structure only (BASELINE.json configs; bin edges = floor((N+1) f / sr) on HTK mel) ----
struct Var16k {   // configs[0], [1], [4]: 16 kHz, 25/10 ms, 512-pt, 26 mel over 0 .. 8 kHz, 13 cepstra
    static constexpr int L = 400, HOP = 160, RB = 32, RA = 16, NMEL = 26, NCEP = 13;
    static constexpr int bin(int j)
    {
        constexpr int b[NMEL + 2] = {0, 2, 4, 7, 10, 13, 16, 20, 24, 29, 34, 40, 46, 53,
                                     60, 68, 77, 87, 97, 109, 122, 136, 152, 169, 188, 209, 231, 256};
        return b[j];
    }
    static constexpr const char *name() { return "fused_sp_tile32_L400_H160_real32x16_mel26_cep13"; }
};
struct Var8k {    // configs[2]: 8 kHz telephony, 25/10 ms, 256-pt, 20 mel over 0 .. 4 kHz, 13 cepstra
    static constexpr int L = 200, HOP = 80, RB = 16, RA = 16, NMEL = 20, NCEP = 13;
    static constexpr int bin(int j)
    {
        constexpr int b[NMEL + 2] = {0, 2, 4, 7, 9, 12, 16, 19, 23, 28, 33, 38, 44, 50, 57, 65, 73, 82, 92, 103, 115, 128};
        return b[j];
    }
    static constexpr const char *name() { return "fused_sp_tile32_L200_H80_real16x16_mel20_cep13"; }
};

template <class V>
struct Geo {
    static constexpr int L = V::L, HOP = V::HOP, RB = V::RB, RA = V::RA, NMEL = V::NMEL, NCEP = V::NCEP;
    static constexpr int NFFT = RB * RA, NB = NFFT / 2 + 1, H = RB / 2;
    static constexpr int NZ = (L + RA - 1) / RA;           // rows b of a column that carry samples
    static constexpr int NZP = (NZ + 1) / 2 * 2;
    static constexpr int STRIDE = HOP + kPad;              // staged words per hop block
    static constexpr int padded(int i) { return i + kPad * (i / HOP); }
    static constexpr int SLACK = RA * NZ - L;              // words past a frame's end read with a zero window
    static constexpr int tceil(int n_frames) { return ((n_frames - 1) * HOP + L + SLACK + 7) / 8 * 8; }
    static constexpr int TCEIL = tceil(32);                // samples staged for a full tile
    static constexpr int STAGED = padded(TCEIL) + 8;
    static constexpr int PW = NB * 32;
    static constexpr int UNION = ((STAGED > PW ? STAGED : PW) + 3) / 4 * 4;
    static constexpr int WS = H * RA * 32 * 2;             // floats
    static constexpr int RAW = (TCEIL + 8) / 2;            // floats holding TCEIL + 8 int16 samples
    // table blob (floats)
    static constexpr int T_WIN = 0;                        // [RA/2][NZP] float2
    static constexpr int T_TW = T_WIN + RA / 2 * NZP * 2;  // [RA][H] float2, slot k1 - 1
    static constexpr int T_TWH = T_TW + RA * H * 2;        // [RA] float2
    static constexpr int TABF = T_TWH + RA * 2;
    static constexpr int SMEM_FLOATS = TABF + UNION + WS + RAW + 4;   // + mbarrier (8 B, 16-B slot)
    static constexpr int PS = NCEP | 1;                    // partial-cepstra row stride (odd: conflict-free)
    static constexpr int LS = NMEL | 1;                    // log-mel staging row stride
    static_assert(HOP % 8 == 0, "an 8-sample chunk must not straddle a hop block");
    static_assert(HOP % RA == 0, "a column pair must not straddle a hop block");
    static_assert(RA == 2 * kWarps && RA == 16, "one column pair per warp, 16-point second pass");
    static_assert(L <= NFFT && NZ <= RB, "frame does not fit the transform");
    static_assert(TABF % 4 == 0 && UNION % 4 == 0 && WS % 4 == 0 && RAW % 4 == 0, "16-byte aligned regions");
    static_assert(kWarps * 32 * PS <= WS && 32 * LS <= WS, "tail scratch must fit in the workspace");
    static_assert(V::bin(NMEL + 1) <= NFFT / 2 && V::bin(0) >= 0, "bin edges inside the spectrum");
};

// Filters [beg(w), beg(w+1)) belong to warp w: contiguous, balanced by FMA count
// (bins of the triangle + the filter's column of the DCT + the log).
template <class V>
struct Part {
    static constexpr int cost(int m) { return (V::bin(m + 2) - V::bin(m)) + V::NCEP + 6; }
    static constexpr int beg(int w)
    {
        if (w <= 0) return 0;
        if (w >= kWarps) return V::NMEL;
        int total = 0;
        for (int m = 0; m < V::NMEL; ++m) total += cost(m);
        int acc = 0, m = 0;
        // first filter whose midpoint lies past w/kWarps of the total cost
        while (m < V::NMEL && (2 * acc + cost(m)) * kWarps <= 2 * total * w) acc += cost(m++);
        return m;
    }
};

template <class V>
struct SpArgs {
    const Tile *tiles;
    int64_t n_tiles;
    float *out;
    const float *tab;               // global copy of the table blob (Geo::TABF floats)
    int logmel;
    float preemph, log_floor;
    float rise[Geo<V>::NB];         // pre-scaled by 1 / NFFT (pass 2 leaves |X|^2)
    float fall[Geo<V>::NB];
    float dct[V::NCEP * V::NMEL];   // [k][m]
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

// ---- mbarrier + bulk async copy (TMA engine, 1-D) ----
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// ---- tail: warp W's filters -> log -> its share of every cepstrum ----
template <class V, int W>
__device__ __forceinline__ void tail_group(const float *__restrict__ pl, const SpArgs<V> &a, float *__restrict__ scr,
                                           int lane)
{
    using G = Geo<V>;
    constexpr int m0 = Part<V>::beg(W), m1 = Part<V>::beg(W + 1), NF = m1 - m0;
    if constexpr (NF > 0) {
        float acc[2 * NF];
#pragma unroll
        for (int i = 0; i < 2 * NF; ++i) acc[i] = 0.0f;
        // segments j = m0 .. m1: bin k of segment j rises into filter j and falls out of filter j - 1
        static_for<NF + 1>([&](auto jc) {
            constexpr int j = m0 + decltype(jc)::value;
            constexpr int k0 = V::bin(j), k1 = V::bin(j + 1);
            static_for<(k1 > k0 ? k1 - k0 : 0)>([&](auto kc) {
                constexpr int k = k0 + decltype(kc)::value;
                const float p = pl[k * 32];
                if constexpr (j < m1) acc[2 * (j - m0) + (k & 1)] = fmaf(a.rise[k], p, acc[2 * (j - m0) + (k & 1)]);
                if constexpr (j > m0)
                    acc[2 * (j - 1 - m0) + (k & 1)] = fmaf(a.fall[k], p, acc[2 * (j - 1 - m0) + (k & 1)]);
            });
        });
        float lg[NF];
#pragma unroll
        for (int i = 0; i < NF; ++i) lg[i] = __logf(fmaxf(acc[2 * i] + acc[2 * i + 1], a.log_floor));
        if (a.logmel) {
#pragma unroll
            for (int i = 0; i < NF; ++i) scr[lane * G::LS + m0 + i] = lg[i];
        } else {
            float c[V::NCEP];
#pragma unroll
            for (int k = 0; k < V::NCEP; ++k) c[k] = 0.0f;
            static_for<NF>([&](auto mc) {
                constexpr int i = decltype(mc)::value;
                static_for<V::NCEP>([&](auto kc) {
                    constexpr int k = decltype(kc)::value;
                    c[k] = fmaf(a.dct[k * V::NMEL + m0 + i], lg[i], c[k]);
                });
            });
            float *dst = scr + (W * 32 + lane) * G::PS;
#pragma unroll
            for (int k = 0; k < V::NCEP; ++k) dst[k] = c[k];
        }
    } else if (!a.logmel) {
        float *dst = scr + (W * 32 + lane) * G::PS;
#pragma unroll
        for (int k = 0; k < V::NCEP; ++k) dst[k] = 0.0f;
    }
}

template <typename PcmT, class V>
__global__ void __launch_bounds__(kThreads, 2)
fused_sp_kernel(const PcmT *__restrict__ pcm, const __grid_constant__ SpArgs<V> a)
{
    using G = Geo<V>;
    constexpr int L = G::L, HOP = G::HOP, RB = G::RB, RA = G::RA, STRIDE = G::STRIDE, H = G::H, NZ = G::NZ;
    extern __shared__ __align__(16) float smem[];
    float *tab = smem;
    float *staged = smem + G::TABF;       // S0-S1
    float *pw = staged;                   // S2-S3 (aliases staged)
    float2 *ws = reinterpret_cast<float2 *>(staged + G::UNION);
    float *scr = reinterpret_cast<float *>(ws);   // tail scratch (aliases ws)
    const int16_t *raw16 = reinterpret_cast<const int16_t *>(staged + G::UNION + G::WS);
    const uint32_t raw_s = smem_u32(raw16);
    const uint32_t bar = smem_u32(staged + G::UNION + G::WS + G::RAW);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    if (threadIdx.x == 0) mbar_init(bar, 1);
    for (int i = threadIdx.x * 4; i < G::TABF; i += kThreads * 4)
        *reinterpret_cast<float4 *>(tab + i) = __ldg(reinterpret_cast<const float4 *>(a.tab + i));
    const float *t_win = tab + G::T_WIN, *t_tw = tab + G::T_TW, *t_twh = tab + G::T_TWH;

    // A tile takes the bulk-copy path when its samples are 16-byte aligned in HBM and lie inside
    // the utterance up to the staging granule (no zero fill needed).
    auto tile_fast = [&](const Tile &tl) -> bool {
        if constexpr (sizeof(PcmT) != 2) return false;
        return ((reinterpret_cast<uintptr_t>(pcm) & 15) == 0) && ((tl.first_sample & 7) == 0) &&
               (tl.utt_end - tl.first_sample >= G::tceil(tl.n_frames));
    };
    // raw16[8 + i] = x[first_sample + i]; the 8 samples before the tile ride along when they exist
    auto issue_copy = [&](const Tile &tl) {
        const int lead = tl.first_sample >= 8 ? 8 : 0;
        const uint32_t bytes = static_cast<uint32_t>(lead + G::tceil(tl.n_frames)) * 2u;
        mbar_expect_tx(bar, bytes);
        bulk_g2s(raw_s + (8 - lead) * 2, pcm + tl.first_sample - lead, bytes, bar);
    };

    __syncthreads();   // mbarrier initialised, tables visible
    Tile nt{};
    bool nfast = false;
    if (static_cast<int64_t>(blockIdx.x) < a.n_tiles) {
        nt = a.tiles[blockIdx.x];
        nfast = tile_fast(nt);
        if (threadIdx.x == 0 && nfast) issue_copy(nt);
    }
    uint32_t phase = 0;
    for (int64_t t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
        const Tile tile = nt;
        const bool fast = nfast;
        const bool has_next = t + gridDim.x < a.n_tiles;
        if (has_next) nt = a.tiles[t + gridDim.x];   // arrives while S0 runs
        const int n_frames = tile.n_frames;
        const int tc = G::tceil(n_frames);

        // ---- S0: stage y[s] = x[s] - a x[s-1] once per sample, padded by kPad words per hop ----
        if (fast) {
            mbar_wait(bar, phase);
            phase ^= 1u;
            const bool at_start = tile.first_sample == tile.utt_begin;
            const int nchunks = tc >> 3;
#pragma unroll 1
            for (int c = threadIdx.x; c < nchunks; c += kThreads) {
                const uint4 q = *reinterpret_cast<const uint4 *>(raw16 + 8 + 8 * c);
                const uint32_t pwd = *reinterpret_cast<const uint32_t *>(raw16 + 6 + 8 * c);
                const float2 x01 = s16x2_to_f32(q.x), x23 = s16x2_to_f32(q.y);
                const float2 x45 = s16x2_to_f32(q.z), x67 = s16x2_to_f32(q.w);
                float xp = s16x2_to_f32(pwd).y;
                if (c == 0 && at_start) xp = 0.0f;
                const float na = -a.preemph;
                float *dst = staged + 8 * c + kPad * (c / (HOP / 8));
                *reinterpret_cast<float2 *>(dst + 0) = make_float2(fmaf(na, xp, x01.x), fmaf(na, x01.x, x01.y));
                *reinterpret_cast<float2 *>(dst + 2) = make_float2(fmaf(na, x01.y, x23.x), fmaf(na, x23.x, x23.y));
                *reinterpret_cast<float2 *>(dst + 4) = make_float2(fmaf(na, x23.y, x45.x), fmaf(na, x45.x, x45.y));
                *reinterpret_cast<float2 *>(dst + 6) = make_float2(fmaf(na, x45.y, x67.x), fmaf(na, x67.x, x67.y));
            }
        } else {
            const int64_t room_lo = tile.first_sample - tile.utt_begin;
            const int64_t room_hi = tile.utt_end - tile.first_sample;
            const PcmT *x = pcm + tile.first_sample;
            for (int i = threadIdx.x; i < tc; i += kThreads) {
                float y = 0.0f;
                if (i < room_hi) {
                    const float x0 = to_f32(x[i]);
                    const float x1 = (i > -room_lo) ? to_f32(x[i - 1]) : 0.0f;
                    y = fmaf(-a.preemph, x1, x0);
                }
                staged[G::padded(i)] = y;
            }
        }
        __syncthreads();   // B1: staged complete; raw buffer and (previous tile's) scratch free
        nfast = false;
        if (has_next) {
            nfast = tile_fast(nt);
            if (threadIdx.x == 0 && nfast) issue_copy(nt);
        }

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
        __syncthreads();   // B2

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
        __syncthreads();   // B3: P complete, workspace free

        // ---- S3: per-warp filter group: sparse mel -> log -> partial DCT (or log-mel staging) ----
        {
            const float *pl = pw + lane;
            switch (warp) {
                case 0: tail_group<V, 0>(pl, a, scr, lane); break;
                case 1: tail_group<V, 1>(pl, a, scr, lane); break;
                case 2: tail_group<V, 2>(pl, a, scr, lane); break;
                case 3: tail_group<V, 3>(pl, a, scr, lane); break;
                case 4: tail_group<V, 4>(pl, a, scr, lane); break;
                case 5: tail_group<V, 5>(pl, a, scr, lane); break;
                case 6: tail_group<V, 6>(pl, a, scr, lane); break;
                default: tail_group<V, 7>(pl, a, scr, lane); break;
            }
        }
        __syncthreads();   // B4

        // ---- S4: add the partial cepstra of the 8 warps and store coalesced ----
        if (a.logmel) {
            const int total = n_frames * V::NMEL;
            float *o = a.out + tile.out_row * V::NMEL;
            for (int i = threadIdx.x; i < total; i += kThreads) {
                const int f = i / V::NMEL, m = i - f * V::NMEL;
                o[i] = scr[f * G::LS + m];
            }
        } else {
            const int total = n_frames * V::NCEP;
            float *o = a.out + tile.out_row * V::NCEP;
            for (int i = threadIdx.x; i < total; i += kThreads) {
                const int f = i / V::NCEP, k = i - f * V::NCEP;
                const float *src = scr + f * G::PS + k;
                float s = src[0];
#pragma unroll
                for (int w = 1; w < kWarps; ++w) s += src[w * 32 * G::PS];
                o[i] = s;
            }
        }
        // no barrier here: the next S0 writes `staged`, which nobody reads any more; the scratch is
        // next written by S1, after B1.
    }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
struct SpState {
    int variant = -1;          // 0 = Var16k, 1 = Var8k
    std::vector<char> args;    // SpArgs<V> image with the per-launch fields blank
    float *d_tab = nullptr;
    int sm_count = 0;
};

template <class V>
bool structure_matches(const mfcc_params &p, const HostTables &h)
{
    if (p.frame_len != V::L || p.hop_len != V::HOP || p.nfft != V::RB * V::RA) return false;
    if (p.n_mel != V::NMEL || p.n_cep != V::NCEP) return false;
    if (static_cast<int>(h.mel_bins.size()) != V::NMEL + 2) return false;
    for (int j = 0; j < V::NMEL + 2; ++j)
        if (h.mel_bins[j] != V::bin(j)) return false;
    return true;
}

template <class V>
void build_tab_and_args(const mfcc_plan *plan, std::vector<float> &tab, std::vector<char> &args_img)
{
    using G = Geo<V>;
    const mfcc_params &p = plan->p;
    const HostTables &h = plan->host;
    constexpr int RB = V::RB, RA = V::RA, N = RB * RA, H = RB / 2;
    tab.assign(G::TABF, 0.0f);
    // window of the sample pairs (a, a + 1) + RA b per column pair, zero past the frame
    for (int pr = 0; pr < RA / 2; ++pr)
        for (int b = 0; b < G::NZP; ++b)
            for (int e = 0; e < 2; ++e) {
                const int i = 2 * pr + e + RA * b;
                tab[G::T_WIN + (pr * G::NZP + b) * 2 + e] = (b < G::NZ && i < p.frame_len) ? h.window[i] : 0.0f;
            }
    // inter-pass twiddles W_N^(a k1), k1 = 1 .. H-1 at slot k1 - 1 (last slot: 1)
    for (int col = 0; col < RA; ++col)
        for (int sl = 0; sl < H; ++sl) {
            const double ang = sl < H - 1 ? -2.0 * M_PI * static_cast<double>(col) * (sl + 1) / N : 0.0;
            tab[G::T_TW + (col * H + sl) * 2 + 0] = static_cast<float>(std::cos(ang));
            tab[G::T_TW + (col * H + sl) * 2 + 1] = static_cast<float>(std::sin(ang));
        }
    // twiddle of row H: W_N^(a H) = W_(2 RA)^a
    for (int col = 0; col < RA; ++col) {
        const double ang = -2.0 * M_PI * col / (2.0 * RA);
        tab[G::T_TWH + col * 2 + 0] = static_cast<float>(std::cos(ang));
        tab[G::T_TWH + col * 2 + 1] = static_cast<float>(std::sin(ang));
    }
    args_img.assign(sizeof(SpArgs<V>), 0);
    SpArgs<V> *a = reinterpret_cast<SpArgs<V> *>(args_img.data());
    const double scale = 1.0 / N;
    for (int k = 0; k < G::NB; ++k) {
        a->rise[k] = static_cast<float>(static_cast<double>(h.rise[k]) * scale);
        a->fall[k] = static_cast<float>(static_cast<double>(h.fall[k]) * scale);
    }
    for (int k = 0; k < V::NCEP; ++k)
        for (int m = 0; m < V::NMEL; ++m) a->dct[k * V::NMEL + m] = h.dct[static_cast<size_t>(k) * V::NMEL + m];
    a->logmel = p.output == MFCC_OUT_LOGMEL;
    a->preemph = p.preemph;
    a->log_floor = p.log_floor;
}

template <typename PcmT, class V>
int launch_variant(const SpState *st, const Tile *d_tiles, int64_t n_tiles, const PcmT *d_pcm, float *d_out,
                   cudaStream_t stream)
{
    using G = Geo<V>;
    constexpr size_t smem = sizeof(float) * G::SMEM_FLOATS;
    auto kern = fused_sp_kernel<PcmT, V>;
    static thread_local const void *configured = nullptr;
    if (configured != reinterpret_cast<const void *>(kern)) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) !=
            cudaSuccess) {
            cudaGetLastError();
            return MFCC_ECUDA;
        }
        configured = reinterpret_cast<const void *>(kern);
    }
    SpArgs<V> a;
    std::memcpy(&a, st->args.data(), sizeof(a));
    a.tiles = d_tiles;
    a.n_tiles = n_tiles;
    a.out = d_out;
    a.tab = st->d_tab;
    const int per_sm = (smem + 1024) * 2 <= 228 * 1024 ? 2 : 1;
    const int64_t grid = std::min<int64_t>(n_tiles, static_cast<int64_t>(st->sm_count) * per_sm);
    kern<<<static_cast<unsigned>(grid), kThreads, smem, stream>>>(d_pcm, a);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError() == cudaSuccess ? MFCC_OK : MFCC_ECUDA;
}

}  // namespace

const char *sp_match(const mfcc_params &p, const HostTables &h)
{
    if (structure_matches<Var16k>(p, h)) return Var16k::name();
    if (structure_matches<Var8k>(p, h)) return Var8k::name();
    return nullptr;
}

int sp_prepare(mfcc_plan *plan)
{
    SpState *st = new SpState();
    std::vector<float> tab;
    if (structure_matches<Var16k>(plan->p, plan->host)) {
        st->variant = 0;
        build_tab_and_args<Var16k>(plan, tab, st->args);
    } else if (structure_matches<Var8k>(plan->p, plan->host)) {
        st->variant = 1;
        build_tab_and_args<Var8k>(plan, tab, st->args);
    } else {
        delete st;
        return MFCC_ENOTSUP;
    }
    st->sm_count = plan->sm_count;
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
    plan->sp_state = st;
    return MFCC_OK;
}

void sp_release(mfcc_plan *plan)
{
    SpState *st = static_cast<SpState *>(plan->sp_state);
    if (st == nullptr) return;
    if (st->d_tab) cudaFree(st->d_tab);
    delete st;
    plan->sp_state = nullptr;
}

template <typename PcmT>
int sp_launch(const mfcc_plan *plan, const Tile *d_tiles, int64_t n_tiles, const PcmT *d_pcm, float *d_out,
              cudaStream_t stream)
{
    const SpState *st = static_cast<const SpState *>(plan->sp_state);
    if (st == nullptr) return MFCC_ENOTSUP;
    if (st->variant == 0) return launch_variant<PcmT, Var16k>(st, d_tiles, n_tiles, d_pcm, d_out, stream);
    return launch_variant<PcmT, Var8k>(st, d_tiles, n_tiles, d_pcm, d_out, stream);
}

template int sp_launch<int16_t>(const mfcc_plan *, const Tile *, int64_t, const int16_t *, float *, cudaStream_t);
template int sp_launch<float>(const mfcc_plan *, const Tile *, int64_t, const float *, float *, cudaStream_t);

}  // namespace mfcc
