// mfcc_fused_wide.cu — fused tile kernel for the LARGE transform (BASELINE.json configs[3]: 48 kHz,
// 1200-sample frames, 2048-point FFT, 80 mel bands, 40 cepstra): the "large-FFT, smem-pressure path".
//
// A 2048-point frame needs 8 KB of f32 workspace between the two FFT passes, so the 32-frames-per-tile,
// lane = frame layout of mfcc_fused_sp.cu (64 KB for 512 points) cannot hold it.  Here a tile is F = 8
// frames and a warp carries 4 work items at once: lane = sub * 8 + frame, sub = 0 .. 3.  Constants a
// butterfly needs are then uniform per quarter-warp instead of per warp (shared-memory loads with 4
// distinct addresses), every array is [...][8 frames], and the layouts are padded so that the four
// quarter-warps of a load or store fall into different banks:
//   staged  lane stride HOP + 4 words (= 4 mod 32): 8 frames x 4 consecutive columns = 32 distinct banks per scalar load
//   ws      [row k1 - 1][column, bit 0 flipped when bit 1 is set][8] float2, rows padded by 16 words
//   P       [bin][8]: four consecutive bins per warp store
// One CTA per SM = two independent halves of 8 warps (named barriers, 128 registers, 16 resident warps), 32 work
// slots each (slot = warp * 4 + sub):
//   S0 stage     PCM -> f32 -> pre-emphasis -> staged (each sample read from HBM once per tile; the next unit's PCM
//                arrives by bulk async copy into the idle part of the workspace)
//   S1 pass 1    slot = column a: windowed REAL DFT-64 over b (38 live rows), inter-pass twiddle
//   S2 pass 2    slot = row k1 - 1, k1 = 1 .. 32: complex DFT-32 over a; row 0 (real) through a real DFT-32 by
//                slot 0; |X|^2 -> P
//   S3           slot s walks mel segments s, s + 32, s + 64: S / T sums per segment -> rise / fall halves
//   S3b          band = rise + fall of the neighbouring segments -> log (and the frame energy term)
//   S4           thread (slot, frame) forms cepstra slot and slot + 32 from the mirrored band pairs; store
// N = RB * RA = 64 * 32, n = a + 32 b, k = k1 + 64 k2 (mfcc_rfft.cuh; tests/codelets checks the data flow).
//
// No reference code corresponds to this (SURVEY.md §8a "Ref file:line = none").
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

// FMA-fused twiddled butterflies in the codelets (mfcc_rfft.cuh): measured +2.7 % on configs[3] for this kernel (the
// 64-point real and 32-point complex transforms carry 45 constant twiddles each), parity-tested with it; the 512/256-point
// kernel gains 0.4 % / loses 0.3 % and keeps the product form.
#ifndef MFCC_RFFT_FUSED
#define MFCC_RFFT_FUSED 1
#endif
#include "mfcc_rfft.cuh"
#include "mfcc_host.h"

// Mutual exclusion of the two halves of a CTA on a span of phases (see MFCC_SP_LOCK_* in mfcc_fused_sp.cu): 10 * from + to,
// the half is held at barrier B<from> (1 = B1, 2 = B2, 3 = B3, 4 = B4a, 5 = B4) until it owns the lock and gives it back
// after barrier B<to>; 0 = none.  Measured with 16 warps (one box, tools/time_variants.py): pass 1 exclusive (12) +6.3 %,
// pass 2 exclusive (23) +4.6 %, S3 (34) +3.0 %, S4 + S0 (51) +3.5 %.
#ifndef MFCC_WIDE_LOCK
#define MFCC_WIDE_LOCK 12
#endif
// Two bank-conflict fixes found with the per-line "excessive wavefronts" of an ncu capture (profiles/r2_wide_C.md), each a
// build switch for A/B timing:
//   MFCC_WIDE_VMPAD   words between the two folded log-energy arrays of S3b / S4.  The even and the odd cepstra read v+ and
//                     v- at the same index; with the arrays a multiple of 32 words apart (hmp * F = 320 for 80 bands) the two
//                     halves of a warp met on the same banks: 9.6 M of the 27.8 M excessive wavefronts of a launch.
//   MFCC_WIDE_S0SWAP  lanes 4 .. 7 of every eight write the two 16-byte halves of their staged chunk in the opposite order.
//                     Chunks are 32 bytes apart, so a quarter-warp's eight 16-byte stores covered only half of the banks
//                     (2-way conflict on both stores: 7.9 M excessive wavefronts); swapped, they cover all 32.
#ifndef MFCC_WIDE_VMPAD
#define MFCC_WIDE_VMPAD 8
#endif
#ifndef MFCC_WIDE_S0SWAP
#define MFCC_WIDE_S0SWAP 1
#endif
//   MFCC_WIDE_SEGPERM the walk order of S3 is permuted on the host so that the four segments a warp walks at once start on bins
//                     that differ modulo 4 (see wide_prepare)
#ifndef MFCC_WIDE_SEGPERM
#define MFCC_WIDE_SEGPERM 1
#endif
// Poison build (see mfcc_fused_sp.cu): NaN-fill every aliased buffer at the point where the comments say it is dead.
#ifndef MFCC_POISON
#define MFCC_POISON 0
#endif

namespace mfcc {

namespace {

constexpr int F = 8;                      // frames per tile
constexpr int SUBS = 32 / F;              // work items per warp
// 8 warps per half = 32 slots: a slot carries ONE column in pass 1 and ONE row in pass 2, so the codelets fit 128 registers
// (the 64-point real and the 32-point complex transform compile to 72 and 128 registers on their own) and the SM holds 16
// warps.  Round 1 had 4 warps per half, a column PAIR and two rows per slot: 255 registers, 8 resident warps, issue 53 % —
// 275.6 M against 310.1 M frames/s on the same box (profiles/r2_wide_16warps.md).
constexpr int kWarps = 8;                 // per half
constexpr int kSlots = kWarps * SUBS;     // 32
constexpr int kHalfThreads = kWarps * 32;
constexpr int kThreads = 2 * kHalfThreads;
constexpr int kPad = 4;
constexpr int KC = 64;                    // cepstra per frame (n_cep <= KC): thread (slot, frame) forms k = slot and k = slot + 32
constexpr size_t kSmemMax = 227 * 1024;

template <int L_, int HOP_>
struct Geo {
    static constexpr int L = L_, HOP = HOP_, RB = 64, RA = 32;
    static constexpr int NFFT = RB * RA, NB = NFFT / 2 + 1, H = RB / 2;
    static constexpr int NZ = (L + RA - 1) / RA;
    static constexpr int NZP = (NZ + 3) / 4 * 4;            // window rows per column, padded for 16-byte loads
    static constexpr int STRIDE = HOP + kPad;
    static constexpr int padded(int i) { return i + kPad * (i / HOP); }
    static constexpr int SLACK = RA * NZ - L;
    static constexpr int tceil(int n_frames) { return ((n_frames - 1) * HOP + L + SLACK + 7) / 8 * 8; }
    static constexpr int tceil_s(int n_frames, int shift) { return (shift + (n_frames - 1) * HOP + L + SLACK + 7) / 8 * 8; }
    static constexpr int TCEIL = tceil(F) + 8;              // + up to 7 samples of alignment shift
    static constexpr int STAGED = padded(TCEIL) + 8;
    static constexpr int PW = (NB + 7) * F;                 // 7 zeroed slack rows for the tail's 4-bin reads
    static constexpr int UNION = ((STAGED > PW ? STAGED : PW) + 3) / 4 * 4;
    static constexpr int WSROW = RA * F * 2 + 16;           // floats per complex row, padded
    static constexpr int WS = H * WSROW;                    // rows k1 = 1 .. H
    static constexpr int R0 = RA * F;                       // row 0 (real)
    static constexpr int RAWOFF = 8192;                     // raw PCM of the NEXT unit lands in the workspace past the tail scratch
    static constexpr int RAW = (TCEIL + 16) / 2;            // floats holding 8 lead + TCEIL + 1 look-ahead int16 samples
    static constexpr int HALF = UNION + WS + R0 + 4;        // + mbarrier (8 B in a 16-B slot)
    static constexpr int T_WIN = 0;                         // [RA][NZP] float: the window values of column a, row b
    // The four quarter-warps of a pass-1 load read four DIFFERENT columns' rows: a row stride of 2 H = 64 floats would put all
    // four on the same banks (4 wavefronts per LDS.128: 1,536 of the 18,600 wavefronts per 32 frames); 4 floats of padding per
    // row (stride = 4 mod 32) separates them.  MFCC_WIDE_TWPAD 0 restores the unpadded table for A/B timing.
#ifndef MFCC_WIDE_TWPAD
#define MFCC_WIDE_TWPAD 4
#endif
    static constexpr int TWS = 2 * H + MFCC_WIDE_TWPAD;     // floats per twiddle row
    static constexpr int T_TW = T_WIN + RA * NZP;           // [RA][TWS]: float2 W_N^(a k1), k1 = 1 .. H at slot k1 - 1
    static constexpr int TABF = T_TW + RA * TWS;
    static constexpr int pcol(int c) { return c ^ ((c >> 1) & 1); }
    static_assert(STRIDE % 32 == 4, "quarter-warp bank separation of the staged tile");
    static_assert(HOP % 8 == 0 && HOP % RA == 0, "chunks and column pairs must not straddle a hop block");
    static_assert(RA == kSlots, "one column per slot");
    static_assert(H == kSlots, "one complex row per slot");
    static_assert(L <= NFFT && NZ <= RB, "frame does not fit the transform");
    static_assert(TABF % 4 == 0 && UNION % 4 == 0 && WS % 4 == 0, "16-byte aligned regions");
    static_assert(129 * F <= RAWOFF && RAWOFF + RAW <= WS && RAWOFF % 4 == 0,
                  "tail scratch, then the raw PCM buffer, must fit in the workspace");
    static_assert(SLACK + 8 <= kTileSpanSlack, "the host's span check (Tile::flags) must cover the staged span");
};

struct WideLayout {
    int seg;      // float4 per walk position: {first bin * F (int), width w (int), s = 1 / (w NFFT), segment index (int)}
    int dct;      // [32 slots][hmp] d[s][m], then [n2 slots][hmp] d[s + 32][m]; m < hmp = ceil(n_mel / 2) rounded up to
                  // even (the mirrored half follows from the symmetry); zero past n_cep
    int total;
};

struct WideArgs {
    const Tile *tiles;
    int64_t n_tiles;
    float *out;
    const float *tab;
    WideLayout lay;
    int n_mel, n_cep, logmel;
    int energy, od;       // MFCC_ENERGY_*; floats per output row
    int nseg;             // segments walked: n_mel + 1, + the two pseudo-segments outside the filterbank when the energy term is on
    float inv_n;          // 1 / NFFT
    int ls, mel_magic, hmp;
    int n2;               // slots that form a second cepstrum (k = slot + 32): n_cep - 32 rounded up to a multiple of 4, or 0
    int rf;               // scratch offset (floats) of the per-segment rise / fall sums [n_mel + 3][F] x 2
    int ef;               // scratch offset of the log frame energy [F]
    float preemph, log_floor;
};

__device__ __forceinline__ float2 lds_f2(const float *p) { return *reinterpret_cast<const float2 *>(p); }
__device__ __forceinline__ float4 lds_f4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
// int16 pair -> two exact floats: (v ^ 0x8000) = v + 32768 in the mantissa of 2^23; one XOR, one PRMT per float
__device__ __forceinline__ float2 s16x2_to_f32(uint32_t w)
{
    const uint32_t b = w ^ 0x80008000u;
    const uint32_t lo = __byte_perm(b, 0x4B000000u, 0x7610);
    const uint32_t hi = __byte_perm(b, 0x4B000000u, 0x7632);
    return make_float2(__uint_as_float(lo) - 8421376.0f, __uint_as_float(hi) - 8421376.0f);
}
// log2 of a NORMAL positive float (callers clamp to log_floor first): one MUFU.LG2, no denormal rescaling
__device__ __forceinline__ float lg2_fast(float x)
{
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
constexpr float kLn2 = 0.69314718055994531f;
__device__ __forceinline__ float to_f32(int16_t v) { return static_cast<float>(v); }
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ void half_sync(int half)
{
    asm volatile("bar.sync %0, %1;" ::"r"(half + 1), "n"(kHalfThreads) : "memory");
}
// ---- mbarrier + bulk async copy (TMA engine, 1-D), as in mfcc_fused_sp.cu ----
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
// orders the half's earlier generic-proxy accesses to the workspace before the async-proxy write of the copy
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ float pwr(const rf::cplx &z) { return fmaf(z.re, z.re, z.im * z.im); }

// MEL > 0: the plan has exactly MEL (even) bands and cepstral output, so the DCT loop of S4 is unrolled; MEL = 0: any plan.
template <typename PcmT, int L_, int HOP_, int MEL>
__global__ void __launch_bounds__(kThreads, 1) fused_wide_kernel(const PcmT *__restrict__ pcm, const WideArgs a)
{
    using G = Geo<L_, HOP_>;
    constexpr int HOP = G::HOP, RB = G::RB, RA = G::RA, STRIDE = G::STRIDE, H = G::H, NZ = G::NZ;
    extern __shared__ __align__(16) float smem[];
    const int half = threadIdx.x / kHalfThreads, tid = threadIdx.x % kHalfThreads;
    const int lane = tid & 31, warp = tid >> 5;
    const int f = lane % F, sub = lane / F, slot = warp * SUBS + sub;

    float *tab = smem;
    float *mine = smem + a.lay.total + half * G::HALF;
    float *staged = mine;                 // S0-S1
    float *pw = staged;                   // S2-S3 (aliases staged)
    float *ws = mine + G::UNION;          // complex rows
    float *r0row = ws + G::WS;            // row 0 (real) [RA][F]
    float *scr = ws;                      // tail scratch (aliases ws)
    // The NEXT unit's PCM is fetched by one bulk async copy into the part of the workspace the tail does not
    // use, issued once pass 2 has released the workspace (B3) and consumed by the next S0 before S1 rewrites it.
    const int16_t *raw16 = reinterpret_cast<const int16_t *>(ws + G::RAWOFF);
    const uint32_t raw_s = smem_u32(raw16), bar = smem_u32(r0row + G::R0);
    if (tid == 0) mbar_init(bar, 1);
    constexpr int kLockFrom = MFCC_WIDE_LOCK / 10, kLockTo = MFCC_WIDE_LOCK % 10;
    // (one word next to half 0's mbarrier, which uses 8 bytes of its 16-byte slot)
    [[maybe_unused]] int *phase_lock = reinterpret_cast<int *>(smem + a.lay.total + G::UNION + G::WS + G::R0 + 2);
    if (MFCC_WIDE_LOCK != 0 && threadIdx.x == 0) *phase_lock = 0;
    bool held = false;
    [[maybe_unused]] auto lock_at = [&](int point) {
        if (MFCC_WIDE_LOCK != 0 && point == kLockFrom && tid == 0) {
            while (atomicCAS(phase_lock, 0, 1) != 0) __nanosleep(32);
            held = true;
        }
    };
    [[maybe_unused]] auto unlock_at = [&](int point) {
        if (MFCC_WIDE_LOCK != 0 && point == kLockTo && held) {
            atomicExch(phase_lock, 0);
            held = false;
        }
    };
    [[maybe_unused]] auto poison = [&](float *p, int n) {
        for (int i = tid; i < n; i += kHalfThreads) p[i] = __int_as_float(0x7fc00000);
    };

    for (int i = G::NB * F + tid; i < G::PW; i += kHalfThreads) pw[i] = 0.0f;   // slack rows stay zero
    for (int i = threadIdx.x * 4; i < a.lay.total; i += kThreads * 4)
        *reinterpret_cast<float4 *>(tab + i) = __ldg(reinterpret_cast<const float4 *>(a.tab + i));
    const float *t_win = tab + G::T_WIN, *t_tw = tab + G::T_TW;
    const float4 *t_seg = reinterpret_cast<const float4 *>(tab + a.lay.seg);
    const float *t_dct = tab + a.lay.dct;
    __syncthreads();

    // The batch's tile table holds 32-frame tiles (mfcc_host.h kTileFrames); a work unit here is a quarter of one.
    constexpr int kQ = kTileFrames / F;
    const int64_t n_units = a.n_tiles * kQ;
    const int64_t first = 2 * static_cast<int64_t>(blockIdx.x) + half, step = 2 * static_cast<int64_t>(gridDim.x);
    const bool base_ok = sizeof(PcmT) == 2 && (reinterpret_cast<uintptr_t>(pcm) & 15) == 0;
    // unit uu of tile tl: its first sample and frame count; "vector-loadable" = frames inside the utterance and
    // staged span inside the array (Tile::flags, checked on the host), whatever the utterance's alignment
    auto unit_geom = [&](const Tile &tl, int64_t uu, int64_t &fs, int &nf) -> bool {
        const int qq = static_cast<int>(uu % kQ);
        nf = min(F, tl.n_frames - qq * F);
        fs = tl.first_sample + static_cast<int64_t>(qq) * F * HOP;
        return base_ok && nf > 0 && (tl.flags & kTileInside) != 0;
    };
    // raw16[8 + i] = x[o + i], o = the 8-sample boundary below fs; the 8 samples before o ride along when they exist
    auto issue_copy = [&](int64_t fs, int nf) {
        const int s = static_cast<int>(fs & 7);
        const int64_t o = fs - s;
        const int lead = o >= 8 ? 8 : 0;
        const uint32_t bytes = static_cast<uint32_t>(lead + G::tceil_s(nf, s)) * 2u;
        fence_proxy_async();
        mbar_expect_tx(bar, bytes);
        bulk_g2s(raw_s + (8 - lead) * 2, pcm + o - lead, bytes, bar);
    };
    Tile nt{};
    bool copied = false;          // a bulk copy for the upcoming unit is in flight (same value in every thread of the half)
    uint32_t phase = 0;
    if (first < n_units) {
        nt = a.tiles[first / kQ];
        int64_t fs;
        int nf;
        copied = unit_geom(nt, first, fs, nf);
        if (copied && tid == 0) issue_copy(fs, nf);
    }
    for (int64_t u = first; u < n_units; u += step) {
        Tile tile = nt;
        if (u + step < n_units) nt = a.tiles[(u + step) / kQ];   // arrives during the unit
        const int q = static_cast<int>(u % kQ);
        const int n_frames = min(F, tile.n_frames - q * F);
        if (n_frames <= 0) continue;                              // (uniform for the half: no barrier is skipped by part of it; no copy was issued for it)
        tile.first_sample += static_cast<int64_t>(q) * F * HOP;
        tile.out_row += q * F;
        const int tc = G::tceil(n_frames);
        const bool have_raw = copied;
        copied = false;

        // ---- S0: stage y[n] = x[n] - a x[n-1] once per sample.  A tile whose frames lie inside the utterance is
        // read with 16-byte loads from the 8-sample boundary o below its first sample, whatever the alignment of
        // the utterance: the shift s = first_sample - o = e + d is absorbed by the layout (staged index i holds
        // sample o + d + i, kPad words inserted at i = e + k HOP, so frame f starts at word e + f STRIDE) ----
        int e = 0;
        {
            bool fast = false;
            const int sh = static_cast<int>(tile.first_sample & 7);
            const int64_t o = tile.first_sample - sh;
            if constexpr (sizeof(PcmT) == 2)   // frames inside the utterance, span inside the array: Tile::flags (host)
                fast = ((reinterpret_cast<uintptr_t>(pcm) & 15) == 0) && (tile.flags & kTileInside) != 0;
            if (fast) {
                e = sh & 6;
                const int d = sh & 1;
                // samples from o on: in the raw buffer when the copy was issued a unit ahead, else straight from HBM
                if (have_raw) {
                    mbar_wait(bar, phase);
                    phase ^= 1u;
                }
                const int16_t *x = have_raw ? raw16 + 8 : reinterpret_cast<const int16_t *>(pcm) + o;
                const float na = -a.preemph;
                const int nchunks = G::tceil_s(n_frames, sh) >> 3;
                for (int c = tid; c < nchunks; c += kHalfThreads) {
                    const uint4 q = *(reinterpret_cast<const uint4 *>(x) + c);
                    // d = 0: the sample before the chunk; d = 1: the sample after it
                    // (in bounds: only x[-1] of a tile at the very start of the array does not exist; the sample
                    // after the last chunk lies inside the span the host checked, kTileSpanSlack)
                    uint32_t pv = (d || c > 0 || o > 0 || have_raw) ? static_cast<uint16_t>(x[8 * c + (d ? 8 : -1)]) : 0u;
                    uint32_t w0 = q.x, w1 = q.y, w2 = q.z, w3 = q.w;
                    if (d) {   // odd shift: move the chunk down one half-word; its old first sample becomes the predecessor
                        const uint32_t first = w0;
                        w0 = __byte_perm(w0, w1, 0x5432);
                        w1 = __byte_perm(w1, w2, 0x5432);
                        w2 = __byte_perm(w2, w3, 0x5432);
                        w3 = __byte_perm(w3, pv, 0x5432);
                        pv = first;
                    }
                    const float xp = s16x2_to_f32(pv).x;
                    const float2 x01 = s16x2_to_f32(w0), x23 = s16x2_to_f32(w1);
                    const float2 x45 = s16x2_to_f32(w2), x67 = s16x2_to_f32(w3);
                    const float4 lo4 = make_float4(fmaf(na, xp, x01.x), fmaf(na, x01.x, x01.y),
                                                   fmaf(na, x01.y, x23.x), fmaf(na, x23.x, x23.y));
                    const float4 hi4 = make_float4(fmaf(na, x23.y, x45.x), fmaf(na, x45.x, x45.y),
                                                   fmaf(na, x45.y, x67.x), fmaf(na, x67.x, x67.y));
                    // chunk c = (HOP / 8) k + m lies in hop block k, except the words below e of the chunks with
                    // m = 0, which still belong to block k - 1 (one chunk in HOP / 8 takes the split stores)
                    const int k = c / (HOP / 8);
                    float *dst = staged + 8 * c + kPad * k;
                    if (e != 0 && k > 0 && c == k * (HOP / 8)) {
                        float *dl = dst - kPad;
                        *reinterpret_cast<float2 *>((0 < e ? dl : dst) + 0) = make_float2(lo4.x, lo4.y);
                        *reinterpret_cast<float2 *>((2 < e ? dl : dst) + 2) = make_float2(lo4.z, lo4.w);
                        *reinterpret_cast<float2 *>((4 < e ? dl : dst) + 4) = make_float2(hi4.x, hi4.y);
                        *reinterpret_cast<float2 *>(dst + 6) = make_float2(hi4.z, hi4.w);
                    } else if (MFCC_WIDE_S0SWAP) {
                        const bool sw = (lane & 4) != 0;
                        const float4 q0 = sw ? hi4 : lo4, q1 = sw ? lo4 : hi4;
                        *reinterpret_cast<float4 *>(dst + (sw ? 4 : 0)) = q0;
                        *reinterpret_cast<float4 *>(dst + (sw ? 0 : 4)) = q1;
                    } else {
                        *reinterpret_cast<float4 *>(dst) = lo4;
                        *reinterpret_cast<float4 *>(dst + 4) = hi4;
                    }
                    // the utterance's first sample has no predecessor: y = x (word e of chunk 0, written just above)
                    if (c == 0 && tile.first_sample == tile.utt_begin) staged[e] = to_f32(x[sh]);
                }
            } else if (sizeof(PcmT) == 4 && (reinterpret_cast<uintptr_t>(pcm) & 15) == 0 && (tile.flags & kTileInside) != 0) {
                // f32 PCM: the same any-alignment staging with 16-byte loads straight from HBM (no raw buffer)
                if constexpr (sizeof(PcmT) == 4) {
                    e = sh & 6;
                    const int d = sh & 1;
                    const float *x = reinterpret_cast<const float *>(pcm) + o;   // x[i] = sample o + i
                    const float na = -a.preemph;
                    const int nchunks = G::tceil_s(n_frames, sh) >> 3;
                    for (int c = tid; c < nchunks; c += kHalfThreads) {
                        const float4 lo4 = __ldg(reinterpret_cast<const float4 *>(x) + 2 * c);
                        const float4 hi4 = __ldg(reinterpret_cast<const float4 *>(x) + 2 * c + 1);
                        const float nb = (d || c > 0 || o > 0) ? __ldg(x + 8 * c + (d ? 8 : -1)) : 0.0f;
                        const float v0 = d ? lo4.x : nb, v1 = d ? lo4.y : lo4.x, v2 = d ? lo4.z : lo4.y, v3 = d ? lo4.w : lo4.z;
                        const float v4 = d ? hi4.x : lo4.w, v5 = d ? hi4.y : hi4.x, v6 = d ? hi4.z : hi4.y, v7 = d ? hi4.w : hi4.z;
                        const float v8 = d ? nb : hi4.w;    // staged word j = v[j + 1] - a v[j]
                        const float4 y0 = make_float4(fmaf(na, v0, v1), fmaf(na, v1, v2), fmaf(na, v2, v3), fmaf(na, v3, v4));
                        const float4 y1 = make_float4(fmaf(na, v4, v5), fmaf(na, v5, v6), fmaf(na, v6, v7), fmaf(na, v7, v8));
                        const int k = c / (HOP / 8);
                        float *dst = staged + 8 * c + kPad * k;
                        if (e != 0 && k > 0 && c == k * (HOP / 8)) {
                            float *dl = dst - kPad;
                            *reinterpret_cast<float2 *>((0 < e ? dl : dst) + 0) = make_float2(y0.x, y0.y);
                            *reinterpret_cast<float2 *>((2 < e ? dl : dst) + 2) = make_float2(y0.z, y0.w);
                            *reinterpret_cast<float2 *>((4 < e ? dl : dst) + 4) = make_float2(y1.x, y1.y);
                            *reinterpret_cast<float2 *>(dst + 6) = make_float2(y1.z, y1.w);
                        } else {
                            *reinterpret_cast<float4 *>(dst) = y0;
                            *reinterpret_cast<float4 *>(dst + 4) = y1;
                        }
                        if (c == 0 && tile.first_sample == tile.utt_begin) staged[e] = x[sh];
                    }
                }
            } else {
                const int64_t room_lo = tile.first_sample - tile.utt_begin;
                const int64_t room_hi = tile.utt_end - tile.first_sample;
                const PcmT *x = pcm + tile.first_sample;
                for (int i = tid; i < tc; i += kHalfThreads) {
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
        lock_at(1);
        half_sync(half);   // B1: staged complete; previous tile's scratch free
        unlock_at(1);
        if constexpr (MFCC_POISON) {   // the raw PCM has been consumed and the tail scratch is dead: pass 1 rewrites the workspace
            poison(ws, G::WS + G::R0);
            half_sync(half);
        }

        // ---- S1: pass 1.  Slot = column a: windowed real DFT-64 over b, inter-pass twiddle ----
        {
            {
                // 38 scalar loads (lane stride 4 words + the slot: 32 distinct banks), window from the
                // column's own row of the table (four rows per 16-byte load).  (The window rows as kernel parameters — what gave
                // the 512-point kernel 5 % — were measured here too: a warp reads four different columns, so the loads are
                // indexed, divergent constant loads, and the kernel runs at 0.62 x: 0.334 -> 0.207 G frames/s.)
                const int col = slot;
                const float *base = staged + e + f * STRIDE + col;
                const float *wrow = t_win + col * G::NZP;
                float x[RB];
#pragma unroll
                for (int b = 0; b < RB; b += 4) {
                    if (b < NZ) {
                        const float4 w = lds_f4(wrow + b);
                        x[b] = base[G::padded(RA * b)] * w.x;
                        x[b + 1] = b + 1 < NZ ? base[G::padded(RA * (b + 1))] * w.y : 0.0f;
                        x[b + 2] = b + 2 < NZ ? base[G::padded(RA * (b + 2))] * w.z : 0.0f;
                        x[b + 3] = b + 3 < NZ ? base[G::padded(RA * (b + 3))] * w.w : 0.0f;
                    } else {
                        x[b] = x[b + 1] = x[b + 2] = x[b + 3] = 0.0f;
                    }
                }
                rf::cplx X[H + 1];
                rf::rdft64<NZ>(x, X);
                r0row[col * F + f] = X[0].re;
                float *wsa = ws + ((col ^ ((col >> 1) & 1)) * F + f) * 2;   // column slot pcol(col)
                const float *trow = t_tw + col * G::TWS;
#pragma unroll
                for (int k1 = 1; k1 <= H; k1 += 2) {
                    const float4 tw = lds_f4(trow + 2 * (k1 - 1));       // twiddles of k1, k1 + 1
                    const rf::cplx v = rf::cmulc(X[k1], tw.x, tw.y);
                    *reinterpret_cast<float2 *>(wsa + (k1 - 1) * G::WSROW) = make_float2(v.re, v.im);
                    const rf::cplx u = rf::cmulc(X[k1 + 1], tw.z, tw.w);
                    *reinterpret_cast<float2 *>(wsa + k1 * G::WSROW) = make_float2(u.re, u.im);
                }
            }
        }
        lock_at(2);
        half_sync(half);   // B2
        unlock_at(2);
        if constexpr (MFCC_POISON) {   // the staged samples are dead: pass 2 writes P over them (slack rows stay zero)
            poison(staged, G::UNION);
            half_sync(half);
            for (int i = G::NB * F + tid; i < G::PW; i += kHalfThreads) pw[i] = 0.0f;
            half_sync(half);
        }

        // ---- S2: pass 2.  Item = row k1 = 1 .. 32: complex DFT-32 over a gives bins k1 + 64 k2; power ----
        {
#pragma unroll 1
            {
                const int k1 = slot + 1;
                const float *row = ws + (k1 - 1) * G::WSROW + f * 2;
                rf::cplx z[RA];
#pragma unroll
                for (int c = 0; c < RA; ++c) {
                    const float2 p = lds_f2(row + G::pcol(c) * F * 2);
                    z[c] = rf::cplx{p.x, p.y};
                }
                rf::cdft32(z);
                float *p_lo = pw + k1 * F + f;                    // bins k1 + 64 k2, k2 < 16
#pragma unroll
                for (int k2 = 0; k2 < RA / 2; ++k2) p_lo[RB * k2 * F] = pwr(z[k2]);
                if (k1 < H) {                                     // row 32 mirrors onto itself
                    float *p_hi = pw + (RB - k1) * F + f;         // N - k = (64 - k1) + 64 (31 - k2)
#pragma unroll
                    for (int k2 = RA / 2; k2 < RA; ++k2) p_hi[RB * (RA - 1 - k2) * F] = pwr(z[k2]);
                }
            }
            if (slot == 0) {
                // row 0 is real: real DFT-32 over a gives bins 64 k2, k2 = 0 .. 16
                float x[RA];
#pragma unroll
                for (int c = 0; c < RA; ++c) x[c] = r0row[c * F + f];
                rf::cplx X0[RA / 2 + 1];
                rf::rdft32<32>(x, X0);
                pw[f] = X0[0].re * X0[0].re;
                pw[RB * (RA / 2) * F + f] = X0[RA / 2].re * X0[RA / 2].re;
#pragma unroll
                for (int k2 = 1; k2 < RA / 2; ++k2) pw[RB * k2 * F + f] = pwr(X0[k2]);
            }
        }
        lock_at(3);
        half_sync(half);   // B3: P complete, workspace free
        unlock_at(3);
        if constexpr (MFCC_POISON) {   // both the tail scratch and the raw buffer of the next unit start from dead memory
            poison(ws, G::WS + G::R0);
            half_sync(half);
        }
        if (u + step < n_units) {
            int64_t fs;
            int nf;
            copied = unit_geom(nt, u + step, fs, nf);
            if (copied && tid == 0) issue_copy(fs, nf);
        }

        // ---- S3: filterbank sums (see mfcc_fused_sp.cu S3): per segment S = sum P, T = sum i P give the rise into
        // filter j, s T, and the fall out of filter j - 1, s (w S - T).  Slot s takes segments s, s + 32, s + 64:
        // the four slots of a warp then walk NEIGHBOURING segments of nearly equal width in every round, so their
        // loop counts agree (with a contiguous filter group per slot the warp ran the longest of four different
        // walks).  rise[j][frame] and fall[j][frame] go through the scratch; S3b adds the two halves of each band.
        // Measured and dropped in round 2 (profiles/r2_wide_16warps.md): segments dealt out by width in a snake over the
        // slots (equal bin totals per slot: -0.6 %, more index arithmetic than the balance gains) and a warp-cooperative walk
        // (lane = sub * 8 + frame takes every fourth bin, two shuffles add the partial sums: -19 %, ten segments per warp
        // one after the other, each paying the descriptor-load and shuffle latencies, instead of four in flight). ----
        {
            float *rise = scr + a.rf + f, *fall = rise + (a.n_mel + 3) * F;
#pragma unroll 1
            for (int j = slot; j < a.nseg; j += kSlots) {
                const float4 sg = t_seg[j];
                const float *p = pw + __float_as_int(sg.x) + f;
                const int w = __float_as_int(sg.y);
                float S = 0.0f, T = 0.0f, U = 0.0f, i0 = 0.0f;
                int c = w >> 2;
#pragma unroll 1
                for (; c >= 2; c -= 2) {                    // 8 bins per step: all loads first, two independent chains
                    const float a0 = p[0], a1 = p[F], a2 = p[2 * F], a3 = p[3 * F];
                    const float b0 = p[4 * F], b1 = p[5 * F], b2 = p[6 * F], b3 = p[7 * F];
                    const float sa = (a0 + a1) + (a2 + a3), sb = (b0 + b1) + (b2 + b3);
                    T = fmaf(i0, sa, T) + fmaf(3.0f, a3, fmaf(2.0f, a2, a1));
                    U = fmaf(i0 + 4.0f, sb, U) + fmaf(3.0f, b3, fmaf(2.0f, b2, b1));
                    S += sa + sb;
                    i0 += 8.0f;
                    p += 8 * F;
                }
                if (c) {
                    const float a0 = p[0], a1 = p[F], a2 = p[2 * F], a3 = p[3 * F];
                    const float sa = (a0 + a1) + (a2 + a3);
                    T = fmaf(i0, sa, T) + fmaf(3.0f, a3, fmaf(2.0f, a2, a1));
                    S += sa;
                    i0 += 4.0f;
                    p += 4 * F;
                }
                T += U;
                if (const int lo = w & 3) {                 // 1..3 leftover bins, no loop: the rows past the segment are
                    const float a0 = p[0];                  // finite (next segment or the zeroed slack rows) and deselected
                    const float a1 = lo > 1 ? p[F] : 0.0f;
                    const float a2 = lo > 2 ? p[2 * F] : 0.0f;
                    const float sa = (a0 + a1) + a2;
                    T = fmaf(i0, sa, T) + fmaf(2.0f, a2, a1);
                    S += sa;
                }
                const float rs = sg.z * T;                  // s = 1 / (w NFFT); pseudo-segments (energy term): s = 0
                const int row = __float_as_int(sg.w);      // the segment this walk position holds (the walk order is permuted)
                rise[row * F] = rs;
                fall[row * F] = fmaf(a.inv_n, S, -rs);
            }
        }
        lock_at(4);
        half_sync(half);   // B4a: every segment's two sums are in the scratch
        unlock_at(4);
        if constexpr (MFCC_POISON) {   // P is dead: the next S0 stages over it
            poison(pw, G::NB * F);
            half_sync(half);
        }
        // ---- S3b: band m = rise of segment m + fall of segment m + 1, log.  Log-mel output: the frame's row, staged for a
        // coalesced copy.  Cepstra: the DCT symmetry d[k][M-1-m] = (-1)^k d[k][m] is applied HERE, once per frame: thread
        // (q, frame) forms both logs of the mirrored pair and stores vp[q] = lg[q] + lg[M-1-q] (even cepstra) and
        // vm[q] = lg[q] - lg[M-1-q] (odd cepstra); S4 then needs one load per term instead of re-folding the pair in
        // every one of its 32 slots. ----
        {
            const float *rise = scr + a.rf, *fall = rise + (a.n_mel + 3) * F;
            if (a.logmel) {
                const int total = a.n_mel * F;
                for (int i = tid; i < total; i += kHalfThreads) {
                    const int m = i / F, fr = i % F;
                    scr[fr * a.ls + m] = kLn2 * lg2_fast(fmaxf(rise[i] + fall[i + F], a.log_floor));
                }
            } else {
                const int hm = (a.n_mel + 1) >> 1, total = a.hmp * F;     // hmp = hm rounded up to even: the pad row is zeroed
                float *vp = scr, *vm = scr + a.hmp * F + MFCC_WIDE_VMPAD;
                for (int i = tid; i < total; i += kHalfThreads) {
                    const int q = i / F;
                    float sum = 0.0f, dif = 0.0f;
                    if (q < hm) {
                        const int im = (a.n_mel - 1 - q) * F + (i % F);
                        const float la = kLn2 * lg2_fast(fmaxf(rise[i] + fall[i + F], a.log_floor));
                        const float lb = kLn2 * lg2_fast(fmaxf(rise[im] + fall[im + F], a.log_floor));
                        sum = la + lb;
                        dif = la - lb;
                    }
                    vp[i] = sum;
                    vm[i] = dif;
                }
            }
            // frame energy = all segment sums of the frame (the filterbank's and the two pseudo-segments outside it)
            if (a.energy != MFCC_ENERGY_NONE && tid >= kHalfThreads - F) {
                const int fr = tid - (kHalfThreads - F);
                float e0 = 0.0f, e1 = 0.0f;
                for (int jj = 0; jj < a.n_mel + 3; ++jj) {
                    e0 += rise[jj * F + fr];
                    e1 += fall[jj * F + fr];
                }
                const float le = kLn2 * lg2_fast(fmaxf(e0 + e1, a.log_floor));
                if (a.logmel) scr[fr * a.ls + a.n_mel] = le;
                else scr[a.ef + fr] = le;
            }
        }
        lock_at(5);
        half_sync(half);   // B4: the folded log energies (or the log-mel rows) are in the scratch
        unlock_at(5);

        // ---- S4: log-mel rows are copied out coalesced.  Cepstra: thread (slot, frame) forms c[slot] and, for
        // slot < n2, c[slot + 32] — both of one parity — from the folded pairs: ceil(M/2) terms per cepstrum, one load
        // of v and half a 64-bit load of the table per term ----
        if (a.logmel) {
            const int M = a.od, total = n_frames * M;
            float *o = a.out + tile.out_row * M;
            for (int i = tid; i < total; i += kHalfThreads) {
                const int fr = (i * a.mel_magic) >> 20, m = i - fr * M;
                o[i] = scr[fr * a.ls + m];
            }
        } else if (slot < a.n_cep) {
            const bool second = slot < a.n2;               // this thread also forms c[slot + 32] (uniform per warp: n2 is a multiple of 4)
            float c0 = 0.0f, c1 = 0.0f, e0 = 0.0f, e1 = 0.0f;
            auto dct_terms = [&](int hq) {
                const float *v = scr + ((slot & 1) ? a.hmp * F + MFCC_WIDE_VMPAD : 0) + f;
                const float2 *da = reinterpret_cast<const float2 *>(t_dct) + slot * hq;                    // {d[s][q], d[s][q+1]}
                const float2 *db = reinterpret_cast<const float2 *>(t_dct + kSlots * hq * 2) + slot * hq;  // {d[s+32][q], d[s+32][q+1]}, s < n2
#pragma unroll 4
                for (int q2 = 0; q2 < hq; ++q2) {
                    const float v0 = v[2 * q2 * F], v1 = v[(2 * q2 + 1) * F];
                    const float2 d = da[q2];
                    c0 = fmaf(d.x, v0, c0);
                    e0 = fmaf(d.y, v1, e0);
                    if (second) {
                        const float2 g = db[q2];
                        c1 = fmaf(g.x, v0, c1);
                        e1 = fmaf(g.y, v1, e1);
                    }
                }
            };
            if constexpr (MEL > 0) dct_terms(MEL / 4);     // two folded pairs per step: hmp / 2 steps
            else dct_terms(a.hmp >> 1);
            if (f < n_frames) {
                float *o = a.out + (tile.out_row + f) * a.od + slot;
                o[0] = c0 + e0;
                if (slot + kSlots < a.n_cep) o[kSlots] = c1 + e1;
            }
        }
        if (!a.logmel && a.energy != MFCC_ENERGY_NONE && slot == 0 && f < n_frames) {
            // slot 0 formed c[0] of its frame above (same thread, program order): replace it, or append after the cepstra
            float *o = a.out + (tile.out_row + f) * a.od;
            o[a.energy == MFCC_ENERGY_REPLACE_C0 ? 0 : a.n_cep] = scr[a.ef + f];
        }
        // next S0 writes `staged` (nobody reads P any more); the scratch is next written by S1, after B1
    }
    if (held) atomicExch(phase_lock, 0);   // a span that wraps around the loop is still held after the last unit
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
constexpr int kL = 1200, kHop = 480;
using G0 = Geo<kL, kHop>;

struct WideState {
    WideArgs args{};
    float *d_tab = nullptr;
    size_t smem = 0;
    int sm_count = 0;
};

size_t table_floats(const mfcc_params &p)
{
    const size_t second = p.output == MFCC_OUT_CEPSTRA && p.n_cep > kSlots ? static_cast<size_t>((p.n_cep - kSlots + 3) / 4 * 4) : 0;
    return G0::TABF + 2 * kSlots + 4 * static_cast<size_t>(p.n_mel + 3) + (kSlots + second) * ((p.n_mel + 1) / 2 + 1) + 16;
}

}  // namespace

const char *wide_match(const mfcc_params &p, const HostTables &h)
{
    if (p.frame_len != kL || p.hop_len != kHop || p.nfft != G0::NFFT) return nullptr;
    if (p.output == MFCC_OUT_CEPSTRA && p.n_cep > KC) return nullptr;
    for (int j = 0; j + 1 < static_cast<int>(h.mel_bins.size()); ++j)
        if (h.mel_bins[j + 1] < h.mel_bins[j]) return nullptr;
    if ((table_floats(p) + 2 * static_cast<size_t>(G0::HALF)) * sizeof(float) > kSmemMax) return nullptr;
    if (p.n_mel < 2 || p.log_floor < 1.17549435e-38f) return nullptr;   // band pairs; lg2.approx.ftz needs a normal floor
    // tail scratch (log band energies [n_mel][F] or log-mel rows [F][n_mel | 1], then the per-segment rise / fall
    // sums [n_mel + 1][F] x 2) must stay below the raw PCM buffer
    if (3 * F * static_cast<size_t>(p.n_mel + 4) + F + MFCC_WIDE_VMPAD > static_cast<size_t>(G0::RAWOFF)) return nullptr;
    return "fused_wide_tile8_L1200_H480_real64x32";
}

int wide_prepare(mfcc_plan *plan)
{
    const mfcc_params &p = plan->p;
    const HostTables &h = plan->host;
    if (wide_match(p, h) == nullptr) return MFCC_ENOTSUP;
    constexpr int RA = G0::RA, N = G0::NFFT, H = G0::H, NZ = G0::NZ, NZP = G0::NZP;
    const int M = p.n_mel;
    std::vector<float> tab;
    auto align4 = [&]() { while (tab.size() % 4) tab.push_back(0.0f); };
    auto push_int = [&](int v) { float fl; std::memcpy(&fl, &v, 4); tab.push_back(fl); };
    for (int col = 0; col < RA; ++col)
        for (int b = 0; b < NZP; ++b) {
            const int i = col + RA * b;
            tab.push_back(b < NZ && i < p.frame_len ? h.window[i] : 0.0f);
        }
    for (int col = 0; col < RA; ++col) {
        for (int sl = 0; sl < H; ++sl) {
            const double ang = -2.0 * M_PI * static_cast<double>(col) * (sl + 1) / N;
            tab.push_back(static_cast<float>(std::cos(ang)));
            tab.push_back(static_cast<float>(std::sin(ang)));
        }
        for (int pad = 2 * H; pad < G0::TWS; ++pad) tab.push_back(0.0f);   // row padding (bank separation of the quarter-warps)
    }
    if (static_cast<int>(tab.size()) != G0::TABF) return MFCC_ECUDA;

    WideLayout lay{};
    align4();
    lay.seg = static_cast<int>(tab.size());
    {
        // segments 0 .. M of the filterbank, then (energy term only) the two pseudo-segments below and above it: plain sums
        struct SegD { int k0, w; float sc; };
        std::vector<SegD> segs;
        for (int j = 0; j <= M; ++j) {
            const int k0 = h.mel_bins[j], w = h.mel_bins[j + 1] - k0;
            segs.push_back({k0, w, static_cast<float>(w > 0 ? 1.0 / (static_cast<double>(w) * N) : 0.0)});
        }
        segs.push_back({0, h.mel_bins[0], 0.0f});
        segs.push_back({h.mel_bins[M + 1], h.nbins - h.mel_bins[M + 1], 0.0f});
        // Walk order.  Position j is walked by slot j mod 32 in round j / 32, and the four slots of a warp run in lockstep over
        // FOUR segments at once: P rows are F = 8 words, so two of them meet on the same banks whenever their first bins agree
        // modulo 4 (with neighbouring segments in natural order: 7.4 M excessive wavefronts per launch, every load of the walk
        // twice).  So each group of four positions takes, from the next few segments not yet placed, four whose first bins
        // differ modulo 4 where that is possible — still neighbours, so their widths (loop counts) stay close.
        // (Only the filterbank's own segments move: the two pseudo-segments stay at the end, where a plan without the energy
        // term — nseg = M + 1 — does not walk them.)
        const int total = M + 1;
        std::vector<int> order;
        std::vector<char> used(total, 0);
        int next = 0;
        while (static_cast<int>(order.size()) < total) {
            while (next < total && used[next]) ++next;
            bool taken[4] = {false, false, false, false};
            int got = 0;
            std::vector<int> group;
            for (int c = next; c < total && c < next + (MFCC_WIDE_SEGPERM ? 10 : 4) && got < 4; ++c) {
                if (used[c]) continue;
                const int r = segs[c].k0 & 3;
                if (MFCC_WIDE_SEGPERM && taken[r]) continue;
                taken[r] = true;
                used[c] = 1;
                group.push_back(c);
                ++got;
            }
            for (int c = next; c < total && got < 4; ++c)      // no distinct residue left nearby: fill up in natural order
                if (!used[c]) { used[c] = 1; group.push_back(c); ++got; }
            for (int c : group) order.push_back(c);
        }
        order.push_back(M + 1);
        order.push_back(M + 2);
        for (int pos = 0; pos < M + 3; ++pos) {
            const SegD &g = segs[order[pos]];
            push_int(g.k0 * F);
            push_int(g.w);
            tab.push_back(g.sc);
            push_int(order[pos]);                              // where its two sums go: rise / fall row of the segment
        }
    }
    align4();
    // DCT entries by slot (see S4).  Band m pairs with M - 1 - m; for odd M the middle band pairs with itself,
    // so its entry is halved (exact); the padding column of an odd half is zero.
    const int half_mel = (M + 1) / 2, hmp = (half_mel + 1) / 2 * 2;
    auto dct_at = [&](int k, int m) -> float {
        if (p.output != MFCC_OUT_CEPSTRA || k >= p.n_cep || m >= half_mel) return 0.0f;
        const float v = h.dct[static_cast<size_t>(k) * M + m];
        return (M % 2 == 1 && m == half_mel - 1) ? 0.5f * v : v;
    };
    lay.dct = static_cast<int>(tab.size());
    const int n2 = p.output == MFCC_OUT_CEPSTRA && p.n_cep > kSlots ? (p.n_cep - kSlots + 3) / 4 * 4 : 0;
    for (int sl = 0; sl < kSlots; ++sl)
        for (int m = 0; m < hmp; ++m) tab.push_back(dct_at(sl, m));
    for (int sl = 0; sl < n2; ++sl)
        for (int m = 0; m < hmp; ++m) tab.push_back(dct_at(sl + kSlots, m));
    align4();
    lay.total = static_cast<int>(tab.size());

    WideState *st = new WideState();
    st->sm_count = plan->sm_count;
    st->smem = sizeof(float) * (static_cast<size_t>(lay.total) + 2 * static_cast<size_t>(G0::HALF));
    if (st->smem > kSmemMax) { delete st; return MFCC_ENOTSUP; }
    st->args.lay = lay;
    st->args.n_mel = M;
    st->args.n_cep = p.n_cep;
    st->args.logmel = p.output == MFCC_OUT_LOGMEL;
    st->args.energy = p.energy;
    st->args.od = h.out_dim;
    st->args.nseg = M + 1 + (p.energy != MFCC_ENERGY_NONE ? 2 : 0);
    st->args.inv_n = static_cast<float>(1.0 / N);
    st->args.ls = h.out_dim | 1;
    st->args.mel_magic = (1 << 20) / h.out_dim + 1;
    st->args.hmp = hmp;
    st->args.n2 = n2;
    st->args.rf = (std::max(F * std::max(h.out_dim | 1, M), 2 * F * hmp + MFCC_WIDE_VMPAD) + 3) / 4 * 4;
    st->args.ef = st->args.rf + 2 * F * (M + 3);
    st->args.preemph = p.preemph;
    st->args.log_floor = p.log_floor;
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
    plan->wide_state = st;
    return MFCC_OK;
}

void wide_release(mfcc_plan *plan)
{
    WideState *st = static_cast<WideState *>(plan->wide_state);
    if (st == nullptr) return;
    if (st->d_tab) cudaFree(st->d_tab);
    delete st;
    plan->wide_state = nullptr;
}

template <typename PcmT>
int wide_launch(const mfcc_plan *plan, const Tile *d_tiles, int64_t n_tiles, const PcmT *d_pcm, int64_t /*pcm_len*/,
                float *d_out, cudaStream_t stream)
{
    const WideState *st = static_cast<const WideState *>(plan->wide_state);
    if (st == nullptr) return MFCC_ENOTSUP;
    // the BASELINE.json band count (80) with cepstral output gets the unrolled DCT
    const bool fixed = !st->args.logmel && st->args.n_mel == 80;
    auto kern = fixed ? fused_wide_kernel<PcmT, kL, kHop, 80> : fused_wide_kernel<PcmT, kL, kHop, 0>;
    static std::atomic<uint64_t> optin[2];
    if (ensure_smem_optin(kern, plan->device, kSmemMax, optin[fixed]) != MFCC_OK) return MFCC_ECUDA;
    WideArgs a = st->args;
    a.tiles = d_tiles;
    a.n_tiles = n_tiles;
    a.out = d_out;
    a.tab = st->d_tab;
    const int64_t grid = std::min<int64_t>((n_tiles * (kTileFrames / F) + 1) / 2, st->sm_count);
    kern<<<static_cast<unsigned>(grid), kThreads, st->smem, stream>>>(d_pcm, a);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError() == cudaSuccess ? MFCC_OK : MFCC_ECUDA;
}

template int wide_launch<int16_t>(const mfcc_plan *, const Tile *, int64_t, const int16_t *, int64_t, float *, cudaStream_t);
template int wide_launch<float>(const mfcc_plan *, const Tile *, int64_t, const float *, int64_t, float *, cudaStream_t);

}  // namespace mfcc
