// mfcc_fused_sp.cu — the streamlined fused tile kernel ("sp"): compile-time frame geometry,
// table-driven filterbank, bulk-copy staging, one 512-thread CTA per SM working as two
// independent 8-warp halves.
//
// What changed against mfcc_fused_ct.cu and why (profiles/r1_v4_real32x16_A.md: the two FFT
// passes were 56 % of the instructions but 37 % of the time; staging, mel, log, DCT and store
// took the rest in seven barrier-separated, latency-bound phases):
//   * S0 input: the next tile's PCM arrives by ONE bulk async copy (cp.async.bulk, TMA engine)
//     into a raw int16 buffer while the current tile is transformed — no prefetch registers,
//     no LDG/LDL instructions, completion on an mbarrier.  Staging then converts 8 samples per
//     thread (LDS.128) instead of 2.
//   * filterbank: a segment [b_j, b_j+1) rises into filter j and falls out of filter j - 1 with linear weights,
//     so two running sums per segment give both halves; every segment is walked once, by the warp the host
//     gave it to (cost-balanced lists).  In the tail-warp variants (the BASELINE filter / cepstrum counts) the
//     segment descriptors are kernel PARAMETERS read under a warp-uniform index: uniform registers, uniform
//     branches, one load and two additions per bin.
//   * tail: band assembly + log + DCT of tile t run on TWO "tail warps" (lane = frame, everything in registers,
//     DCT entries fetched from the parameter bank four at a time, even cepstra on one warp, odd on the other)
//     WHILE the other six warps stage tile t + 1.  4 barriers per tile; the all-warp S3b + S4 of the generic
//     variant (1,400 instructions per tile, 8 x redundant loads of the log energies) is gone from the hot path:
//     1.545 -> 1.75 G frames/s on configs[1].
//   * hot code stays under 32 KB: profiles/r1_v5_sp_unrolled_tail_A.md shows what happens when the
//     tail is unrolled per warp (60 KB of code, 44 % of stall samples = no_instruction), so the
//     tail is a table-driven loop shared by all warps.
//   * one CTA per SM, two halves: the tables are held once per SM, which is what makes the raw
//     buffer fit; each half has its own staging/workspace/raw buffers, mbarrier and NAMED barrier,
//     so the halves drift apart like two CTAs would and cover each other's barrier waits.
// Phases S1 (windowed real DFT-RB per column pair + inter-pass twiddle) and S2 (complex DFT-RA per
// row + power) are those of mfcc_fused_ct.cu.
//
// No reference code corresponds to this (SURVEY.md §8a "Ref file:line = none").
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <type_traits>
#include <vector>

#include "mfcc_rfft.cuh"
#include "mfcc_host.h"

// Build-time switches for A/B measurements on one box (tools/variants_sp.py); the defaults are the measured winners.
#ifndef MFCC_SP_TW2_16
#define MFCC_SP_TW2_16 1
#endif
#ifndef MFCC_SP_TW2_32
#define MFCC_SP_TW2_32 0
#endif
// Phase ablation (timing only, results are wrong by construction; profiles/r1_ablation_sp_A.md): bit 0 S0 staging
// arithmetic, 1 S1 pass 1, 2 S2 pass 2, 3 S3 filterbank sums, 4 tail (band assembly + log + DCT + store).  Barriers,
// bulk copies and tile bookkeeping stay.
// Groups per CTA (timing experiments: how much one group gains from its neighbours)
#ifndef MFCC_SP_GROUPS_512
#define MFCC_SP_GROUPS_512 2
#endif
#ifndef MFCC_SP_GROUPS_256
#define MFCC_SP_GROUPS_256 3
#endif
#ifndef MFCC_SP_GROUP_DELAY_NS
#define MFCC_SP_GROUP_DELAY_NS 2700
#endif
#ifndef MFCC_SP_ABLATE
#define MFCC_SP_ABLATE 0
#endif
// Mutual exclusion of the groups of a CTA on a span of phases: 10 * from + to, the lock is taken before barrier B<from> and
// released after barrier B<to> (B1 | S1 | B2 | S2 | B3 | S3 | B4): 12 = pass 1 exclusive, 23 = pass 2 exclusive, 0 = none.
// Inside a phase the 8 warps of a group do the same thing at the same time (a burst of shared-memory loads, FP32, a burst
// of stores); with the FFT pass of one group always running over the latency-bound phases of the others (S3, S0, tail) the
// issue slots and the shared-memory port are shared better than when the groups drift freely.  Measured on one box
// (tools/time_variants.py, profiles/r2_phase_lock.md), against no lock: 512-point S3 exclusive +2.5 % (S2 +1.7 %, S1 +1.4 %,
// any two of them together less than either alone, S1 + S2 as one section -17 %); 256-point (three groups) S1 exclusive
// +2.5 % (S2 -2.5 %, S3 -7 %, S0 -11 %).  The winners keep two latency-bound phases from coinciding (512-point: both
// groups walking the filterbank leaves the SM idle) or the FP32-heaviest phase from tripling up (256-point).
#ifndef MFCC_SP_LOCK_512
#define MFCC_SP_LOCK_512 34
#endif
#ifndef MFCC_SP_LOCK_256
#define MFCC_SP_LOCK_256 12
#endif
#ifndef MFCC_SP_LOCK_NS
#define MFCC_SP_LOCK_NS 32
#endif
// how many groups may hold the lock at once (a counting semaphore when > 1; only meaningful for the three-group CTA)
#ifndef MFCC_SP_LOCK_HOLDERS
#define MFCC_SP_LOCK_HOLDERS 1
#endif
// S0 of bulk-copied int16 tiles, second form: the predecessor sample loaded with the chunk, the rare split-pad chunk on a
// branch instead of four selects per chunk; MFCC_SP_I2F: int16 -> f32 on the conversion pipe (one I2F per sample).
#ifndef MFCC_SP_S0V2
#define MFCC_SP_S0V2 0
#endif
#ifndef MFCC_SP_I2F
#define MFCC_SP_I2F 0
#endif
// S0 of bulk-copied int16 tiles: lanes rotate the order of their four 8-byte stores (see the comment at the stores)
#ifndef MFCC_SP_S0ROT
#define MFCC_SP_S0ROT 0
#endif
// Pass-1 constants (window pairs, and for 32 x 16 the inter-pass twiddles) from the kernel PARAMETER bank under a warp-uniform
// index (uniform constant loads into uniform registers: LDCU.64 pairs in the shipped binary, profiles/r2_sass_sp_A.md) instead of shared memory: takes 232 broadcast LDS.128 per 32-frame tile off the
// shared-memory port (ncu: 117.1 M -> 100.9 M wavefronts per configs[1] launch) and the constants out of the vector registers.  Measured on one box, bit-identical
// results (tools/time_variants.py): 512-point +5.1 % (1.818 -> 1.910 G frames/s on configs[1]); 256-point, window only (its
// twiddles are applied in pass 2) -4.6 % — the switch is per geometry: bit 0 = 512-point pass 1, bit 1 = 256-point window,
// bit 2 = 256-point pass-2 twiddles.
#ifndef MFCC_SP_CONST_TABLES
#define MFCC_SP_CONST_TABLES 1
#endif
// The 16 twiddles of the special pass-2 item (rows 0 and H) as constant-bank operands instead of 8 LDS.128: bit 0 = 512-point,
// bit 1 = 256-point.  Measured (one box, bit-identical): 512-point -0.5 %, 256-point +0.7 % -> on for 256-point only.
#ifndef MFCC_SP_CONST_TWH
#define MFCC_SP_CONST_TWH 2
#endif
// Which PCM entries this translation unit instantiates (the file is compiled once per input type so that the three sets
// of kernel variants build in parallel): bit 0 int16 (+ the host half), bit 1 f32, bit 2 G.711 codes.
#ifndef MFCC_SP_PCM_TYPES
#define MFCC_SP_PCM_TYPES 7
#endif
// Poison build (libmfcc_b200_poison.so, the compute-sanitizer substitute: the pool's boxes refuse the sanitizer).  The
// kernel's buffers alias each other (staged / P, workspace / tail scratch, raw PCM refilled by the TMA engine); the
// barrier reasoning in the comments says when each one is dead.  With MFCC_POISON every buffer is filled with NaN at
// the point where it is said to be dead (one extra barrier each time), so any value that is read after its buffer
// died, or before it was written for this tile, reaches the output as NaN.  tests/test_gpu_parity.py runs the
// parity batches through this build.
#ifndef MFCC_POISON
#define MFCC_POISON 0
#endif

namespace mfcc {

namespace {

constexpr int kWarps = 8;                 // per half
constexpr int kHalfThreads = kWarps * 32;
// Groups ("halves") of 8 warps per CTA: 2 for the 512-point geometry (128 registers per thread, 108 KB of shared
// memory per group), 3 for the 256-point one (its codelets fit 80 registers and a group needs 55 KB), which buys
// the latency-bound small-FFT path 24 resident warps instead of 16 (measured: 2.32 -> 2.61 G frames/s; 4 groups at
// 64 registers: 2.52 G).
constexpr int groups_for(int rb) { return rb >= 32 ? MFCC_SP_GROUPS_512 : MFCC_SP_GROUPS_256; }
constexpr int kPad = 2;
constexpr int kSegMax = 20;                // segments one warp may be given (n_mel + 1 <= 8 * kSegMax)
constexpr int KC = 16;                    // cepstra per DCT round: warp w forms k = w and k = w + 8 (tail-warp variants: n_cep <= KC)
constexpr int kSegParam = 12;              // segments per warp (+ terminator) a tail-warp variant can take as parameters
constexpr int kFoldMax = 20;                // ceil(n_mel / 2) of a tail-warp variant with cepstral output (n_mel <= 40)
constexpr size_t kSmemMax = 227 * 1024;

template <int L_, int HOP_, int RB_, int RA_>
struct Geo {
    static constexpr int L = L_, HOP = HOP_, RB = RB_, RA = RA_;
    static constexpr int NFFT = RB * RA, NB = NFFT / 2 + 1, H = RB / 2;
    static constexpr int NZ = (L + RA - 1) / RA;           // rows b of a column that carry samples
    static constexpr int NZP = (NZ + 1) / 2 * 2;
    static constexpr int STRIDE = HOP + kPad;              // staged words per hop block
    static constexpr int padded(int i) { return i + kPad * (i / HOP); }
    static constexpr int SLACK = RA * NZ - L;              // words past a frame's end read with a zero window
    static constexpr int tceil(int n_frames) { return ((n_frames - 1) * HOP + L + SLACK + 7) / 8 * 8; }
    static constexpr int tceil_s(int n_frames, int shift) { return (shift + (n_frames - 1) * HOP + L + SLACK + 7) / 8 * 8; }
    static constexpr int TCEIL = tceil(32) + 8;            // samples staged for a full tile (+ up to 7 of alignment shift)
    static constexpr int STAGED = padded(TCEIL) + 8;
    static constexpr int PW = (NB + 3) * 32;               // 3 zeroed slack rows: 4-bin chunks read past the last bin
    static constexpr int UNION = ((STAGED > PW ? STAGED : PW) + 3) / 4 * 4;
    static constexpr int WS = H * RA * 32 * 2;             // floats
    static constexpr int RAW = (TCEIL + 16) / 2;           // floats holding 8 lead + TCEIL + 1 look-ahead int16 samples
    static constexpr int DESC = 24;                        // two 40-byte tile descriptors (current, next)
    static constexpr int HALF = UNION + WS + RAW + 4 + DESC;   // + mbarrier (8 B in a 16-B slot)
    // fixed part of the table blob (floats); the filterbank tables follow at run-time offsets
    static constexpr int T_WIN = 0;                        // [RA/2][NZP] float2
    // Where the inter-pass twiddle is applied.  16 x 16: pass 2 has one item per warp, seven plain rows and the longer
    // special row, so the plain rows take the twiddle on load (measured +1.4 %); 32 x 16: pass 1 keeps it (-0.7 % moved).
    static constexpr bool TW_IN_PASS2 = RB == 16 ? MFCC_SP_TW2_16 : MFCC_SP_TW2_32;
    static constexpr bool CONST_TAB = ((MFCC_SP_CONST_TABLES) >> (RB == 16 ? 1 : 0)) & 1;   // pass-1 constants from the parameter bank
    static constexpr bool CONST_TW2 = RB == 16 && (((MFCC_SP_CONST_TABLES) >> 2) & 1);       // 16 x 16: the pass-2 twiddles from it
    static constexpr int T_TW = T_WIN + RA / 2 * NZP * 2;  // TW_IN_PASS2 ? [H][RA] (row k1 - 1, column a) : [RA][H] float2
    static constexpr int T_TWH = T_TW + RA * H * 2;        // [RA] float2
    static constexpr int TABF = T_TWH + RA * 2;
    static_assert(HOP % 8 == 0, "an 8-sample chunk must not straddle a hop block");
    static_assert(HOP % RA == 0, "a column pair must not straddle a hop block");
    static_assert(RA == 2 * kWarps && RA == 16, "one column pair per warp, 16-point second pass");
    static_assert(L <= NFFT && NZ <= RB, "frame does not fit the transform");
    static_assert(TABF % 4 == 0 && UNION % 4 == 0 && WS % 4 == 0 && RAW % 4 == 0, "16-byte aligned regions");
    static_assert(32 * 129 <= WS, "tail scratch must fit in the workspace");
};

// Run-time part of the table blob (offsets in floats from its start).
struct SpLayout {
    int wseg;     // per warp: kSegMax segment descriptors in walk order, float4 {first bin * 32 (int), width w (int),
                  // s = 1 / (w NFFT), segment index j (int)}; unused slots have j = -1
    int dct;      // [rounds][kWarps][n_mel] float2: DCT entries {d[16 r + w][m], d[16 r + w + 8][m]}, zero past n_cep (n_mel padded to even)
    int total;    // floats, multiple of 4
};

struct SpArgs {
    const Tile *tiles;
    int64_t n_tiles;
    float *out;
    const float *tab;     // global copy of the table blob
    SpLayout lay;
    int n_mel, n_cep, logmel;
    int energy;           // MFCC_ENERGY_*: frame energy = sum of the one-sided power spectrum, taken from the segment sums
    int od;               // floats per output row (n_cep or n_mel, + 1 under MFCC_ENERGY_APPEND)
    int alaw;             // G.711 entry (uint8 codes): 0 mu-law, 1 A-law
    int ls;               // log-mel staging row stride (odd, >= od)
    int rf;               // scratch offset (floats) of the per-segment rise / fall sums [n_mel + 3][32] x 2 (segments 0 .. n_mel, then the two
                          // pseudo-segments below / above the filterbank that only the energy term reads)
    int ef;               // scratch offset of the log frame energy [32] (generic tail)
    int raw32_off;        // f32 PCM: float offset in the workspace of the raw buffer the NEXT tile's samples are bulk-copied into
                          // (the end of the workspace, past the tail scratch); 0: no room, f32 tiles are staged straight from HBM
    float inv_n;          // 1 / NFFT
    int mp;               // n_mel rounded up to even (DCT row length in the table)
    int mel_magic;        // i / od == (i * mel_magic) >> 20 for i < 32 * od
    float preemph, log_floor;
    // tail-warp variants (MEL > 0): ln 2 * d[k][q] for q < MEL / 2 (the mirrored half follows from d[k][M - 1 - q] =
    // (-1)^k d[k][q]).  Kernel parameters live in the constant bank, so with compile-time indices every entry is
    // a constant-bank FFMA operand (LDC.64 in the shipped binary: the index depends on the warp parity): no shared-memory loads in the DCT.
    float4 dctc[KC / 2][2][kFoldMax / 4];   // [k / 2][k & 1][q / 4]; odd n_mel: the middle band pairs with itself, its entry is halved
    // tail-warp variants: the segment walks of S3 as parameters too, {first bin * 128 (byte offset into P), width w,
    // s = 1 / (w NFFT), segment index * 128 (byte offset into the rise / fall scratch; -1 ends the list)}: read with a
    // warp-uniform index they arrive in uniform registers, so every branch of the walk is a uniform branch.
    float4 useg[kWarps][kSegParam];
    // MFCC_SP_CONST_TABLES: the pass-1 constants as parameters too (filled for the geometries that use them)
    float4 winc[kWarps][14];    // [column pair][b / 2]: window values of rows b, b + 1 of the pair's two columns
    float4 twc[16][8];          // 32 x 16 only: [column][(k1 - 1) / 2] inter-pass twiddles of k1, k1 + 1
    float4 twhc[8];             // row-H twiddles W_(2 RA)^a of columns 2 q, 2 q + 1: compile-time indices, so they are plain
                                // constant-bank operands of the multiplications (MFCC_SP_CONST_TWH)
};

__device__ __forceinline__ float2 lds_f2(const float *p) { return *reinterpret_cast<const float2 *>(p); }
__device__ __forceinline__ float4 lds_f4(const float *p) { return *reinterpret_cast<const float4 *>(p); }

// int16 pair -> two exact floats without the XU pipe: (v ^ 0x8000) = v + 32768 sits in the mantissa of
// 2^23 (bits 0x4B00xxxx); one XOR serves both halves, one byte permute (PRMT) builds each float.
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
__device__ __forceinline__ float to_f32(int16_t v, int) { return static_cast<float>(v); }
__device__ __forceinline__ float to_f32(float v, int) { return v; }
// G.711 code -> linear PCM value as an exact float, integer pipe only (no I2F): the magnitude (< 2^15) is dropped into
// the mantissa of 2^23, the sign is XORed in.  Same arithmetic as g711_kernel (mfcc_generic.cu) / ITU-T G.711 tables.
__device__ __forceinline__ float g711_to_f32(uint32_t b, int alaw)
{
    uint32_t mag, neg;
    if (alaw) {
        const uint32_t a = b ^ 0x55u, seg = (a >> 4) & 7u, man = a & 0x0Fu;
        mag = seg == 0 ? (man << 4) + 8u : ((man << 4) + 0x108u) << (seg - 1);
        neg = (~a) & 0x80u;
    } else {
        const uint32_t u = (~b) & 0xFFu;
        mag = ((((u & 0x0Fu) << 3) + 0x84u) << ((u >> 4) & 7u)) - 0x84u;
        neg = u & 0x80u;
    }
    const float f = __uint_as_float(0x4B000000u | mag) - 8388608.0f;
    return __uint_as_float(__float_as_uint(f) ^ (neg << 24));
}
__device__ __forceinline__ float to_f32(uint8_t v, int alaw) { return g711_to_f32(v, alaw); }

// ---- mbarrier + bulk async copy (TMA engine, 1-D) + named barrier ----
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
__device__ __forceinline__ void cp_async8(uint32_t dst, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void half_sync(int half)
{
    asm volatile("bar.sync %0, %1;" ::"r"(half + 1), "n"(kHalfThreads) : "memory");
}

// MEL > 0: the plan has exactly MEL filters and CEP cepstra per frame.  Band assembly, log and DCT of tile t are then
// done by ONE warp (the "tail warp", lane = frame, everything in registers, DCT entries as constant-bank operands)
// while the other seven stage tile t + 1: 364 instructions per tile instead of the 1,400 of the all-warp S3b + S4,
// and one barrier fewer.  MEL = 0: any plan (all-warp S3b + S4, table-driven DCT loop).
template <typename PcmT, int L_, int HOP_, int RB_, int RA_, int MEL, int CEP>
__global__ void __launch_bounds__(groups_for(RB_) * kHalfThreads, 1) fused_sp_kernel(const PcmT *__restrict__ pcm, const SpArgs a)
{
    using G = Geo<L_, HOP_, RB_, RA_>;
    constexpr int GROUPS = groups_for(RB_), kThreads = GROUPS * kHalfThreads;
    constexpr int HOP = G::HOP, RB = G::RB, RA = G::RA, STRIDE = G::STRIDE, H = G::H, NZ = G::NZ;
    extern __shared__ __align__(16) float smem[];
    // (two words next to group 0's mbarrier, which uses 8 bytes of its 16-byte slot)
    [[maybe_unused]] int *phase_lock = reinterpret_cast<int *>(smem + a.lay.total + Geo<L_, HOP_, RB_, RA_>::UNION + Geo<L_, HOP_, RB_, RA_>::WS + Geo<L_, HOP_, RB_, RA_>::RAW + 2);
    // up to two independent locks, two digits each (from, to): the group is held at barrier B<from> until it owns the lock and
    // gives it back after barrier B<to> (to < from: the span wraps around the tile loop, e.g. 41 = S0 exclusive)
    constexpr int kLock = groups_for(RB_) > 1 ? (RB_ >= 32 ? MFCC_SP_LOCK_512 : MFCC_SP_LOCK_256) : 0;
    constexpr int kLockFrom[2] = {(kLock % 100) / 10, kLock / 1000}, kLockTo[2] = {kLock % 10, (kLock / 100) % 10};
    constexpr bool kLockOn[2] = {kLock % 100 != 0, kLock / 100 != 0};
    if (kLock != 0 && threadIdx.x < 2) phase_lock[threadIdx.x] = 0;
    bool held[2] = {false, false};
    [[maybe_unused]] auto lock_at = [&](int point) {      // before barrier B<point> (0: before S0)
#pragma unroll
        for (int i = 0; i < 2; ++i)
            if (kLockOn[i] && point == kLockFrom[i] && threadIdx.x % kHalfThreads == 0) {
                if constexpr (MFCC_SP_LOCK_HOLDERS <= 1) {
                    while (atomicCAS(&phase_lock[i], 0, 1) != 0) __nanosleep(MFCC_SP_LOCK_NS);
                } else {
                    while (atomicAdd(&phase_lock[i], 1) >= MFCC_SP_LOCK_HOLDERS) {
                        atomicSub(&phase_lock[i], 1);
                        __nanosleep(MFCC_SP_LOCK_NS);
                    }
                }
                held[i] = true;
            }
    };
    [[maybe_unused]] auto unlock_at = [&](int point) {    // after barrier B<point>
#pragma unroll
        for (int i = 0; i < 2; ++i)
            if (kLockOn[i] && point == kLockTo[i] && held[i]) {
                if constexpr (MFCC_SP_LOCK_HOLDERS <= 1) atomicExch(&phase_lock[i], 0);
                else atomicSub(&phase_lock[i], 1);
                held[i] = false;
            }
    };
    const int half = threadIdx.x >> 8, tid = threadIdx.x & (kHalfThreads - 1);
    const int lane = tid & 31, warp = tid >> 5;

    float *tab = smem;                                        // shared by both halves
    float *mine = smem + a.lay.total + half * G::HALF;
    float *staged = mine;                 // S0-S1
    float *pw = staged;                   // S2-S3 (aliases staged)
    float2 *ws = reinterpret_cast<float2 *>(mine + G::UNION);
    float *scr = reinterpret_cast<float *>(ws);               // tail scratch (aliases ws)
    const int16_t *raw16 = reinterpret_cast<const int16_t *>(mine + G::UNION + G::WS);
    const uint32_t raw_s = smem_u32(raw16);
    const uint32_t bar = smem_u32(mine + G::UNION + G::WS + G::RAW);
    const Tile *desc = reinterpret_cast<const Tile *>(mine + G::UNION + G::WS + G::RAW + 4);   // [2]
    static_assert(sizeof(Tile) == 40, "descriptor copy assumes five 8-byte words");
    // the tile table is read one tile ahead straight into shared memory (no registers held across the phases)
    auto fetch_desc = [&](int slot, int64_t t) {
        if (tid < 5) cp_async8(smem_u32(desc + slot) + tid * 8, reinterpret_cast<const char *>(a.tiles + t) + tid * 8);
    };

    [[maybe_unused]] auto poison = [&](float *p, int n) {
        for (int i = tid; i < n; i += kHalfThreads) p[i] = __int_as_float(0x7fc00000);
    };
    if (tid == 0) mbar_init(bar, 1);
    // slack rows of P: read with zero weights, so they must hold finite values (0 * NaN would turn a
    // band into NaN and fmaxf would then silently replace it by the floor)
    for (int i = G::NB * 32 + tid; i < G::PW; i += kHalfThreads) pw[i] = 0.0f;
    for (int i = threadIdx.x * 4; i < a.lay.total; i += kThreads * 4)
        *reinterpret_cast<float4 *>(tab + i) = __ldg(reinterpret_cast<const float4 *>(a.tab + i));
    const float *t_win = tab + G::T_WIN, *t_tw = tab + G::T_TW, *t_twh = tab + G::T_TWH;
    const float4 *t_wseg = reinterpret_cast<const float4 *>(tab + a.lay.wseg);
    const float *t_dct = tab + a.lay.dct;

    // A tile takes the bulk-copy path when the PCM array is 16-byte aligned, every frame lies inside the
    // utterance (no zero fill) and the staged span stays inside the array (both checked on the host when the
    // tile table is built: Tile::flags).  The utterance may start at ANY sample: the copy starts at the
    // 8-sample boundary `o` below the tile's first sample and the shift s = first_sample - o (0..7) is
    // absorbed by the staging (see S0).
    const bool base_aligned = (reinterpret_cast<uintptr_t>(pcm) & 15) == 0;
    const bool base_ok = sizeof(PcmT) <= 2 && base_aligned;   // int16 PCM and G.711 codes arrive by bulk copy
    static_assert(G::SLACK + 16 <= kTileSpanSlack, "the host's span check must cover the staged span (16-byte units of 1-byte codes included)");
    auto tile_fast = [&](const Tile &tl) -> bool { return base_ok && (tl.flags & kTileInside) != 0; };
    // f32 PCM: same eligibility, but the samples are read straight from HBM with 16-byte loads in S0 (a raw f32
    // buffer would need 21 KB per group, which the 512-point geometry does not have)
    auto tile_vec = [&](const Tile &tl) -> bool { return sizeof(PcmT) == 4 && base_aligned && (tl.flags & kTileInside) != 0; };
    // ... unless the workspace has room for a raw f32 buffer past the tail scratch: then the next tile's samples are
    // bulk-copied there as soon as pass 2 has released the workspace (after B3) and S0 reads them from shared memory
    // (round 1 staged f32 PCM with exposed HBM latency: 1.22 G against 1.75 G frames/s for int16)
    const float *raw32 = scr + a.raw32_off;
    auto tile_pre32 = [&](const Tile &tl) -> bool { return a.raw32_off > 0 && tile_vec(tl); };
    auto issue_copy32 = [&](const Tile &tl) {
        const int s = static_cast<int>(tl.first_sample & 7);
        const int64_t o = tl.first_sample - s;
        const int lead = o >= 8 ? 8 : 0;
        const uint32_t bytes = static_cast<uint32_t>(lead + G::tceil_s(tl.n_frames, s)) * 4u;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the workspace was last touched through the generic proxy
        mbar_expect_tx(bar, bytes);
        bulk_g2s(smem_u32(raw32) + (8 - lead) * 4, pcm + o - lead, bytes, bar);
    };
    // raw16[8 + i] = x[o + i]; the 8 samples before o ride along when they exist
    auto issue_copy = [&](const Tile &tl) {
        const int s = static_cast<int>(tl.first_sample & 7);
        if constexpr (sizeof(PcmT) == 1) {
            // G.711 codes: 16-byte units are 16 samples, so the copy starts at the 16-sample boundary o16 below the tile
            // (raw bytes [16 + i] = code of sample o16 + i, the 16 codes before o16 ride along when they exist)
            const int64_t o16 = tl.first_sample & ~static_cast<int64_t>(15);
            const int lead = o16 >= 16 ? 16 : 0;
            const int span = (static_cast<int>(tl.first_sample) & 8) + G::tceil_s(tl.n_frames, s);
            const uint32_t bytes = static_cast<uint32_t>(lead + ((span + 15) & ~15));
            mbar_expect_tx(bar, bytes);
            bulk_g2s(raw_s + (16 - lead), pcm + o16 - lead, bytes, bar);
        } else {
            const int64_t o = tl.first_sample - s;
            const int lead = o >= 8 ? 8 : 0;
            const uint32_t bytes = static_cast<uint32_t>(lead + G::tceil_s(tl.n_frames, s)) * 2u;
            mbar_expect_tx(bar, bytes);
            bulk_g2s(raw_s + (8 - lead) * 2, pcm + o - lead, bytes, bar);
        }
    };

    // staging threads of a group: all of it, or all but the tail warp
    constexpr bool kTail = MEL > 0;
    constexpr int kTailWarps = 2;   // cepstra of even k on one, of odd k on the other
    constexpr int kStage = kTail ? kHalfThreads - 32 * kTailWarps : kHalfThreads;
    const bool stager = !kTail || warp < kWarps - kTailWarps;
    // tail warp: band m = rise of segment m + fall of segment m + 1, log2, mirrored-pair fold, DCT with constant-bank
    // entries (ln 2 folded in on the host), CEP stores per frame
    // CEP > 0: cepstra (even k on one tail warp, odd k on the other); CEP == 0: log-mel rows (even bands on one, odd on the
    // other).  The frame energy (MFCC_ENERGY_*) is the sum of ALL segment sums: every band, plus the fall of segment 0
    // and the rise of segment M (the two half-triangles that belong to no band), plus the two pseudo-segments that
    // cover the bins below and above the filterbank.
    auto tail = [&](int64_t out_row, int nf) {
        if constexpr (kTail && (MFCC_SP_ABLATE & 16) == 0) {
            constexpr int HM = (MEL + 1) / 2;
            static_assert((CEP == 0 || HM <= kFoldMax) && CEP <= KC, "tail-warp variant limits");
            const float *rise = scr + a.rf + lane, *fall = rise + (MEL + 3) * 32;
            const int par = warp & 1;
            float l[MEL];
#pragma unroll
            for (int m = 0; m < MEL; ++m) l[m] = rise[m * 32] + fall[(m + 1) * 32];
            float le = 0.0f;
            if (a.energy != MFCC_ENERGY_NONE) {   // uniform branch
                float e0 = fall[0] + rise[MEL * 32], e1 = rise[(MEL + 1) * 32] + fall[(MEL + 1) * 32];
                e1 += rise[(MEL + 2) * 32] + fall[(MEL + 2) * 32];
#pragma unroll
                for (int m = 0; m + 1 < MEL; m += 2) { e0 += l[m]; e1 += l[m + 1]; }
                if (MEL & 1) e0 += l[MEL - 1];
                le = kLn2 * lg2_fast(fmaxf(e0 + e1, a.log_floor));
            }
#pragma unroll
            for (int m = 0; m < MEL; ++m) l[m] = lg2_fast(fmaxf(l[m], a.log_floor));
            if constexpr (CEP == 0) {
                if (lane < nf) {
                    float *o = a.out + (out_row + lane) * a.od;
#pragma unroll
                    for (int m = 0; m < MEL; ++m)
                        if ((m & 1) == par) o[m] = kLn2 * l[m];
                    if (par == 1 && a.energy == MFCC_ENERGY_APPEND) o[MEL] = le;
                }
            } else {
                // this warp's cepstra k = 2 kk + par all take v[q] = l[q] + (-1)^par l[M - 1 - q]; DCT entries from the constant
                // bank (ln 2 folded in on the host; odd M: the middle band meets itself, its entry is halved)
                const float sgn = par ? -1.0f : 1.0f;
                float v[HM];
#pragma unroll
                for (int q = 0; q < HM; ++q) v[q] = fmaf(sgn, l[MEL - 1 - q], l[q]);
                constexpr int CH = (CEP + 1) / 2;
                float c[CH];
#pragma unroll
                for (int kk = 0; kk < CH; ++kk) c[kk] = 0.0f;
#pragma unroll
                for (int q4 = 0; q4 < (HM + 3) / 4; ++q4)
#pragma unroll
                    for (int kk = 0; kk < CH; ++kk) {
                        const float4 d = a.dctc[kk][par][q4];   // one 16-byte uniform load from the parameter bank (zeros past n_cep)
                        c[kk] = fmaf(d.x, v[4 * q4], c[kk]);
                        if (4 * q4 + 1 < HM) c[kk] = fmaf(d.y, v[4 * q4 + 1], c[kk]);
                        if (4 * q4 + 2 < HM) c[kk] = fmaf(d.z, v[4 * q4 + 2], c[kk]);
                        if (4 * q4 + 3 < HM) c[kk] = fmaf(d.w, v[4 * q4 + 3], c[kk]);
                    }
                if (par == 0 && a.energy == MFCC_ENERGY_REPLACE_C0) c[0] = le;
                if (lane < nf) {
                    float *o = a.out + (out_row + lane) * a.od + par;
#pragma unroll
                    for (int kk = 0; kk < CH; ++kk)
                        if (2 * kk + 1 < CEP || par == 0) o[2 * kk] = c[kk];
                    if (par == 1 && a.energy == MFCC_ENERGY_APPEND) o[CEP - 1] = le;
                }
            }
        }
    };
    int64_t prev_row = 0;
    int prev_nf = 0;      // frames of the tile whose tail is still to be done (0: none)

    const int64_t first = GROUPS * static_cast<int64_t>(blockIdx.x) + half, step = GROUPS * static_cast<int64_t>(gridDim.x);
    if (first < a.n_tiles) fetch_desc(0, first);
    cp_async_wait_all();
    __syncthreads();   // mbarriers initialised, tables and first descriptors visible; from here on the halves only meet themselves
    if (first < a.n_tiles && tid == 0) {
        const Tile t0 = desc[0];
        if (tile_fast(t0)) issue_copy(t0);
        if (tile_pre32(t0)) issue_copy32(t0);
    }
    // Start the two groups of the 512-point CTA half a tile (2.7 us) apart.  Inside a phase the 8 warps of a group do
    // the same thing at the same time (a burst of loads, FP32, a burst of stores), so how well the groups fill each other's
    // gaps depends on their relative phase; released together by the __syncthreads above they start IN phase.  Measured
    // +0.9 % on configs[1] (1.764 G against 1.746-1.751 G frames/s); not applied to the three-group 256-point CTA
    // (not measured there).  Timing only: results do not depend on it.
    if constexpr (RB == 32 && MFCC_SP_GROUP_DELAY_NS > 0) {
        if (half > 0) __nanosleep(static_cast<unsigned>(half) * MFCC_SP_GROUP_DELAY_NS);
    }
    uint32_t phase = 0;
    int cur = 0;
    for (int64_t t = first;; t += step, cur ^= 1) {
        // the previous tile's tail, concurrent with this tile's S0 (one call site: the tail is 300 instructions)
        if (!stager && prev_nf > 0) tail(prev_row, prev_nf);
        if (t >= a.n_tiles) break;
        const Tile tile = desc[cur];
        const bool fast = tile_fast(tile), pre32 = tile_pre32(tile), vec = tile_vec(tile) && !pre32;
        const bool has_next = t + step < a.n_tiles;
        if (has_next) fetch_desc(cur ^ 1, t + step);   // lands while S0 runs
        const int n_frames = tile.n_frames;
        const int tc = G::tceil(n_frames);
        // alignment shift of a bulk-copied tile: s = e + d, e even (absorbed as a word offset of the frame
        // columns in the staged layout), d = 0/1 (absorbed by staging y one sample ahead of x)
        const int sh = (fast || vec || pre32) ? static_cast<int>(tile.first_sample & 7) : 0;
        const int e = sh & 6, d = sh & 1;

        // ---- S0: stage y[n] = x[n] - a x[n-1] once per sample: staged index i holds sample o + d + i (o = the
        // 8-sample boundary below the tile), with kPad words inserted at i = e + k HOP, so that frame f starts
        // at word e + f STRIDE ----
        if (!stager) {
            if (fast || pre32) phase ^= 1u;
        } else if (fast && sizeof(PcmT) == 1) {
            // G.711 codes (mu-law / A-law bytes): expanded here, on the way from the raw buffer to the staged tile — the
            // PCM never exists as int16 anywhere (1 byte per sample over PCIe and from HBM)
            mbar_wait(bar, phase);
            phase ^= 1u;
            const uint8_t *raw8 = reinterpret_cast<const uint8_t *>(raw16) + 16 + (static_cast<int>(tile.first_sample) & 8);   // raw8[i] = code of sample o + i
            const int nchunks = G::tceil_s(n_frames, sh) >> 3;
            const float na = -a.preemph;
            const int alaw = a.alaw;
#pragma unroll 1
            for (int c = tid; c < nchunks; c += kStage) {
                const uint2 q = *reinterpret_cast<const uint2 *>(raw8 + 8 * c);
                uint32_t pv = raw8[8 * c + (d ? 8 : -1)];      // d = 0: the code before the chunk; d = 1: the code after it
                uint32_t w0 = q.x, w1 = q.y;
                if (d) {   // odd shift: move the chunk down one byte; its old first code becomes the predecessor
                    const uint32_t first = w0 & 0xFFu;
                    w0 = __funnelshift_r(w0, w1, 8);
                    w1 = __funnelshift_r(w1, pv, 8);
                    pv = first;
                }
                const float xe = g711_to_f32(pv, alaw);
                const float x0 = g711_to_f32(w0 & 0xFFu, alaw), x1 = g711_to_f32((w0 >> 8) & 0xFFu, alaw);
                const float x2 = g711_to_f32((w0 >> 16) & 0xFFu, alaw), x3 = g711_to_f32(w0 >> 24, alaw);
                const float x4 = g711_to_f32(w1 & 0xFFu, alaw), x5 = g711_to_f32((w1 >> 8) & 0xFFu, alaw);
                const float x6 = g711_to_f32((w1 >> 16) & 0xFFu, alaw), x7 = g711_to_f32(w1 >> 24, alaw);
                const int k = c / (HOP / 8);
                const int pad = kPad * k;
                const int pad_first = (c == k * (HOP / 8) && k > 0) ? pad - kPad : pad;
                float *dst = staged + 8 * c;
                *reinterpret_cast<float2 *>(dst + (0 < e ? pad_first : pad)) = make_float2(fmaf(na, xe, x0), fmaf(na, x0, x1));
                *reinterpret_cast<float2 *>(dst + 2 + (2 < e ? pad_first : pad)) = make_float2(fmaf(na, x1, x2), fmaf(na, x2, x3));
                *reinterpret_cast<float2 *>(dst + 4 + (4 < e ? pad_first : pad)) = make_float2(fmaf(na, x3, x4), fmaf(na, x4, x5));
                *reinterpret_cast<float2 *>(dst + 6 + pad) = make_float2(fmaf(na, x5, x6), fmaf(na, x6, x7));
            }
            if (tile.first_sample == tile.utt_begin) {
                if (tid == 0) staged[e] = g711_to_f32(raw8[sh], alaw);   // the utterance's first sample has no predecessor: y = x
            }
#if MFCC_SP_S0V2
        } else if (fast) {
            mbar_wait(bar, phase);
            phase ^= 1u;
            const int nchunks = (MFCC_SP_ABLATE & 1) ? 0 : G::tceil_s(n_frames, sh) >> 3;
            const float na = -a.preemph;
            // chunk c = 8 samples; the thread takes chunks tid, tid + kStage, ...: all loads first, then the arithmetic
            constexpr int NU = (G::TCEIL / 8 + kStage - 1) / kStage;
            uint4 q[NU];
            uint32_t pvs[NU];
#pragma unroll
            for (int u = 0; u < NU; ++u) {
                const int c = tid + u * kStage;
                if (c < nchunks) {
                    q[u] = *reinterpret_cast<const uint4 *>(raw16 + 8 + 8 * c);
                    pvs[u] = static_cast<uint16_t>(raw16[7 + 9 * d + 8 * c]);   // d = 0: the sample before the chunk; d = 1: the sample after it
                }
            }
#pragma unroll
            for (int u = 0; u < NU; ++u) {
                const int c = tid + u * kStage;
                if (c < nchunks) {
                    uint32_t pv = pvs[u];
                    uint32_t w0 = q[u].x, w1 = q[u].y, w2 = q[u].z, w3 = q[u].w;
                    if (d) {   // odd shift (uniform): move the chunk down one half-word; its old first sample becomes the predecessor
                        const uint32_t first = w0;
                        w0 = __byte_perm(w0, w1, 0x5432);
                        w1 = __byte_perm(w1, w2, 0x5432);
                        w2 = __byte_perm(w2, w3, 0x5432);
                        w3 = __byte_perm(w3, pv, 0x5432);
                        pv = first;
                    }
#if MFCC_SP_I2F
                    // conversion pipe: one I2F per sample (16 lanes/clk, otherwise idle) instead of XOR/2 + PRMT + FADD
                    const float2 x01 = make_float2(static_cast<float>(static_cast<int16_t>(w0)), static_cast<float>(static_cast<int32_t>(w0) >> 16));
                    const float2 x23 = make_float2(static_cast<float>(static_cast<int16_t>(w1)), static_cast<float>(static_cast<int32_t>(w1) >> 16));
                    const float2 x45 = make_float2(static_cast<float>(static_cast<int16_t>(w2)), static_cast<float>(static_cast<int32_t>(w2) >> 16));
                    const float2 x67 = make_float2(static_cast<float>(static_cast<int16_t>(w3)), static_cast<float>(static_cast<int32_t>(w3) >> 16));
                    const float xe = static_cast<float>(static_cast<int16_t>(pv));
#else
                    const float2 x01 = s16x2_to_f32(w0), x23 = s16x2_to_f32(w1);
                    const float2 x45 = s16x2_to_f32(w2), x67 = s16x2_to_f32(w3);
                    const float xe = s16x2_to_f32(pv).x;
#endif
                    const float2 y0 = make_float2(fmaf(na, xe, x01.x), fmaf(na, x01.x, x01.y));
                    const float2 y2 = make_float2(fmaf(na, x01.y, x23.x), fmaf(na, x23.x, x23.y));
                    const float2 y4 = make_float2(fmaf(na, x23.y, x45.x), fmaf(na, x45.x, x45.y));
                    const float2 y6 = make_float2(fmaf(na, x45.y, x67.x), fmaf(na, x67.x, x67.y));
                    // hop-block padding: block k = [e + k HOP, e + (k + 1) HOP).  Chunk c = (HOP / 8) k + m lies in
                    // block k, except the words below e of the chunks with m = 0, which still belong to block k - 1:
                    // one chunk in HOP / 8, and only when the tile is shifted (e != 0) — a branch, not four selects
                    const int k = c / (HOP / 8);
                    float *dst = staged + 8 * c + kPad * k;
                    if (e != 0 && c == k * (HOP / 8) && k > 0) {
                        float *dl = dst - kPad;
                        *reinterpret_cast<float2 *>(0 < e ? dl : dst) = y0;
                        *reinterpret_cast<float2 *>((2 < e ? dl : dst) + 2) = y2;
                        *reinterpret_cast<float2 *>((4 < e ? dl : dst) + 4) = y4;
                        *reinterpret_cast<float2 *>(dst + 6) = y6;
                    } else {
                        *reinterpret_cast<float2 *>(dst) = y0;
                        *reinterpret_cast<float2 *>(dst + 2) = y2;
                        *reinterpret_cast<float2 *>(dst + 4) = y4;
                        *reinterpret_cast<float2 *>(dst + 6) = y6;
                    }
                }
            }
            // the utterance's first sample has no predecessor: y = x.  It is word e of chunk 0 (thread 0 wrote it
            // just above), sample raw16[8 + sh].  A branch, not a predicate: one tile in 32 starts an utterance.
            if (tile.first_sample == tile.utt_begin) {
                if (tid == 0) staged[e] = static_cast<float>(raw16[8 + sh]);
            }
#else
        } else if (fast) {
            mbar_wait(bar, phase);
            phase ^= 1u;
            const int nchunks = (MFCC_SP_ABLATE & 1) ? 0 : G::tceil_s(n_frames, sh) >> 3;
            const float na = -a.preemph;
            // chunk c = 8 samples; the thread takes chunks tid, tid + 256, ...: all loads first, then the arithmetic
            constexpr int NU = (G::TCEIL / 8 + kStage - 1) / kStage;
            uint4 q[NU];
#pragma unroll
            for (int u = 0; u < NU; ++u) {
                const int c = tid + u * kStage;
                if (c < nchunks) q[u] = *reinterpret_cast<const uint4 *>(raw16 + 8 + 8 * c);
            }
#pragma unroll
            for (int u = 0; u < NU; ++u) {
                const int c = tid + u * kStage;
                if (c < nchunks) {
                    // d = 0: the sample before the chunk; d = 1: the sample after it (low half-word of pv)
                    uint32_t pv = static_cast<uint16_t>(raw16[7 + 9 * d + 8 * c]);
                    uint32_t w0 = q[u].x, w1 = q[u].y, w2 = q[u].z, w3 = q[u].w;
                    if (d) {   // odd shift: move the chunk down one half-word; its old first sample becomes the predecessor
                        const uint32_t first = w0;
                        w0 = __byte_perm(w0, w1, 0x5432);
                        w1 = __byte_perm(w1, w2, 0x5432);
                        w2 = __byte_perm(w2, w3, 0x5432);
                        w3 = __byte_perm(w3, pv, 0x5432);
                        pv = first;
                    }
                    const float2 x01 = s16x2_to_f32(w0), x23 = s16x2_to_f32(w1);
                    const float2 x45 = s16x2_to_f32(w2), x67 = s16x2_to_f32(w3);
                    // hop-block padding: block k = [e + k HOP, e + (k + 1) HOP).  Chunk c = (HOP / 8) k + m lies in
                    // block k, except the words below e of the chunks with m = 0, which still belong to block k - 1.
                    const int k = c / (HOP / 8);
                    const int pad = kPad * k;
                    const int pad_first = (c == k * (HOP / 8) && k > 0) ? pad - kPad : pad;
                    float *dst = staged + 8 * c;
                    float *d0 = dst + (0 < e ? pad_first : pad);
                    float *d2 = dst + 2 + (2 < e ? pad_first : pad);
                    float *d4 = dst + 4 + (4 < e ? pad_first : pad);
                    float *d6 = dst + 6 + pad;
                    const float xe = s16x2_to_f32(pv).x;
                    float2 y0 = make_float2(fmaf(na, xe, x01.x), fmaf(na, x01.x, x01.y));
                    float2 y2 = make_float2(fmaf(na, x01.y, x23.x), fmaf(na, x23.x, x23.y));
                    float2 y4 = make_float2(fmaf(na, x23.y, x45.x), fmaf(na, x45.x, x45.y));
                    float2 y6 = make_float2(fmaf(na, x45.y, x67.x), fmaf(na, x67.x, x67.y));
                    // Chunks are 32 bytes apart, so in one 8-byte store the lanes j, j + 4, j + 8, ... of a half-warp all hit
                    // the same bank pair (4-way conflict: 11.9 M of the 13.9 M excessive wavefronts of a configs[1] launch).
                    // MFCC_SP_S0ROT: lanes rotate the ORDER in which they write their four pieces by (lane / 4) mod 4 (2: full
                    // rotation, conflict-free; 1: halves swapped for lanes with bit 2 set, 2-way) — selects, no extra stores.
                    if constexpr (MFCC_SP_S0ROT >= 1) {
                        const bool r2 = MFCC_SP_S0ROT >= 2 ? (lane & 8) != 0 : (lane & 4) != 0;
                        const float2 t0 = r2 ? y4 : y0, t2 = r2 ? y6 : y2, t4 = r2 ? y0 : y4, t6 = r2 ? y2 : y6;
                        float *e0 = r2 ? d4 : d0, *e2 = r2 ? d6 : d2, *e4 = r2 ? d0 : d4, *e6 = r2 ? d2 : d6;
                        y0 = t0; y2 = t2; y4 = t4; y6 = t6;
                        d0 = e0; d2 = e2; d4 = e4; d6 = e6;
                        if constexpr (MFCC_SP_S0ROT >= 2) {
                            const bool r1 = (lane & 4) != 0;
                            const float2 u0 = r1 ? y2 : y0, u2 = r1 ? y4 : y2, u4 = r1 ? y6 : y4, u6 = r1 ? y0 : y6;
                            float *g0 = r1 ? d2 : d0, *g2 = r1 ? d4 : d2, *g4 = r1 ? d6 : d4, *g6 = r1 ? d0 : d6;
                            y0 = u0; y2 = u2; y4 = u4; y6 = u6;
                            d0 = g0; d2 = g2; d4 = g4; d6 = g6;
                        }
                    }
                    *reinterpret_cast<float2 *>(d0) = y0;
                    *reinterpret_cast<float2 *>(d2) = y2;
                    *reinterpret_cast<float2 *>(d4) = y4;
                    *reinterpret_cast<float2 *>(d6) = y6;
                }
            }
            // the utterance's first sample has no predecessor: y = x.  It is word e of chunk 0 (thread 0 wrote it
            // just above), sample raw16[8 + sh].  A branch, not a predicate: one tile in 32 starts an utterance.
            if (tile.first_sample == tile.utt_begin) {
                if (tid == 0) staged[e] = static_cast<float>(raw16[8 + sh]);
            }
#endif
        } else if (pre32) {
            if constexpr (sizeof(PcmT) == 4) {
                mbar_wait(bar, phase);
                phase ^= 1u;
                const float *x = raw32 + 8;                     // x[i] = sample o + i, o = the 8-sample boundary below the tile
                const bool has_before = tile.first_sample - sh > 0;
                const int nchunks = G::tceil_s(n_frames, sh) >> 3;
                const float na = -a.preemph;
#pragma unroll 1
                for (int c = tid; c < nchunks; c += kStage) {
                    const float4 lo4 = lds_f4(x + 8 * c), hi4 = lds_f4(x + 8 * c + 4);
                    // d = 0: the sample before the chunk; d = 1: the sample after it (inside the copied span + 1: finite, unused)
                    const float nb = (d || c > 0 || has_before) ? x[8 * c + (d ? 8 : -1)] : 0.0f;
                    const float v0 = d ? lo4.x : nb, v1 = d ? lo4.y : lo4.x, v2 = d ? lo4.z : lo4.y, v3 = d ? lo4.w : lo4.z;
                    const float v4 = d ? hi4.x : lo4.w, v5 = d ? hi4.y : hi4.x, v6 = d ? hi4.z : hi4.y, v7 = d ? hi4.w : hi4.z;
                    const float v8 = d ? nb : hi4.w;        // staged word j = v[j + 1] - a v[j]
                    const int k = c / (HOP / 8);
                    const int pad = kPad * k;
                    const int pad_first = (c == k * (HOP / 8) && k > 0) ? pad - kPad : pad;
                    float *dst = staged + 8 * c;
                    *reinterpret_cast<float2 *>(dst + (0 < e ? pad_first : pad)) = make_float2(fmaf(na, v0, v1), fmaf(na, v1, v2));
                    *reinterpret_cast<float2 *>(dst + 2 + (2 < e ? pad_first : pad)) = make_float2(fmaf(na, v2, v3), fmaf(na, v3, v4));
                    *reinterpret_cast<float2 *>(dst + 4 + (4 < e ? pad_first : pad)) = make_float2(fmaf(na, v4, v5), fmaf(na, v5, v6));
                    *reinterpret_cast<float2 *>(dst + 6 + pad) = make_float2(fmaf(na, v6, v7), fmaf(na, v7, v8));
                }
                if (tile.first_sample == tile.utt_begin) {
                    if (tid == 0) staged[e] = x[sh];           // the utterance's first sample has no predecessor: y = x
                }
            }
        } else if (vec) {
            if constexpr (sizeof(PcmT) == 4) {
                const float *x = reinterpret_cast<const float *>(pcm) + (tile.first_sample - sh);   // x[i] = sample o + i
                const bool has_before = tile.first_sample - sh > 0;
                const int nchunks = G::tceil_s(n_frames, sh) >> 3;
                const float na = -a.preemph;
#pragma unroll 1
                for (int c = tid; c < nchunks; c += kStage) {
                    const float4 lo4 = __ldg(reinterpret_cast<const float4 *>(x) + 2 * c);
                    const float4 hi4 = __ldg(reinterpret_cast<const float4 *>(x) + 2 * c + 1);
                    // d = 0: the sample before the chunk; d = 1: the sample after it (inside the span the host checked)
                    const float nb = (d || c > 0 || has_before) ? __ldg(x + 8 * c + (d ? 8 : -1)) : 0.0f;
                    const float v0 = d ? lo4.x : nb, v1 = d ? lo4.y : lo4.x, v2 = d ? lo4.z : lo4.y, v3 = d ? lo4.w : lo4.z;
                    const float v4 = d ? hi4.x : lo4.w, v5 = d ? hi4.y : hi4.x, v6 = d ? hi4.z : hi4.y, v7 = d ? hi4.w : hi4.z;
                    const float v8 = d ? nb : hi4.w;        // staged word j = v[j + 1] - a v[j]
                    const int k = c / (HOP / 8);
                    const int pad = kPad * k;
                    const int pad_first = (c == k * (HOP / 8) && k > 0) ? pad - kPad : pad;
                    float *dst = staged + 8 * c;
                    *reinterpret_cast<float2 *>(dst + (0 < e ? pad_first : pad)) = make_float2(fmaf(na, v0, v1), fmaf(na, v1, v2));
                    *reinterpret_cast<float2 *>(dst + 2 + (2 < e ? pad_first : pad)) = make_float2(fmaf(na, v2, v3), fmaf(na, v3, v4));
                    *reinterpret_cast<float2 *>(dst + 4 + (4 < e ? pad_first : pad)) = make_float2(fmaf(na, v4, v5), fmaf(na, v5, v6));
                    *reinterpret_cast<float2 *>(dst + 6 + pad) = make_float2(fmaf(na, v6, v7), fmaf(na, v7, v8));
                }
                if (tile.first_sample == tile.utt_begin) {
                    if (tid == 0) staged[e] = x[sh];   // the utterance's first sample has no predecessor: y = x
                }
            }
        } else {
            const int64_t room_lo = tile.first_sample - tile.utt_begin;
            const int64_t room_hi = tile.utt_end - tile.first_sample;
            const PcmT *x = pcm + tile.first_sample;
            for (int i = tid; i < tc; i += kStage) {
                float y = 0.0f;
                if (i < room_hi) {
                    const float x0 = to_f32(x[i], a.alaw);
                    const float x1 = (i > -room_lo) ? to_f32(x[i - 1], a.alaw) : 0.0f;
                    y = fmaf(-a.preemph, x1, x0);
                }
                staged[G::padded(i)] = y;
            }
        }
        cp_async_wait_all();
        lock_at(1);
        half_sync(half);   // B1: staged complete; raw buffer and (previous tile's) scratch free; next descriptor visible
        unlock_at(1);
        if constexpr (MFCC_POISON) {   // the raw buffer is dead until the next bulk copy lands; so is the tail scratch
            poison(mine + G::UNION + G::WS, G::RAW);
            poison(scr, G::WS);
            half_sync(half);
        }
        if (has_next && tid == 0) {
            const Tile nt = desc[cur ^ 1];
            if (tile_fast(nt)) issue_copy(nt);
        }

        // ---- S1: pass 1.  Warp = column pair (a, a + 1): windowed real DFT-RB over b, inter-pass twiddle ----
        if constexpr ((MFCC_SP_ABLATE & 2) == 0) {
            const int pr = warp;
            const float *base = staged + e + lane * STRIDE + 2 * pr;
            [[maybe_unused]] const float *wrow = t_win + pr * (2 * G::NZP);
            [[maybe_unused]] const int upr = __shfl_sync(0xffffffffu, pr, 0);
            float2 in[NZ];
#pragma unroll
            for (int b = 0; b < NZ; b += 2) {
                const float4 w = G::CONST_TAB ? a.winc[upr][b / 2] : lds_f4(wrow + 2 * b);
                const float2 y0 = lds_f2(base + G::padded(RA * b));
                in[b] = make_float2(y0.x * w.x, y0.y * w.y);
                if (b + 1 < NZ) {
                    const float2 y1 = lds_f2(base + G::padded(RA * (b + 1)));
                    in[b + 1] = make_float2(y1.x * w.z, y1.y * w.w);
                }
            }
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int col = 2 * pr + hh;
                float x[RB];
#pragma unroll
                for (int b = 0; b < RB; ++b) x[b] = b < NZ ? (hh ? in[b < NZ ? b : 0].y : in[b < NZ ? b : 0].x) : 0.0f;
                rf::cplx X[H + 1];
                rf::RDft<RB>::template run<NZ>(x, X);
                float2 *wsa = ws + col * 32 + lane;
                wsa[(H - 1) * RA * 32] = make_float2(X[0].re, X[H].re);   // rows 0 and H are real here
                if constexpr (G::TW_IN_PASS2) {
                    // rows 1 .. H - 1 go out UNtwiddled: pass 2 applies the inter-pass twiddle on load
#pragma unroll
                    for (int k1 = 1; k1 < H; ++k1) wsa[(k1 - 1) * RA * 32] = make_float2(X[k1].re, X[k1].im);
                } else {
                    [[maybe_unused]] const float *trow = t_tw + col * (2 * H);
#pragma unroll
                    for (int k1 = 1; k1 < H; k1 += 2) {
                        const float4 tw = G::CONST_TAB ? a.twc[(2 * upr + hh) & 15][(k1 - 1) / 2]
                                                        : lds_f4(trow + 2 * (k1 - 1));       // twiddles of k1, k1 + 1
                        const rf::cplx v = rf::cmulc(X[k1], tw.x, tw.y);
                        wsa[(k1 - 1) * RA * 32] = make_float2(v.re, v.im);
                        if (k1 + 1 < H) {
                            const rf::cplx u = rf::cmulc(X[k1 + 1], tw.z, tw.w);
                            wsa[k1 * RA * 32] = make_float2(u.re, u.im);
                        }
                    }
                }
            }
        }
        lock_at(2);
        half_sync(half);   // B2
        unlock_at(2);
        if constexpr (MFCC_POISON) {   // the staged samples are dead: pass 2 writes P over them (slack rows stay zero)
            poison(staged, G::UNION);
            half_sync(half);
            for (int i = G::NB * 32 + tid; i < G::PW; i += kHalfThreads) pw[i] = 0.0f;
            half_sync(half);
        }

        // ---- S2: pass 2.  Item = one row k1: complex DFT-RA over a gives bins k1 + RB k2; power ----
        if constexpr ((MFCC_SP_ABLATE & 4) == 0) {
#pragma unroll 1
            for (int it = warp; it < H; it += kWarps) {
                if (it < H - 1) {
                    const int k1 = it + 1;
                    const float2 *row = ws + (k1 - 1) * RA * 32 + lane;
                    rf::cplx z[RA];
                    if constexpr (G::TW_IN_PASS2) {
                        [[maybe_unused]] const float *trow = t_tw + (k1 - 1) * (2 * RA);      // W_N^(a k1), a = 0 .. RA - 1
                        [[maybe_unused]] const int uk = __shfl_sync(0xffffffffu, k1 - 1, 0) & 15;
#pragma unroll
                        for (int c = 0; c < RA; c += 2) {
                            const float4 tw = G::CONST_TW2 ? a.twc[uk][c / 2] : lds_f4(trow + 2 * c);
                            const float2 p = row[c * 32], q = row[(c + 1) * 32];
                            z[c] = c == 0 ? rf::cplx{p.x, p.y} : rf::cmulc(rf::cplx{p.x, p.y}, tw.x, tw.y);
                            z[c + 1] = rf::cmulc(rf::cplx{q.x, q.y}, tw.z, tw.w);
                        }
                    } else {
#pragma unroll
                        for (int c = 0; c < RA; ++c) {
                            const float2 p = row[c * 32];
                            z[c] = rf::cplx{p.x, p.y};
                        }
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
                        constexpr bool kConstTwh = ((MFCC_SP_CONST_TWH) >> (RB == 16 ? 1 : 0)) & 1;
                        const float4 tw = kConstTwh ? a.twhc[c / 2] : lds_f4(t_twh + 2 * c);
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
        lock_at(3);
        half_sync(half);   // B3: P complete, workspace free
        unlock_at(3);
        if constexpr (MFCC_POISON) {   // the workspace is dead: S3 writes every segment's two sums into it
            poison(scr, G::WS);
            half_sync(half);
        }
        if constexpr (sizeof(PcmT) == 4) {   // f32 PCM: the next tile's samples land in the idle end of the workspace
            if (has_next && tid == 0) {
                const Tile nt = desc[cur ^ 1];
                if (tile_pre32(nt)) issue_copy32(nt);
            }
        }

        // ---- S3: filterbank sums.  Segment j = bins [b_j, b_j + w) rises into filter j with weight i / w and
        // falls out of filter j - 1 with weight (w - i) / w (i = bin - b_j), so two plain sums per segment,
        // S = sum P and T = sum i P, give both: rise = s T, fall = s (w S - T), s = 1 / (w NFFT).  No weight loads,
        // and the P addresses do not depend on loaded data.  Every segment is walked ONCE, by the warp the host
        // gave it to (longest-first assignment, so the eight walks cost the same); rise[j][lane] and fall[j][lane]
        // go through the scratch and S3b adds the two halves of each band. ----
        if constexpr ((MFCC_SP_ABLATE & 8) != 0) {
        } else if constexpr (kTail) {
            // Forward walk with running sums: run_i = P_0 + .. + P_i and acc = run_0 + .. + run_(w-1) = sum (w - i) P_i,
            // so fall = s acc and rise = S / NFFT - fall: one load and two additions per bin, no constants, no counter.
            // (rise is formed by subtraction: a bin carries weight i / w in it, and the subtraction leaves an error of
            // eps P_i, i.e. eps w / i relative to that bin's contribution — harmless except for a bin of weight 0 (i = 0)
            // that towers over its neighbours by more than 1 / (eps w).  The frames of this kernel are shorter than the
            // transform (L < NFFT), so even a bin-centred tone under a rectangular window leaks at the -13 dB level into the
            // neighbouring bins: the bound is eps w P_peak / P_neighbour ~ 3e-5 relative, and
            // tests/test_gpu_parity.py::test_bin_centred_tone_known_answer holds every band to the stated tolerance there.)
            // The descriptors come from the parameter bank under a warp-uniform index (uniform registers, uniform
            // branches); the width is taken apart in binary (8-bin loop, then 4, 2, 1).
            const int uw = __shfl_sync(0xffffffffu, warp, 0);
            const char *pl = reinterpret_cast<const char *>(pw + lane);
            char *rl = reinterpret_cast<char *>(scr + a.rf + lane);
#pragma unroll 1
            for (int q = 0;; ++q) {
                const float4 sg = a.useg[uw][q];
                const int oo = __float_as_int(sg.w);
                if (oo < 0) break;
                const float *p = reinterpret_cast<const float *>(pl + __float_as_int(sg.x));
                const int w = __float_as_int(sg.y);
                float run = 0.0f, acc = 0.0f;
#pragma unroll 1
                for (int c = w >> 3; c > 0; --c) {
                    const float a0 = p[0], a1 = p[32], a2 = p[64], a3 = p[96];
                    const float a4 = p[128], a5 = p[160], a6 = p[192], a7 = p[224];
                    run += a0; acc += run;
                    run += a1; acc += run;
                    run += a2; acc += run;
                    run += a3; acc += run;
                    run += a4; acc += run;
                    run += a5; acc += run;
                    run += a6; acc += run;
                    run += a7; acc += run;
                    p += 256;
                }
                if (w & 4) {
                    const float a0 = p[0], a1 = p[32], a2 = p[64], a3 = p[96];
                    run += a0; acc += run;
                    run += a1; acc += run;
                    run += a2; acc += run;
                    run += a3; acc += run;
                    p += 128;
                }
                if (w & 2) {
                    const float a0 = p[0], a1 = p[32];
                    run += a0; acc += run;
                    run += a1; acc += run;
                    p += 64;
                }
                if (w & 1) {
                    run += p[0];
                    acc += run;
                }
                const float f = sg.z * acc;
                *reinterpret_cast<float *>(rl + oo) = fmaf(a.inv_n, run, -f);
                *reinterpret_cast<float *>(rl + oo + (MEL + 3) * 128) = f;
            }
        } else
        {
            const float4 *wd = t_wseg + warp * kSegMax;
            float *rise = scr + a.rf + lane, *fall = rise + (a.n_mel + 3) * 32;
            float4 nx = wd[0];                              // descriptors run one segment ahead of their use
#pragma unroll 1
            for (int q = 1; __float_as_int(nx.w) >= 0; ++q) {
                const float4 sg = nx;                       // {first bin * 32, w, s, j}
                nx = wd[q];                                 // (the list ends with a j = -1 entry)
                const float *p = pw + __float_as_int(sg.x) + lane;
                const int w = __float_as_int(sg.y), j = __float_as_int(sg.w);
                int c = w >> 2;
                float S = 0.0f, T = 0.0f, i0 = 0.0f;
#pragma unroll 1
                for (; c >= 2; c -= 2) {
                    const float a0 = p[0], a1 = p[32], a2 = p[64], a3 = p[96];
                    const float b0 = p[128], b1 = p[160], b2 = p[192], b3 = p[224];
                    const float sa = (a0 + a1) + (a2 + a3), sb = (b0 + b1) + (b2 + b3);
                    const float ta = fmaf(3.0f, a3, fmaf(2.0f, a2, a1)), tb = fmaf(3.0f, b3, fmaf(2.0f, b2, b1));
                    T = fmaf(i0, sa, T) + ta;
                    T = fmaf(i0 + 4.0f, sb, T) + tb;
                    S += sa + sb;
                    i0 += 8.0f;
                    p += 256;
                }
                if (c) {
                    const float a0 = p[0], a1 = p[32], a2 = p[64], a3 = p[96];
                    const float sa = (a0 + a1) + (a2 + a3);
                    T = fmaf(i0, sa, T) + fmaf(3.0f, a3, fmaf(2.0f, a2, a1));
                    S += sa;
                    i0 += 4.0f;
                    p += 128;
                }
                if (const int e = w & 3) {                  // 1..3 leftover bins, no loop: the rows past the segment are
                    const float a0 = p[0];                  // finite (next segment or the zeroed slack rows) and deselected
                    const float a1 = e > 1 ? p[32] : 0.0f;
                    const float a2 = e > 2 ? p[64] : 0.0f;
                    const float sa = (a0 + a1) + a2;
                    T = fmaf(i0, sa, T) + fmaf(2.0f, a2, a1);
                    S += sa;
                }
                const float r = sg.z * T;
                rise[j * 32] = r;
                fall[j * 32] = fmaf(a.inv_n, S, -r);
            }
        }
        lock_at(4);
        half_sync(half);   // B4a: every segment's two sums are in the scratch
        unlock_at(4);
        if constexpr (MFCC_POISON) {   // P is dead: the next S0 stages over it
            poison(pw, G::NB * 32);
            half_sync(half);
        }
        if constexpr (kTail) {
            // the tail warp takes it from here during the next tile's S0 (or after the loop): the scratch is next
            // written by S1, after B1, which the tail warp joins only when it is done
            prev_row = tile.out_row;
            prev_nf = n_frames;
        } else {
        // ---- S3b: band m = rise of segment m + fall of segment m + 1; log -> lg[m][lane] or the frame's log-mel row ----
        {
            const float *rise = scr + a.rf, *fall = rise + (a.n_mel + 3) * 32;
            const int total = a.n_mel * 32;
            for (int i = tid; i < total; i += kHalfThreads) {
                const float lg = kLn2 * lg2_fast(fmaxf(rise[i] + fall[i + 32], a.log_floor));
                if (a.logmel) scr[(i & 31) * a.ls + (i >> 5)] = lg;
                else scr[i] = lg;
            }
            // frame energy: all segment sums of the frame (lane = frame), the last warp's job
            if (a.energy != MFCC_ENERGY_NONE && warp == kWarps - 1) {
                float e0 = 0.0f, e1 = 0.0f;
                for (int j = 0; j < a.n_mel + 3; ++j) {
                    e0 += rise[j * 32 + lane];
                    e1 += fall[j * 32 + lane];
                }
                const float le = kLn2 * lg2_fast(fmaxf(e0 + e1, a.log_floor));
                if (a.logmel) scr[lane * a.ls + a.n_mel] = le;
                else scr[a.ef + lane] = le;
            }
        }
        half_sync(half);   // B4: every band's log energy is in the scratch

        // ---- S4: log-mel rows are copied out coalesced; cepstra: warp w forms c[w] and c[w + 8] of frame = lane
        // from the log energies (conflict-free column reads, warp-uniform DCT entries) and stores them ----
        if (a.logmel) {
            const int M = a.od, total = n_frames * M;     // n_mel columns, + the energy column under MFCC_ENERGY_APPEND
            float *o = a.out + tile.out_row * M;
            for (int i = tid; i < total; i += kHalfThreads) {
                const int f = (i * a.mel_magic) >> 20, m = i - f * M;
                o[i] = scr[f * a.ls + m];
            }
        } else {
            // 16 cepstra per round: warp w forms c[kb + w] and c[kb + w + 8]
#pragma unroll 1
            for (int kb = 0; kb + warp < a.n_cep; kb += 2 * kWarps) {
                const float *lg = scr + lane;
                const float4 *dc = reinterpret_cast<const float4 *>(t_dct + ((kb >> 4) * kWarps + warp) * (2 * a.mp));   // {d[k][m], d[k+8][m], d[k][m+1], d[k+8][m+1]}
                float c0 = 0.0f, c1 = 0.0f, e0 = 0.0f, e1 = 0.0f;
                const int M2 = a.n_mel >> 1;
#pragma unroll 2
                for (int q = 0; q < M2; ++q) {
                    const float l0 = lg[(2 * q) * 32], l1 = lg[(2 * q + 1) * 32];
                    const float4 d = dc[q];
                    c0 = fmaf(d.x, l0, c0);
                    c1 = fmaf(d.y, l0, c1);
                    e0 = fmaf(d.z, l1, e0);
                    e1 = fmaf(d.w, l1, e1);
                }
                if (a.n_mel & 1) {
                    const float l0 = lg[(a.n_mel - 1) * 32];
                    const float4 d = dc[M2];
                    c0 = fmaf(d.x, l0, c0);
                    c1 = fmaf(d.y, l0, c1);
                }
                if (lane < n_frames) {
                    float *o = a.out + (tile.out_row + lane) * a.od + kb + warp;
                    o[0] = c0 + e0;
                    if (kb + warp + kWarps < a.n_cep) o[kWarps] = c1 + e1;
                }
            }
            // energy term: warp 0 formed c[0] of its frames above (same thread, program order), so it replaces it here
            if (a.energy != MFCC_ENERGY_NONE && warp == 0 && lane < n_frames) {
                float *o = a.out + (tile.out_row + lane) * a.od;
                o[a.energy == MFCC_ENERGY_REPLACE_C0 ? 0 : a.n_cep] = scr[a.ef + lane];
            }
        }
        }
        // no barrier here: the next S0 writes `staged`, which nobody reads any more; the scratch is
        // next written by S1, after B1.
    }
    // a lock whose span wraps around the loop is still held after the last tile
    for (int i = 0; i < 2; ++i)
        if (held[i]) {
            if constexpr (MFCC_SP_LOCK_HOLDERS <= 1) atomicExch(&phase_lock[i], 0);
            else atomicSub(&phase_lock[i], 1);
        }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
struct SpVariant {
    int L, hop, rb, ra;
    const char *name;
};
constexpr SpVariant kVariants[] = {
    {400, 160, 32, 16, "fused_sp_tile32_L400_H160_real32x16"},   // BASELINE.json configs 1, 2, 5 (16 kHz)
    {200, 80, 16, 16, "fused_sp_tile32_L200_H80_real16x16"},     // BASELINE.json config 3 (8 kHz telephony)
};

struct SpState {
    const SpVariant *v = nullptr;
    SpArgs args{};             // per-launch fields blank
    float *d_tab = nullptr;
    size_t smem = 0;
    int sm_count = 0;
    int device = 0;
    bool useg_ok = false;      // every warp's segment list fits SpArgs::useg (tail-warp variants need it)
};

const SpVariant *find_variant(const mfcc_params &p)
{
    for (const auto &v : kVariants)
        if (p.frame_len == v.L && p.hop_len == v.hop && p.nfft == v.rb * v.ra) return &v;
    return nullptr;
}

template <int L, int HOP, int RB, int RA>
void geo_sizes(int &tabf, int &half_floats) { tabf = Geo<L, HOP, RB, RA>::TABF; half_floats = Geo<L, HOP, RB, RA>::HALF; }

void variant_sizes(const SpVariant &v, int &tabf, int &half_floats)
{
    if (v.L == 400) geo_sizes<400, 160, 32, 16>(tabf, half_floats);
    else geo_sizes<200, 80, 16, 16>(tabf, half_floats);
}

// One walk of S3: bins [k0, k0 + w) summed into scratch row j.  Rows 0 .. M are the filterbank's segments; rows M + 1 and
// M + 2 are the pseudo-segments below and above the filterbank (s = 0: plain sums), walked only when the plan has an
// energy term.
struct Seg {
    int k0, w, row;
    float s;        // 1 / (w NFFT): pass 2 leaves |X|^2, hence the 1 / N
};

std::vector<Seg> plan_segments(const mfcc_params &p, const HostTables &h)
{
    const int M = p.n_mel, N = p.nfft;
    std::vector<Seg> segs;
    for (int j = 0; j <= M; ++j) {
        const int k0 = h.mel_bins[j], w = h.mel_bins[j + 1] - k0;
        segs.push_back({k0, w, j, w > 0 ? static_cast<float>(1.0 / (static_cast<double>(w) * N)) : 0.0f});
    }
    if (p.energy != MFCC_ENERGY_NONE) {
        segs.push_back({0, h.mel_bins[0], M + 1, 0.0f});
        segs.push_back({h.mel_bins[M + 1], h.nbins - h.mel_bins[M + 1], M + 2, 0.0f});
    }
    return segs;
}

// Instruction estimate of one walk (S3): 8-bin chunks, then 4, 2, 1 bins, fixed part.
inline int seg_cost(const Seg &g) { return 28 * (g.w / 8) + 14 * ((g.w / 4) & 1) + 8 * ((g.w / 2) & 1) + 4 * (g.w & 1) + 18; }

// Segments over the warps, longest first onto the least loaded warp.  Each warp's list is kept in ascending bin order
// (the walk then moves forward through P).  Empty when a warp would need more than kSegMax - 1 segments.
std::vector<std::vector<Seg>> assign_segments(const std::vector<Seg> &segs)
{
    std::vector<int> order(segs.size());
    for (size_t j = 0; j < segs.size(); ++j) order[j] = static_cast<int>(j);
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return seg_cost(segs[x]) > seg_cost(segs[y]); });
    std::vector<std::vector<Seg>> lists(kWarps);
    std::vector<int> load(kWarps, 0);
    for (int j : order) {
        int best = 0;
        for (int w = 1; w < kWarps; ++w)
            if (load[w] < load[best]) best = w;
        lists[best].push_back(segs[j]);
        load[best] += seg_cost(segs[j]);
    }
    for (auto &l : lists) {
        if (static_cast<int>(l.size()) > kSegMax - 1) return {};
        std::sort(l.begin(), l.end(), [](const Seg &x, const Seg &y) { return x.k0 != y.k0 ? x.k0 < y.k0 : x.row < y.row; });
    }
    return lists;
}

inline int out_stride(const mfcc_params &p) { return (p.output == MFCC_OUT_LOGMEL ? p.n_mel : p.n_cep) + (p.energy == MFCC_ENERGY_APPEND ? 1 : 0); }
inline int logmel_stride(const mfcc_params &p) { return out_stride(p) | 1; }
inline int scratch_rf(const mfcc_params &p) { return (32 * std::max(logmel_stride(p), p.n_mel) + 3) / 4 * 4; }

}  // namespace

#if MFCC_SP_PCM_TYPES & 1   // the host half is compiled once, with the int16 entry
const char *sp_match(const mfcc_params &p, const HostTables &h)
{
    const SpVariant *v = find_variant(p);
    if (v == nullptr) return nullptr;
    if (p.log_floor < 1.17549435e-38f) return nullptr;   // the tail takes lg2.approx.ftz of max(E, floor): floor must be a normal float
    for (int j = 0; j + 1 < static_cast<int>(h.mel_bins.size()); ++j)
        if (h.mel_bins[j + 1] < h.mel_bins[j]) return nullptr;
    // the run-time tables must fit next to the two halves
    int tabf = 0, half_floats = 0;
    variant_sizes(*v, tabf, half_floats);
    if (p.n_mel + 3 > kWarps * (kSegMax - 1) || assign_segments(plan_segments(p, h)).empty()) return nullptr;
    const size_t rounds = p.output == MFCC_OUT_CEPSTRA ? (p.n_cep + KC - 1) / KC : 1;
    const size_t total = tabf + 4 * static_cast<size_t>(kWarps) * kSegMax + rounds * KC * (p.n_mel + 1) + 16;
    if ((total + groups_for(v->rb) * static_cast<size_t>(half_floats)) * sizeof(float) > kSmemMax) return nullptr;
    // tail scratch (log band energies [n_mel][32] or log-mel rows [32][ls], then the per-segment rise / fall sums
    // [n_mel + 3][32] x 2, then the log frame energy [32]) must fit in the workspace
    const size_t scratch = static_cast<size_t>(scratch_rf(p)) + 2 * 32 * static_cast<size_t>(p.n_mel + 3) + 32;
    if (scratch > static_cast<size_t>(v->rb / 2) * v->ra * 32 * 2) return nullptr;
    return v->name;
}

int sp_prepare(mfcc_plan *plan)
{
    const mfcc_params &p = plan->p;
    const HostTables &h = plan->host;
    const SpVariant *var = find_variant(p);
    if (var == nullptr || sp_match(p, h) == nullptr) return MFCC_ENOTSUP;
    const int RB = var->rb, RA = var->ra, N = RB * RA, H = RB / 2, M = p.n_mel;
    const int NZ = (p.frame_len + RA - 1) / RA, NZP = (NZ + 1) / 2 * 2;
    int tabf = 0, half_floats = 0;
    variant_sizes(*var, tabf, half_floats);
    std::vector<float> tab;
    auto align4 = [&]() { while (tab.size() % 4) tab.push_back(0.0f); };
    auto push_int = [&](int v) { float f; std::memcpy(&f, &v, 4); tab.push_back(f); };

    // fixed part, in Geo's order: window pairs, inter-pass twiddles, row-H twiddles
    for (int pr = 0; pr < RA / 2; ++pr)
        for (int b = 0; b < NZP; ++b)
            for (int e = 0; e < 2; ++e) {
                const int i = 2 * pr + e + RA * b;
                tab.push_back(b < NZ && i < p.frame_len ? h.window[i] : 0.0f);
            }
    const bool tw_in_pass2 = RB == 16 ? MFCC_SP_TW2_16 : MFCC_SP_TW2_32;   // Geo::TW_IN_PASS2
    for (int i = 0; i < RA * H; ++i) {
        const int col = tw_in_pass2 ? i % RA : i / H, sl = tw_in_pass2 ? i / RA : i % H;
        const double ang = sl < H - 1 ? -2.0 * M_PI * static_cast<double>(col) * (sl + 1) / N : 0.0;
        tab.push_back(static_cast<float>(std::cos(ang)));
        tab.push_back(static_cast<float>(std::sin(ang)));
    }
    for (int col = 0; col < RA; ++col) {
        const double ang = -2.0 * M_PI * col / (2.0 * RA);
        tab.push_back(static_cast<float>(std::cos(ang)));
        tab.push_back(static_cast<float>(std::sin(ang)));
    }
    if (static_cast<int>(tab.size()) != tabf) return MFCC_ECUDA;   // layout drifted from Geo

    SpLayout lay{};
    // per-warp segment walks.  The triangles are linear ramps over integer bins (mfcc_tables.cpp build_tables),
    // so a segment is described by its first bin, its width and 1 / (w N).
    const std::vector<std::vector<Seg>> lists = assign_segments(plan_segments(p, h));
    if (lists.empty()) return MFCC_ENOTSUP;
    lay.wseg = static_cast<int>(tab.size());
    for (int w = 0; w < kWarps; ++w)
        for (int q = 0; q < kSegMax; ++q) {
            if (q < static_cast<int>(lists[w].size())) {
                const Seg &g = lists[w][q];
                push_int(g.k0 * 32);
                push_int(g.w);
                tab.push_back(g.s);
                push_int(g.row);
            } else {
                push_int(0);
                push_int(0);
                tab.push_back(0.0f);
                push_int(-1);
            }
        }
    align4();
    // DCT entries for warp w: {d[w][m], d[w + 8][m]} per filter m, zero past n_cep; rows padded to even length
    lay.dct = static_cast<int>(tab.size());
    const int MP = (M + 1) / 2 * 2;
    auto dct_at = [&](int k, int m) {
        return p.output == MFCC_OUT_CEPSTRA && k < p.n_cep && m < M ? h.dct[static_cast<size_t>(k) * M + m] : 0.0f;
    };
    const int rounds = p.output == MFCC_OUT_CEPSTRA ? (p.n_cep + KC - 1) / KC : 1;
    for (int r = 0; r < rounds; ++r)
        for (int w = 0; w < kWarps; ++w)
            for (int m = 0; m < MP; ++m) {
                tab.push_back(dct_at(KC * r + w, m));
                tab.push_back(dct_at(KC * r + w + kWarps, m));
            }
    align4();
    lay.total = static_cast<int>(tab.size());

    SpState *st = new SpState();
    st->v = var;
    st->sm_count = plan->sm_count;
    st->device = plan->device;
    st->smem = sizeof(float) * (static_cast<size_t>(lay.total) + groups_for(RB) * static_cast<size_t>(half_floats));
    if (st->smem > kSmemMax) { delete st; return MFCC_ENOTSUP; }
    st->args.lay = lay;
    {
        const float *twh = tab.data() + RA / 2 * NZP * 2 + RA * H * 2;
        for (int q = 0; q < RA / 2 && q < 8; ++q) st->args.twhc[q] = make_float4(twh[4 * q], twh[4 * q + 1], twh[4 * q + 2], twh[4 * q + 3]);
    }
    {
        const float *win = tab.data(), *tw = tab.data() + RA / 2 * NZP * 2;
        for (int pr = 0; pr < RA / 2 && pr < kWarps; ++pr)
            for (int q = 0; q < NZP / 2 && q < 14; ++q)
                st->args.winc[pr][q] = make_float4(win[pr * 2 * NZP + 4 * q], win[pr * 2 * NZP + 4 * q + 1], win[pr * 2 * NZP + 4 * q + 2], win[pr * 2 * NZP + 4 * q + 3]);
        if (tw_in_pass2)      // rows k1 - 1 of the [H][RA] table, two columns per float4
            for (int r = 0; r < H && r < 16; ++r)
                for (int q = 0; q < RA / 2 && q < 8; ++q)
                    st->args.twc[r][q] = make_float4(tw[r * 2 * RA + 4 * q], tw[r * 2 * RA + 4 * q + 1], tw[r * 2 * RA + 4 * q + 2], tw[r * 2 * RA + 4 * q + 3]);
        else
            for (int col = 0; col < RA && col < 16; ++col)
                for (int q = 0; q < H / 2 && q < 8; ++q)
                    st->args.twc[col][q] = make_float4(tw[col * 2 * H + 4 * q], tw[col * 2 * H + 4 * q + 1], tw[col * 2 * H + 4 * q + 2], tw[col * 2 * H + 4 * q + 3]);
    }
    st->args.n_mel = M;
    st->args.n_cep = p.n_cep;
    st->args.logmel = p.output == MFCC_OUT_LOGMEL;
    st->args.energy = p.energy;
    st->args.od = out_stride(p);
    st->args.alaw = 0;
    st->args.ls = logmel_stride(p);
    st->args.mp = MP;
    st->args.rf = scratch_rf(p);
    st->args.ef = st->args.rf + 2 * 32 * (M + 3);
    {   // raw f32 buffer (8 lead + TCEIL + 8 samples) at the end of the workspace, when the tail scratch leaves room for it
        const int ws_floats = H * RA * 32 * 2;
        const int tceil32 = ((31 * p.hop_len + p.frame_len + (RA * NZ - p.frame_len) + 7) / 8 * 8 + 8 + 16 + 3) / 4 * 4;
        const int scratch = st->args.ef + 32;
        st->args.raw32_off = scratch + tceil32 <= ws_floats ? ws_floats - tceil32 : 0;
    }
    st->args.inv_n = static_cast<float>(1.0 / N);
    st->args.mel_magic = (1 << 20) / st->args.od + 1;
    st->args.preemph = p.preemph;
    st->args.log_floor = p.log_floor;
    const int HM = (M + 1) / 2;
    if (p.output == MFCC_OUT_CEPSTRA && HM <= kFoldMax)
        for (int k = 0; k < p.n_cep && k < KC; ++k)
            for (int q = 0; q < HM; ++q) {
                double d = std::log(2.0) * static_cast<double>(h.dct[static_cast<size_t>(k) * M + q]);
                if ((M & 1) && q == HM - 1) d *= 0.5;   // the middle band of an odd bank is folded onto itself
                (&st->args.dctc[k / 2][k & 1][q / 4].x)[q % 4] = static_cast<float>(d);
            }
    st->useg_ok = true;
    for (int w = 0; w < kWarps; ++w) {
        const int n = static_cast<int>(lists[w].size());
        if (n > kSegParam - 1) st->useg_ok = false;
        for (int q = 0; q < kSegParam; ++q) {
            float4 d = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            int x = 0, y = 0, z = -1;
            if (q < n && q < kSegParam - 1) {
                const Seg &g = lists[w][q];
                x = g.k0 * 128;
                y = g.w;
                z = g.row * 128;
                d.z = g.s;
            }
            std::memcpy(&d.x, &x, 4);
            std::memcpy(&d.y, &y, 4);
            std::memcpy(&d.w, &z, 4);
            st->args.useg[w][q] = d;
        }
    }
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
#endif

template <typename PcmT, int L, int HOP, int RB, int RA, int MEL, int CEP>
static int launch_variant(const SpState *st, const Tile *d_tiles, int64_t n_tiles, const PcmT *d_pcm, float *d_out,
                          int alaw, cudaStream_t stream)
{
    auto kern = fused_sp_kernel<PcmT, L, HOP, RB, RA, MEL, CEP>;
    static std::atomic<uint64_t> optin{0};
    if (ensure_smem_optin(kern, st->device, kSmemMax, optin) != MFCC_OK) return MFCC_ECUDA;
    SpArgs a = st->args;
    a.tiles = d_tiles;
    a.n_tiles = n_tiles;
    a.out = d_out;
    a.tab = st->d_tab;
    a.alaw = alaw;
    constexpr int GROUPS = groups_for(RB);
    const int64_t grid = std::min<int64_t>((n_tiles + GROUPS - 1) / GROUPS, st->sm_count);
    kern<<<static_cast<unsigned>(grid), GROUPS * kHalfThreads, st->smem, stream>>>(d_pcm, a);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError() == cudaSuccess ? MFCC_OK : MFCC_ECUDA;
}

// Tail-warp variants (band assembly + log + DCT / log-mel rows on two tail warps, everything in registers) exist for the
// filter / cepstrum counts front ends actually use; everything else takes the generic tail (MEL = 0).  CEP = 0: log-mel.
#define MFCC_SP_SHAPES_512(X) X(26, 13) X(40, 13) X(23, 13) X(26, 0) X(40, 0) X(80, 0)
#define MFCC_SP_SHAPES_256(X) X(20, 13) X(23, 13) X(20, 0) X(40, 0)

template <typename PcmT>
int sp_launch(const mfcc_plan *plan, const Tile *d_tiles, int64_t n_tiles, const PcmT *d_pcm, int64_t /*pcm_len*/,
              float *d_out, int alaw, cudaStream_t stream)
{
    const SpState *st = static_cast<const SpState *>(plan->sp_state);
    if (st == nullptr) return MFCC_ENOTSUP;
    const int mel = st->useg_ok ? st->args.n_mel : -1, cep = st->args.logmel ? 0 : st->args.n_cep;
    if (st->v->L == 400) {
#define X(M_, C_) if (mel == M_ && cep == C_) return launch_variant<PcmT, 400, 160, 32, 16, M_, C_>(st, d_tiles, n_tiles, d_pcm, d_out, alaw, stream);
        MFCC_SP_SHAPES_512(X)
#undef X
        return launch_variant<PcmT, 400, 160, 32, 16, 0, 0>(st, d_tiles, n_tiles, d_pcm, d_out, alaw, stream);
    }
#define X(M_, C_) if (mel == M_ && cep == C_) return launch_variant<PcmT, 200, 80, 16, 16, M_, C_>(st, d_tiles, n_tiles, d_pcm, d_out, alaw, stream);
    MFCC_SP_SHAPES_256(X)
#undef X
    return launch_variant<PcmT, 200, 80, 16, 16, 0, 0>(st, d_tiles, n_tiles, d_pcm, d_out, alaw, stream);
}

#if MFCC_SP_PCM_TYPES & 1
template int sp_launch<int16_t>(const mfcc_plan *, const Tile *, int64_t, const int16_t *, int64_t, float *, int, cudaStream_t);
#endif
#if MFCC_SP_PCM_TYPES & 2
template int sp_launch<float>(const mfcc_plan *, const Tile *, int64_t, const float *, int64_t, float *, int, cudaStream_t);
#endif
#if MFCC_SP_PCM_TYPES & 4
template int sp_launch<uint8_t>(const mfcc_plan *, const Tile *, int64_t, const uint8_t *, int64_t, float *, int, cudaStream_t);
#endif

}  // namespace mfcc
