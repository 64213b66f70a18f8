// tc_probe.cu — measured prototype of "pass 1 of the 512-point frame transform on tcgen05" (VERDICT r1, next-round 1d).
//
// What it is.  512 = 8 x 64: sample m = a + 8 b of a frame (a < 8 "column", b < 50).  Pass 1 is, per column a,
//     Y_a[k1] = sum_b w[a + 8b] y[a + 8b] W_512^((a + 8b) k1),   k1 = 0 .. 32
// (a windowed real DFT-64 over b with the inter-pass twiddle folded in): a [frames x 50] x [50 x 64] real GEMM per
// column — the part of the present CUDA-core kernel that costs 34.6 % of its time.  Here it runs on the tensor cores:
//   * frames are rows (M = 128 frames of ONE utterance = a "supertile"), accumulators live in TMEM: column a owns TMEM
//     columns [64 a, 64 a + 64) -> the 128 x 512 fp32 accumulator tile is exactly the SM's tensor memory;
//   * the A operand is NOT materialised per frame (frames overlap 2.5 x): the pre-emphasised samples y are staged ONCE
//     per sample as rows of hop blocks (160 samples, row j = y[160 j .. 160 j + 159]), regrouped by column with
//     ldmatrix.trans / stmatrix into 16-byte K chunks [chunk (a, c)][row j][8 b'].  Rows are 16 bytes apart (no-swizzle
//     K-major canonical layout with SBO = 128), so frame f's three hop blocks are rows f, f + 1, f + 2: block s of every
//     frame is the SAME buffer addressed 16 s bytes further on.  K = 64 per column in four K = 16 steps:
//         step s = 0, 1, 2 : chunks (a, c = 0), (a, c = 1) of row f + s      -> b = 20 s + 0 .. 15
//         step 3           : chunk (a, c = 2) of rows f and f + 1 (LBO = 16 B) -> b = 16 .. 19 and 36 .. 39
//   * fp32-grade accuracy from fp16 operands: y = yh + yl (yh = y truncated to 11 significant bits: exact in fp16;
//     yl = y - yh), B = Bh + Bl (the table split the same way, scaled by 2^10 so that Bl stays normal):
//     D = yh Bh + yl Bh + yh Bl, three products, 12 MMAs (M128 N64 K16) per column, 96 per supertile;
//   * B (8 columns x 2 pieces x 8 KB = 128 KB) does not fit next to A (100 KB): it streams from L2 through a ring of
//     8 KB slots filled by cp.async.bulk, full / empty mbarriers, tcgen05.commit releasing the slots.
// The probe runs the front half (staging, MMAs, TMEM read-back) on real descriptors, checks the accumulators against a
// double-precision host evaluation (and the host finishes pass 2 in double to check the whole factorisation against a
// direct DFT), and times the phases.  Prints JSON lines.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tools/tc_probe tools/tc_probe.cu
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace tc {

constexpr int L = 400, HOP = 160, NFFT = 512, NCOL = 8, NOUT = 64;
constexpr int FR = 128;                      // frames per supertile
constexpr int UNIT_ROWS = 3, UNIT_GROUPS = UNIT_ROWS * HOP / 8;   // 60 groups of 8 samples
constexpr int ROWS = 132;                    // hop-block rows staged per supertile (130 needed)
constexpr int UNITS = ROWS / UNIT_ROWS;
constexpr int CH = 24;                       // K chunks per row: 8 columns x 3
constexpr int LBO0 = 2128;                   // bytes between chunk planes: >= ROWS * 16, (3 LBO0 / 4) mod 32 = 4 * odd
constexpr int A_PIECE = CH * LBO0;
constexpr int B_SLOT = 8192;                 // one (column, piece): [8 k-chunks][64 n][8 k] fp16
#ifndef TC_RING
#define TC_RING 6
#endif
constexpr int RING = TC_RING;                // slots; RING / 2 columns in flight
constexpr int RAW_BYTES = (16 + ROWS * HOP * 2 + 15) / 16 * 16;
constexpr int STG_WARP = UNIT_GROUPS * 16;   // one piece of one unit
constexpr int kComputeWarps = 8, kComputeThreads = kComputeWarps * 32, kThreads = kComputeThreads + 32;
constexpr int OFF_A = 0, OFF_B = OFF_A + 2 * A_PIECE, OFF_RAW = OFF_B + RING * B_SLOT, OFF_STG = OFF_RAW + RAW_BYTES;
constexpr int OFF_BAR = OFF_STG + kComputeWarps * STG_WARP;
constexpr int SMEM_BYTES = OFF_BAR + 256;
static_assert(ROWS * 16 <= LBO0 && (3 * LBO0 / 4) % 8 == 4, "chunk plane stride");
static_assert(OFF_B % 16 == 0 && OFF_RAW % 16 == 0 && OFF_STG % 16 == 0 && OFF_BAR % 16 == 0, "alignment");

constexpr float kBScale = 1024.0f;
constexpr uint32_t kIdesc = (1u << 4) | (uint32_t(NOUT >> 3) << 17) | (uint32_t(FR >> 4) << 24);   // f16 x f16 -> f32, K-major A and B

// ---- PTX helpers ----
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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
__device__ __forceinline__ void compute_sync()
{
    asm volatile("bar.sync 1, %0;" ::"n"(kComputeThreads) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(kIdesc), "r"(accumulate) : "memory");
}
// shared-memory matrix descriptor, no swizzle, version 1 (sm_100): start >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo)
{
    return static_cast<uint64_t>((addr & 0x3FFFFu) >> 4) | (static_cast<uint64_t>(lbo >> 4) << 16) |
           (static_cast<uint64_t>(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3)
{
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr) : "memory");
}
__device__ __forceinline__ void ldsm_x1_t(uint32_t addr, uint32_t &r0)
{
    asm volatile("ldmatrix.sync.aligned.m8n8.x1.trans.shared.b16 {%0}, [%1];" : "=r"(r0) : "r"(addr) : "memory");
}
__device__ __forceinline__ void stsm_x4(uint32_t addr, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3)
{
    asm volatile("stmatrix.sync.aligned.m8n8.x4.shared.b16 [%0], {%1, %2, %3, %4};"
                 ::"r"(addr), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}
__device__ __forceinline__ void stsm_x1(uint32_t addr, uint32_t r0)
{
    asm volatile("stmatrix.sync.aligned.m8n8.x1.shared.b16 [%0], {%1};" ::"r"(addr), "r"(r0) : "memory");
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi)
{
    uint32_t d;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}
__device__ __forceinline__ float2 s16x2_to_f32(uint32_t w)
{
    const uint32_t b = w ^ 0x80008000u;
    const uint32_t lo = __byte_perm(b, 0x4B000000u, 0x7610);
    const uint32_t hi = __byte_perm(b, 0x4B000000u, 0x7632);
    return make_float2(__uint_as_float(lo) - 8421376.0f, __uint_as_float(hi) - 8421376.0f);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&r)[4])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

struct Args {
    const int16_t *pcm;     // supertile st starts at sample 8 + st * FR * HOP (8 lead samples exist)
    const uint8_t *bmat;    // [8 columns][2 pieces][B_SLOT] fp16, shared-memory image
    float *dump;            // [n_super * FR][512] accumulators (mode 0) or nullptr
    float *sink;            // one float per thread (mode 1)
    long long *clocks;      // per CTA: {stage, mma wait, read} cycles summed over its supertiles (thread 0)
    int n_super;
    float preemph;
    int flags;              // bit 0: skip staging arithmetic, bit 1: skip MMAs, bit 2: skip the TMEM read, bit 3: row-wise x4 reads
};

// One unit = 3 hop-block rows (60 groups of 8 samples): convert + pre-emphasise + split, stage one piece in natural
// order, regroup by column with ldmatrix.trans -> stmatrix into the chunk planes.
__device__ __forceinline__ void stage_unit(const int16_t *raw16, uint32_t a_base, uint32_t stg, int j0, int lane, float na,
                                           const uint32_t (&src_off)[3], const uint32_t (&dst_off)[3])
{
    uint4 hh[2], ll[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int g = lane + 32 * u;
        if (g < UNIT_GROUPS) {
            const int G = 20 * j0 + g;
            const uint4 q = *reinterpret_cast<const uint4 *>(raw16 + 8 + 8 * G);
            const float xp = s16x2_to_f32(static_cast<uint16_t>(raw16[7 + 8 * G])).x;
            const float2 x01 = s16x2_to_f32(q.x), x23 = s16x2_to_f32(q.y), x45 = s16x2_to_f32(q.z), x67 = s16x2_to_f32(q.w);
            float y[8];
            y[0] = fmaf(na, xp, x01.x);    y[1] = fmaf(na, x01.x, x01.y);
            y[2] = fmaf(na, x01.y, x23.x); y[3] = fmaf(na, x23.x, x23.y);
            y[4] = fmaf(na, x23.y, x45.x); y[5] = fmaf(na, x45.x, x45.y);
            y[6] = fmaf(na, x45.y, x67.x); y[7] = fmaf(na, x67.x, x67.y);
            float h[8], l[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                h[i] = __uint_as_float(__float_as_uint(y[i]) & 0xFFFFE000u);   // 11 significant bits: exact in fp16
                l[i] = y[i] - h[i];
            }
            hh[u] = make_uint4(pack_h2(h[0], h[1]), pack_h2(h[2], h[3]), pack_h2(h[4], h[5]), pack_h2(h[6], h[7]));
            ll[u] = make_uint4(pack_h2(l[0], l[1]), pack_h2(l[2], l[3]), pack_h2(l[4], l[5]), pack_h2(l[6], l[7]));
        }
    }
#pragma unroll
    for (int piece = 0; piece < 2; ++piece) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int g = lane + 32 * u;
            if (g < UNIT_GROUPS) {
                const uint4 v = piece ? ll[u] : hh[u];
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg + g * 16), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
            }
        }
        __syncwarp();
        const uint32_t dst = a_base + piece * A_PIECE + j0 * 16;
        uint32_t r0, r1, r2, r3;
        ldsm_x4_t(stg + src_off[0], r0, r1, r2, r3);
        stsm_x4(dst + dst_off[0], r0, r1, r2, r3);
        ldsm_x4_t(stg + src_off[1], r0, r1, r2, r3);
        stsm_x4(dst + dst_off[1], r0, r1, r2, r3);
        ldsm_x1_t(stg + src_off[2], r0);
        stsm_x1(dst + dst_off[2], r0);
        __syncwarp();
    }
}

__global__ void __launch_bounds__(kThreads, 1) tc_front_kernel(const Args a)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t s_base = smem_u32(smem);
    const uint32_t s_a = s_base + OFF_A, s_b = s_base + OFF_B, s_raw = s_base + OFF_RAW;
    const uint32_t bar = s_base + OFF_BAR;
    const uint32_t bar_raw = bar, bar_aready = bar + 8, bar_mma = bar + 16, bar_full = bar + 24, bar_empty = bar + 24 + 8 * RING;
    volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(smem + OFF_BAR + 24 + 16 * RING);
    const int16_t *raw16 = reinterpret_cast<const int16_t *>(smem + OFF_RAW);

    const int first = blockIdx.x, step = gridDim.x;
    const int my_super = first < a.n_super ? (a.n_super - first + step - 1) / step : 0;

    if (tid == 0) {
        mbar_init(bar_raw, 1);
        mbar_init(bar_aready, 1);
        mbar_init(bar_mma, 1);
        for (int s = 0; s < RING; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kComputeWarps) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(const_cast<uint32_t *>(tmem_slot))), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    auto issue_raw = [&](int st) {   // 8 lead samples + ROWS hop blocks
        const int16_t *src = a.pcm + 8 + static_cast<size_t>(st) * FR * HOP - 8;
        const uint32_t bytes = 16 + ROWS * HOP * 2;
        mbar_expect_tx(bar_raw, bytes);
        bulk_g2s(s_raw, src, bytes, bar_raw);
    };

    if (warp < kComputeWarps) {
        // ---------------- compute warps: staging, then the accumulators ----------------
        if (tid == 0 && my_super > 0) issue_raw(first);
        const uint32_t stg = s_base + OFF_STG + warp * STG_WARP;
        // per-lane source rows (staged group) and destination rows (chunk plane) of the 9 transposes of a unit
        uint32_t src_off[3], dst_off[3];
#pragma unroll
        for (int bt = 0; bt < 3; ++bt) {
            const int m = bt * 4 + (bt < 2 ? lane / 8 : 0), i = lane & 7;
            const int r = m / 3, c = m % 3;
            int g = 20 * r + 8 * c + i;
            if (g >= UNIT_GROUPS) g = UNIT_GROUPS - 1;
            src_off[bt] = g * 16;
            dst_off[bt] = (i * 3 + c) * LBO0 + r * 16;     // row a = lane % 8 of the transposed matrix
        }
        const float na = -a.preemph;
        long long t_stage = 0, t_wait = 0, t_read = 0;
        float acc = 0.0f;
        for (int it = 0; it < my_super; ++it) {
            const int st = first + it * step;
            long long c0 = clock64();
            mbar_wait(bar_raw, it & 1);
            if (!(a.flags & 1)) {
                for (int u = warp; u < UNITS; u += kComputeWarps) stage_unit(raw16, s_a, stg, u * UNIT_ROWS, lane, na, src_off, dst_off);
            }
            fence_async_smem();     // the chunk planes were written through the generic proxy, the MMAs read them through the async proxy
            tc_fence_before();      // this thread's TMEM reads of the previous supertile are complete (tmem_wait_ld below)
            compute_sync();
            if (tid == 0) {
                if (it + 1 < my_super) issue_raw(st + step);
                mbar_arrive(bar_aready);
            }
            long long c1 = clock64();
            mbar_wait(bar_mma, it & 1);
            tc_fence_after();
            long long c2 = clock64();
            if (!(a.flags & 4)) {
                const int q = warp & 3, half = warp >> 2;
                const uint32_t tq = tmem + (static_cast<uint32_t>(32 * q) << 16);
                if (a.dump != nullptr) {
                    float *drow = a.dump + (static_cast<size_t>(st) * FR + 32 * q + lane) * 512 + 256 * half;
#pragma unroll 1
                    for (int cb = 0; cb < 8; ++cb) {
                        uint32_t r[32];
                        tmem_ld32(tq + 256 * half + 32 * cb, r);
                        tmem_wait_ld();
#pragma unroll
                        for (int i = 0; i < 32; i += 4)
                            *reinterpret_cast<float4 *>(drow + 32 * cb + i) = make_float4(__uint_as_float(r[i]), __uint_as_float(r[i + 1]),
                                                                                          __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
                    }
                } else if (a.flags & 8) {
                    // the access pattern of pass 2: two rows k1, k1 + 1 -> 8 loads of 4 columns, one per column a
#pragma unroll 1
                    for (int rr = 0; rr < 8; ++rr) {
                        const int k1 = 16 * half + 2 * rr;
                        uint32_t r[8][4];
#pragma unroll
                        for (int c = 0; c < 8; ++c) tmem_ld4(tq + 64 * c + 2 * k1, r[c]);
                        tmem_wait_ld();
#pragma unroll
                        for (int c = 0; c < 8; ++c) acc += __uint_as_float(r[c][0]) * __uint_as_float(r[c][1]) + __uint_as_float(r[c][2]) * __uint_as_float(r[c][3]);
                    }
                } else {
#pragma unroll 1
                    for (int cb = 0; cb < 8; ++cb) {
                        uint32_t r[32];
                        tmem_ld32(tq + 256 * half + 32 * cb, r);
                        tmem_wait_ld();
#pragma unroll
                        for (int i = 0; i < 32; ++i) acc += __uint_as_float(r[i]);
                    }
                }
            }
            long long c3 = clock64();
            t_stage += c1 - c0;
            t_wait += c2 - c1;
            t_read += c3 - c2;
        }
        if (a.sink != nullptr) a.sink[blockIdx.x * kComputeThreads + tid] = acc;
        if (tid == 0 && a.clocks != nullptr) {
            a.clocks[3 * blockIdx.x + 0] = t_stage;
            a.clocks[3 * blockIdx.x + 1] = t_wait;
            a.clocks[3 * blockIdx.x + 2] = t_read;
        }
        tc_fence_before();
    } else {
        // ---------------- control warp: B ring + MMA issue (one elected lane issues; all lanes wait) ----------------
        uint32_t leader;
        asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.b32 %0, 1, 0, P;\n}\n" : "=r"(leader));
        const int total_cols = my_super * NCOL;
        auto load_col = [&](int cc) {
            const int col = cc % NCOL;
#pragma unroll
            for (int piece = 0; piece < 2; ++piece) {
                const int slot = (2 * cc + piece) % RING;
                mbar_expect_tx(bar_full + 8 * slot, B_SLOT);
                bulk_g2s(s_b + slot * B_SLOT, a.bmat + static_cast<size_t>(2 * col + piece) * B_SLOT, B_SLOT, bar_full + 8 * slot);
            }
        };
        if (leader)
            for (int cc = 0; cc < RING / 2 && cc < total_cols; ++cc) load_col(cc);
        // descriptors: constant high words, low words = start address >> 4 | LBO >> 4 << 16
        const uint32_t a_hi_word = (128u >> 4) | (1u << 14), b_hi_word = a_hi_word;
        const uint32_t a_lo_ks012 = ((s_a & 0x3FFFFu) >> 4) | (static_cast<uint32_t>(LBO0 >> 4) << 16);
        const uint32_t a_lo_ks3 = (((s_a + 2 * LBO0) & 0x3FFFFu) >> 4) | (1u << 16);
        const uint32_t b_lo0 = ((s_b & 0x3FFFFu) >> 4) | (static_cast<uint32_t>(1024 >> 4) << 16);
        const bool resident = (a.flags & 16) != 0;     // timing only: reuse whatever the slots hold, no refills
        for (int it = 0; it < my_super; ++it) {
            mbar_wait(bar_aready, it & 1);
            tc_fence_after();
#pragma unroll 1
            for (int col = 0; col < NCOL; ++col) {
                const int cc = it * NCOL + col;
                const int slot_h = (2 * cc) % RING, slot_l = (2 * cc + 1) % RING;
                const uint32_t use = static_cast<uint32_t>((2 * cc) / RING);
                if (!resident || use == 0) {
                    mbar_wait(bar_full + 8 * slot_h, use & 1);
                    mbar_wait(bar_full + 8 * slot_l, use & 1);
                }
                tc_fence_after();
                if (leader) {
                    if (!(a.flags & 2)) {
                        const uint32_t acol = col * (3 * LBO0 >> 4);
#pragma unroll
                        for (int prod = 0; prod < 3; ++prod) {
                            const uint32_t ap = acol + (prod == 1 ? (A_PIECE >> 4) : 0);
                            const uint32_t bp = b_lo0 + ((prod == 2 ? slot_l : slot_h) * (B_SLOT >> 4));
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) {
                                const uint32_t alo = ks < 3 ? a_lo_ks012 + ap + ks : a_lo_ks3 + ap;
                                const uint64_t ad = (static_cast<uint64_t>(a_hi_word) << 32) | alo;
                                const uint64_t bd = (static_cast<uint64_t>(b_hi_word) << 32) | (bp + ks * (2048 >> 4));
                                tc_mma(tmem + 64 * col, ad, bd, (prod | ks) != 0);
                            }
                        }
                    }
                    tc_commit(bar_empty + 8 * slot_h);
                    tc_commit(bar_empty + 8 * slot_l);
                }
                // the slots of column cc - 1 take column cc - 1 + RING / 2 once its MMAs have completed
                if (!resident && cc >= 1 && cc - 1 + RING / 2 < total_cols) {
                    const int pc = cc - 1;
                    const uint32_t puse = static_cast<uint32_t>((2 * pc) / RING);
                    mbar_wait(bar_empty + 8 * ((2 * pc) % RING), puse & 1);
                    mbar_wait(bar_empty + 8 * ((2 * pc + 1) % RING), puse & 1);
                    if (leader) load_col(pc + RING / 2);
                }
                __syncwarp();
            }
            if (leader) tc_commit(bar_mma);
            __syncwarp();
        }
    }
    __syncthreads();
    if (warp == kComputeWarps) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

// ---------------------------------------------------------------- host ----------------------------------------------------------------
static uint16_t f32_to_f16_bits(float f)   // round to nearest even, handles subnormals (no NaN / Inf inputs here)
{
    uint32_t x;
    std::memcpy(&x, &f, 4);
    const uint32_t sign = (x >> 16) & 0x8000u;
    int32_t e = static_cast<int32_t>((x >> 23) & 0xFF) - 127 + 15;
    uint32_t m = x & 0x7FFFFFu;
    if (((x >> 23) & 0xFF) == 0) return static_cast<uint16_t>(sign);
    if (e >= 31) return static_cast<uint16_t>(sign | 0x7BFF);
    if (e <= 0) {
        if (e < -10) return static_cast<uint16_t>(sign);
        m |= 0x800000u;
        const int shift = 14 - e;
        uint32_t v = m >> shift;
        const uint32_t rem = m & ((1u << shift) - 1), half = 1u << (shift - 1);
        if (rem > half || (rem == half && (v & 1))) ++v;
        return static_cast<uint16_t>(sign | v);
    }
    uint32_t v = (static_cast<uint32_t>(e) << 10) | (m >> 13);
    const uint32_t rem = m & 0x1FFFu;
    if (rem > 0x1000u || (rem == 0x1000u && (v & 1))) ++v;
    return static_cast<uint16_t>(sign | v);
}
static float f16_bits_to_f32(uint16_t h)
{
    const uint32_t sign = (h & 0x8000u) << 16;
    const int e = (h >> 10) & 0x1F;
    const uint32_t m = h & 0x3FFu;
    float v;
    if (e == 0) v = std::ldexp(static_cast<float>(m), -24);
    else v = std::ldexp(static_cast<float>(m | 0x400u), e - 25);
    uint32_t x;
    std::memcpy(&x, &v, 4);
    x |= sign;
    std::memcpy(&v, &x, 4);
    return v;
}

// K index kappa of column a -> b (row of the column), or -1 (zero row)
static int kappa_to_b(int kappa)
{
    const int ks = kappa / 16, kk = kappa % 16;
    if (ks < 3) {
        const int b = 20 * ks + kk;
        return b < 50 ? b : -1;
    }
    if (kk < 8) return kk < 4 ? 16 + kk : -1;
    return kk - 8 < 4 ? 36 + (kk - 8) : -1;
}
// value of the pass-1 matrix (window x DFT-64 x inter-pass twiddle), output slot n of column a, frame sample m = a + 8 b
static double bvalue(const std::vector<double> &win, int a, int b, int n)
{
    const int m = a + 8 * b;
    if (m >= L) return 0.0;
    const double w = win[m];
    if (n == 0) return w;
    if (n == 1) return (b & 1) ? -w : w;              // k1 = 32 without the inter-pass twiddle: W_64^(32 b) = (-1)^b
    const int k1 = n / 2;
    const double ang = -2.0 * M_PI * static_cast<double>(m) * k1 / NFFT;
    return (n & 1) ? w * std::sin(ang) : w * std::cos(ang);
}

}  // namespace tc

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::printf("{\"error\": \"%s at %s:%d\"}\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

int main(int argc, char **argv)
{
    using namespace tc;
    const int reps = argc > 1 ? std::atoi(argv[1]) : 3;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    std::printf("{\"probe\": \"tc_front\", \"device\": \"%s\", \"sms\": %d, \"smem_bytes\": %d, \"ring_slots\": %d}\n", prop.name, sms, SMEM_BYTES, RING);

    // window and B image
    std::vector<double> win(L);
    for (int i = 0; i < L; ++i) win[i] = 0.54 - 0.46 * std::cos(2.0 * M_PI * i / (L - 1));
    std::vector<uint8_t> bimg(static_cast<size_t>(NCOL) * 2 * B_SLOT, 0);
    std::vector<float> bsum(static_cast<size_t>(NCOL) * 64 * NOUT, 0.0f);   // Bh + Bl as the kernel sees it, [a][kappa][n]
    for (int a = 0; a < NCOL; ++a)
        for (int kappa = 0; kappa < 64; ++kappa) {
            const int b = kappa_to_b(kappa);
            for (int n = 0; n < NOUT; ++n) {
                const double v = b < 0 ? 0.0 : bvalue(win, a, b, n) * kBScale;
                const uint16_t hb = f32_to_f16_bits(static_cast<float>(v));
                const float hf = f16_bits_to_f32(hb);
                const uint16_t lb = f32_to_f16_bits(static_cast<float>(v - hf));
                const size_t off = static_cast<size_t>(kappa / 8) * 1024 + (n / 8) * 128 + (n % 8) * 16 + (kappa % 8) * 2;
                std::memcpy(&bimg[(static_cast<size_t>(2 * a) * B_SLOT) + off], &hb, 2);
                std::memcpy(&bimg[(static_cast<size_t>(2 * a + 1) * B_SLOT) + off], &lb, 2);
                bsum[(static_cast<size_t>(a) * 64 + kappa) * NOUT + n] = hf + f16_bits_to_f32(lb);
            }
        }

    const int n_check = 2 * sms + 3;          // supertiles of the checked run (more than one per CTA)
    const int n_time = sms * 48;
    const size_t n_samples = 8 + static_cast<size_t>(n_time) * FR * HOP + ROWS * HOP + 64;
    std::vector<int16_t> pcm(n_samples);
    uint64_t s = 0x9E3779B97F4A7C15ull;
    for (size_t i = 0; i < n_samples; ++i) {   // sum of uniforms ~ Gaussian sigma 3000, plus a loud tone in the first supertiles
        double g = 0.0;
        for (int k = 0; k < 6; ++k) {
            s = s * 6364136223846793005ull + 1442695040888963407ull;
            g += static_cast<double>((s >> 33) & 0xFFFFFF) / 16777216.0 - 0.5;
        }
        double v = g * 3000.0 * std::sqrt(2.0);
        if (i < 200000) v = v * 0.01 + 20000.0 * std::sin(2.0 * M_PI * 1000.0 * i / 16000.0);
        pcm[i] = static_cast<int16_t>(std::lrint(std::fmax(-32768.0, std::fmin(32767.0, v))));
    }
    int16_t *d_pcm = nullptr;
    uint8_t *d_b = nullptr;
    float *d_dump = nullptr, *d_sink = nullptr;
    long long *d_clk = nullptr;
    CK(cudaMalloc(&d_pcm, n_samples * 2));
    CK(cudaMalloc(&d_b, bimg.size()));
    CK(cudaMalloc(&d_dump, static_cast<size_t>(n_check) * FR * 512 * 4));
    CK(cudaMalloc(&d_sink, static_cast<size_t>(sms) * kComputeThreads * 4));
    CK(cudaMalloc(&d_clk, static_cast<size_t>(sms) * 3 * 8));
    CK(cudaMemcpy(d_pcm, pcm.data(), n_samples * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_b, bimg.data(), bimg.size(), cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(tc_front_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));

    // ---- 1. correctness ----
    Args a{};
    a.pcm = d_pcm;
    a.bmat = d_b;
    a.dump = d_dump;
    a.sink = nullptr;
    a.clocks = nullptr;
    a.n_super = n_check;
    a.preemph = 0.97f;
    a.flags = 0;
    CK(cudaMemset(d_dump, 0xFF, static_cast<size_t>(n_check) * FR * 512 * 4));
    tc_front_kernel<<<sms, kThreads, SMEM_BYTES>>>(a);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<float> dump(static_cast<size_t>(n_check) * FR * 512);
    CK(cudaMemcpy(dump.data(), d_dump, dump.size() * 4, cudaMemcpyDeviceToHost));
    {
        double max_rel_split = 0.0, max_rel_true = 0.0, scale_max = 0.0;
        long long bad = 0, nonfinite = 0;
        double worst_bin_rel = 0.0;
        const int supers[] = {0, 1, sms - 1, sms, 2 * sms + 2};
        for (int st : supers) {
            for (int f = 0; f < FR; f += 7) {
                // y of this frame in float, exactly as the kernel forms it
                const size_t s0 = 8 + static_cast<size_t>(st) * FR * HOP + static_cast<size_t>(f) * HOP;
                std::vector<float> y(L);
                for (int m = 0; m < L; ++m) y[m] = std::fmaf(-0.97f, static_cast<float>(pcm[s0 + m - 1]), static_cast<float>(pcm[s0 + m]));
                const float *row = &dump[(static_cast<size_t>(st) * FR + f) * 512];
                std::vector<double> ytrue(512);
                double fmax = 0.0;
                for (int col = 0; col < NCOL; ++col)
                    for (int n = 0; n < NOUT; ++n) {
                        double split = 0.0, tr = 0.0;
                        for (int kappa = 0; kappa < 64; ++kappa) {
                            const int b = kappa_to_b(kappa);
                            if (b < 0) continue;
                            const int m = col + 8 * b;
                            if (m >= L) continue;
                            split += static_cast<double>(y[m]) * bsum[(static_cast<size_t>(col) * 64 + kappa) * NOUT + n];
                            tr += static_cast<double>(y[m]) * bvalue(win, col, b, n) * kBScale;
                        }
                        ytrue[col * 64 + n] = tr;
                        fmax = std::fmax(fmax, std::fabs(tr));
                        const double got = row[col * 64 + n];
                        if (!std::isfinite(got)) { ++nonfinite; continue; }
                        (void)split;
                    }
                scale_max = std::fmax(scale_max, fmax);
                for (int i = 0; i < 512; ++i) {
                    const double got = row[i];
                    if (!std::isfinite(got)) continue;
                    const double e = std::fabs(got - ytrue[i]) / fmax;
                    max_rel_true = std::fmax(max_rel_true, e);
                    if (e > 1e-5) ++bad;
                }
                // finish pass 2 on the host in double from the GPU accumulators and compare the power spectrum with a direct DFT
                if (f % 21 == 0) {
                    std::vector<double> pw(257, 0.0), pd(257, 0.0);
                    for (int k = 0; k <= 256; ++k) {
                        double re = 0.0, im = 0.0;
                        for (int m = 0; m < L; ++m) {
                            const double ang = -2.0 * M_PI * static_cast<double>(m) * k / NFFT;
                            re += win[m] * y[m] * std::cos(ang);
                            im += win[m] * y[m] * std::sin(ang);
                        }
                        pd[k] = re * re + im * im;
                    }
                    const double inv = 1.0 / kBScale;
                    for (int k1 = 0; k1 <= 32; ++k1)
                        for (int k2 = 0; k2 < 8; ++k2) {
                            const int k = k1 + 64 * k2;
                            double re = 0.0, im = 0.0;
                            for (int col = 0; col < NCOL; ++col) {
                                double yr, yi;
                                if (k1 == 0) { yr = row[col * 64 + 0] * inv; yi = 0.0; }
                                else if (k1 == 32) {
                                    const double t = -2.0 * M_PI * 32.0 * col / NFFT;
                                    yr = row[col * 64 + 1] * inv * std::cos(t);
                                    yi = row[col * 64 + 1] * inv * std::sin(t);
                                } else { yr = row[col * 64 + 2 * k1] * inv; yi = row[col * 64 + 2 * k1 + 1] * inv; }
                                const double t = -2.0 * M_PI * col * k2 / 8.0;
                                re += yr * std::cos(t) - yi * std::sin(t);
                                im += yr * std::sin(t) + yi * std::cos(t);
                            }
                            const int kk = k <= 256 ? k : 512 - k;
                            if (k <= 256 || k1 != 0) pw[kk] = re * re + im * im;
                        }
                    double pmax = 0.0;
                    for (int k = 0; k <= 256; ++k) pmax = std::fmax(pmax, pd[k]);
                    for (int k = 0; k <= 256; ++k) {
                        // relative to the bin itself with a floor 100 dB under the peak (fp32 FFT noise floor is about -130 dB)
                        const double e = std::fabs(pw[k] - pd[k]) / std::fmax(pd[k], pmax * 1e-10);
                        worst_bin_rel = std::fmax(worst_bin_rel, e);
                    }
                }
                (void)max_rel_split;
            }
        }
        std::printf("{\"test\": \"accumulators\", \"max_err_rel_to_frame_max\": %.3e, \"elements_over_1e-5\": %lld, \"nonfinite\": %lld, "
                    "\"max_abs_value\": %.3e, \"power_spectrum_worst_rel_err_floor_-100dB\": %.3e}\n",
                    max_rel_true, bad, nonfinite, scale_max, worst_bin_rel);
    }

    // ---- 2. timing ----
    a.dump = nullptr;
    a.sink = d_sink;
    a.clocks = d_clk;
    a.n_super = n_time;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const struct { const char *name; int flags; } modes[] = {
        {"all (stage + mma + tmem read x32)", 0}, {"all, row-wise x4 reads (pass-2 pattern)", 8}, {"no staging arithmetic", 1},
        {"no mma", 2}, {"all, B resident (no refills; timing only)", 16}, {"mma only, B resident", 1 | 4 | 16}, {"skeleton, B resident", 1 | 2 | 4 | 16}, {"no tmem read", 4}, {"stage only", 2 | 4}, {"mma only", 1 | 4}, {"skeleton", 1 | 2 | 4},
    };
    for (const auto &md : modes) {
        a.flags = md.flags;
        tc_front_kernel<<<sms, kThreads, SMEM_BYTES>>>(a);   // warm
        CK(cudaDeviceSynchronize());
        float best = 1e30f;
        for (int r = 0; r < reps; ++r) {
            CK(cudaEventRecord(e0));
            tc_front_kernel<<<sms, kThreads, SMEM_BYTES>>>(a);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms = 0.0f;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            best = std::fmin(best, ms);
        }
        CK(cudaGetLastError());
        std::vector<long long> clk(static_cast<size_t>(sms) * 3);
        CK(cudaMemcpy(clk.data(), d_clk, clk.size() * 8, cudaMemcpyDeviceToHost));
        double cs = 0, cw = 0, cr = 0;
        for (int i = 0; i < sms; ++i) { cs += clk[3 * i]; cw += clk[3 * i + 1]; cr += clk[3 * i + 2]; }
        const double per = static_cast<double>(n_time) / sms;
        const double frames = static_cast<double>(n_time) * FR;
        std::printf("{\"test\": \"timing\", \"mode\": \"%s\", \"ms\": %.4f, \"frames_per_s\": %.4g, \"us_per_supertile_per_sm\": %.3f, "
                    "\"clk_per_supertile\": {\"stage\": %.0f, \"mma_wait\": %.0f, \"read\": %.0f}}\n",
                    md.name, best, frames / (best * 1e-3), best * 1e3 / per, cs / sms / per, cw / sms / per, cr / sms / per);
    }
    return 0;
}
