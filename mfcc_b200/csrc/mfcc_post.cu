// mfcc_post.cu — fused post-processing of the feature matrix: per-utterance cepstral mean /
// variance normalisation, delta and delta-delta regression, written ONCE as the stacked matrix
// [frames][dim * (1 + order)] = static | delta | delta-delta that acoustic models consume.
//
// Unlike the transform kernels this path is HBM-bound: per frame it reads dim floats (twice when CMVN needs the
// statistics first; the second read is served by the L2 when the matrix fits its 126 MB) and writes dim * (1 + order)
// floats, against ~40 FP32 operations.  So the design rules are the memory ones:
//   * every global access is a sweep over a CONTIGUOUS range (a chunk = up to `post_rows` consecutive rows of one
//     utterance is one contiguous block of the input and one contiguous block of the output), consecutive threads on
//     consecutive floats;
//   * the three output parts of a row are produced by the same sweep (thread = output column walking down the rows), so an
//     output sector is written once, completely, by neighbouring threads — not by three passes with a 3 * dim stride;
//   * the chunk (plus the +-2 * window halo rows the two regressions need, edge rows replicated as the spec says) is
//     staged in shared memory once, by ONE bulk async copy (TMA engine); the delta of the halo rows is recomputed instead
//     of exchanged;
//   * the apply grid is one CTA per chunk with 6 CTAs resident per SM (27 KB of shared memory each for 13 cepstra), the
//     statistics kernel is persistent with a ring of three bulk copies per CTA: enough bytes in flight to cover the HBM
//     latency; nothing is allocated per call.
// Statistics: per group of four chunks sums of (x - pivot) and (x - pivot)^2 (pivot = the group's first row, which keeps the
// uncentred variance formula free of cancellation), combined in group order by a second tiny kernel — deterministic, no
// floating-point atomics, no fences.  Three launches per call with CMVN (statistics, finalise, apply), one without.
//
// Spec (oracle/mfcc_oracle.c oracle_cmvn_f32 + oracle_delta_f32 applied twice): x' = (x - mu) [* 1 / sqrt(max(var,
// 1e-20))], d[t] = sum_n n (x'[t + n] - x'[t - n]) / (2 sum n^2) with frame indices clamped to the utterance,
// dd = the same regression of d.  No reference code corresponds to this (SURVEY.md §8a "Ref file:line = none").
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>

#include "mfcc_host.h"

namespace mfcc {

namespace {

// resident CTAs per SM the apply / statistics kernels are compiled for (register budget 65,536 / (256 x this)): with 8
// the apply kernel spills (32 registers) and runs at 78 % of the HBM roofline, with 6 (40 registers, no spills) at 88 %
#ifndef MFCC_POST_MINB
#define MFCC_POST_MINB 6
#endif
#ifndef MFCC_POST_MINB_STATS
#define MFCC_POST_MINB_STATS 8
#endif
constexpr int kPostThreads = 256;
constexpr int kPostMaxDim = 256;   // dim <= n_mel + 1 <= 129 for any valid plan
constexpr int kPostFront = 8;      // floats kept free in front of a staged block: alignment shift of the bulk copy (<= 7)

// Everything the kernels would otherwise derive with integer divisions by run-time values, prepared on the host: a chunk
// gives a thread only ~40 output elements, so a handful of 20-instruction divisions per thread is a third of the work.
struct PostGeom {
    int dim, rows, order, window, cmvn;
    int part0;            // first output part: 0 = static | delta | ..., 1 = the regressions only (mfcc_delta_batch)
    int per;              // (256 / dim) * dim: thread t < per always meets column t % dim
    int nsub;             // per / dim
    int s_per, s_nsub;    // the same for the thread count of the statistics kernel
    int od, cw, nblk;     // output columns, columns per pass of the stacking sweep, row blocks in it
    unsigned m_dim, m_cw, m_nblk, m_nsub;   // t / d == (t * m) >> 20 for t < 4096, d <= 256
    float inv_den;
};
__device__ __forceinline__ int div20(int t, unsigned m) { return static_cast<int>((static_cast<unsigned>(t) * m) >> 20); }
unsigned magic20(int d) { return ((1u << 20) + static_cast<unsigned>(d) - 1u) / static_cast<unsigned>(d); }

// ---- mbarrier + 1-D bulk async copy (TMA engine) ----
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

// Stage `cnt` floats starting at `src` into shared memory so that element i lands at block[i].  `floor` is a 16-byte aligned
// region; the block starts at least `min_off` floats into it (the apply kernel keeps its halo rows in front) and at most 3
// floats later: where exactly follows from the alignment of src.
// Fast path (matrix base 16-byte aligned): ONE bulk copy from the 16-byte boundary below src up to the last 16-byte boundary
// inside the range — no load instructions, no registers, the whole chunk in flight at once — and the <= 3 floats behind it by
// plain loads; the bytes in front of src belong to the previous row of the same matrix.  Nothing past the range is read.
// Returns the block pointer; every thread must call this (it contains the barrier that publishes the mbarrier) and must
// call stage_wait before reading the block.  min_off >= 4.
struct Staged { float *block; bool bulk; };
__device__ __forceinline__ Staged stage_issue(float *floor, int min_off, uint32_t bar, const float *src, int cnt, bool base_aligned, int tid)
{
    const int a = base_aligned ? static_cast<int>((reinterpret_cast<uintptr_t>(src) >> 2) & 3) : 0;   // floats between the boundary and src
    float *block = floor + min_off + ((a - min_off) & 3);      // block - a is 16-byte aligned
    const int body = base_aligned ? ((a + cnt) & ~3) : 0;      // floats the bulk copy brings (from src - a)
    if (tid == 0) {
        mbar_init(bar, 1);
        if (body > 0) {
            mbar_expect_tx(bar, static_cast<uint32_t>(body) * 4u);
            bulk_g2s(smem_u32(block - a), src - a, static_cast<uint32_t>(body) * 4u, bar);
        }
    }
    // what the bulk copy leaves out: everything without it, else the <= 3 floats behind the last 16-byte boundary
    for (int i = (body > 0 ? body - a : 0) + tid; i < cnt; i += kPostThreads) block[i] = __ldg(src + i);
    __syncthreads();                                // mbarrier initialised and visible to the waiters
    return Staged{block, body > 0};
}
__device__ __forceinline__ void stage_wait(const Staged &st, uint32_t bar)
{
    if (st.bulk) mbar_wait(bar, 0);
}

// ---- per-utterance statistics: persistent CTAs, each walks a contiguous range of chunks through a ring of bulk copies ----
// Work is cut into GROUPS of up to kStatGroup consecutive chunks of one utterance, aligned to the utterance (chunks first +
// 4 k .. first + 4 k + 3): a thread keeps its running sums in registers across the chunks of a group and the expensive part
// — the cross-thread reduction and the partial to global memory — happens once per group, while the bulk copies of the next
// chunks are already in flight (ring of kStatRing slots, one mbarrier each).  There is no fence, no counter and no "last
// CTA" logic in this kernel: the partials are combined by post_finalize_kernel, a separate (tiny) launch — a fence + atomic
// per group cost 21 us of the 71 us pass, the extra launch costs 8 (profiles/r2_post.md).  Which CTA walks which
// group depends on the launch; what is summed in which order depends on the group alone, so every bit of the result is
// independent of the grid and of the range of the batch a launch covers.
// Inside a thread the values it meets are summed in f32 about a pivot (the first row of the group: |x - pivot| is a few sigma,
// so the sums keep 7 digits of a quantity whose mean needs 5, and the uncentred variance formula is free of cancellation);
// across threads and groups everything is double, in a fixed order, re-referenced to the utterance's first row — no
// floating-point atomics.
// threads per CTA, ring depth and resident CTAs per SM of the statistics kernel (build-time switches for A/B timing; one
// box, whole call in ms: 256 / 3 / 4: 0.2142, 256 / 2 / 6: 0.2116, 128 / 2 / 6: 0.2115, 128 / 3 / 4: 0.2156 — not sensitive)
#ifndef MFCC_POST_STAT_THREADS
#define MFCC_POST_STAT_THREADS 256
#endif
#ifndef MFCC_POST_STAT_RING
#define MFCC_POST_STAT_RING 2
#endif
#ifndef MFCC_POST_STAT_CTAS
#define MFCC_POST_STAT_CTAS 6
#endif
constexpr int kStatGroup = 4, kStatRing = MFCC_POST_STAT_RING, kStatThreads = MFCC_POST_STAT_THREADS;
struct StatPartial { double s, q, pivot, n; };    // sums of (x - pivot), (x - pivot)^2 over n rows

__global__ void __launch_bounds__(kStatThreads, MFCC_POST_STAT_CTAS)
post_stats_kernel(const PostChunk *__restrict__ chunks, int chunk0, int n_chunks, const float *__restrict__ feat, const PostGeom g,
                  StatPartial *__restrict__ partial)
{
    extern __shared__ __align__(16) float sm[];     // [16: mbarriers][kStatRing slots of 4 + 4 + rows * dim (rounded up to 4) floats]
    __shared__ double s_s[kStatThreads], s_q[kStatThreads], s2_s[kStatThreads], s2_q[kStatThreads];
    const int tid = threadIdx.x, dim = g.dim, per = g.s_per, nsub = g.s_nsub;
    const int sub = div20(tid, g.m_dim), col = tid - sub * dim;
    const int slot_floats = 8 + ((g.rows * dim + 3) & ~3);
    const bool base_aligned = (reinterpret_cast<uintptr_t>(feat) & 15) == 0;
    const int end = chunk0 + n_chunks;
    // this CTA's range of chunks, both ends moved up to the next start of a group
    auto group_start = [&](int c) {
        while (c < end) {
            const int first = chunks[c].first_chunk;
            if ((c - first) % kStatGroup == 0) break;
            ++c;
        }
        return c;
    };
    const int lo = group_start(chunk0 + static_cast<int>(static_cast<int64_t>(n_chunks) * blockIdx.x / gridDim.x));
    const int hi = group_start(chunk0 + static_cast<int>(static_cast<int64_t>(n_chunks) * (blockIdx.x + 1) / gridDim.x));
    if (lo >= hi) return;

    // Where chunk k lands in its slot: a = floats between the 16-byte boundary below its first float and that float (the
    // matrix base is aligned, so this is (row0 * dim) mod 4: 32-bit arithmetic on the low bits), block - a is 16-byte aligned.
    // Readers need only a and the float count; the issuer (warp 0) also the source address.
    auto shift_of = [&](const PostChunk &ck) { return base_aligned ? static_cast<int>((static_cast<uint32_t>(ck.row0) * static_cast<uint32_t>(dim)) & 3u) : 0; };
    auto block_of = [&](int k, int a) { return sm + 16 + ((k - lo) % kStatRing) * slot_floats + 4 + a; };
    auto issue = [&](int k) {            // by warp 0 (by every thread for a matrix that is not 16-byte aligned: plain loads only)
        const PostChunk ck = chunks[k];
        const float *src = feat + ck.row0 * dim;
        const int cnt = ck.n * dim, a = shift_of(ck);
        const int body = base_aligned ? ((a + cnt) & ~3) : 0;
        float *block = block_of(k, a);
        if (tid == 0 && base_aligned) {     // every use of a slot completes one phase of its mbarrier, with or without a copy
            const uint32_t bar = smem_u32(sm) + 8 * ((k - lo) % kStatRing);
            if (body > 0) {
                mbar_expect_tx(bar, static_cast<uint32_t>(body) * 4u);
                bulk_g2s(smem_u32(block - a), src - a, static_cast<uint32_t>(body) * 4u, bar);
            } else {
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
            }
        }
        // what the bulk copy leaves out: the <= 3 floats behind the last 16-byte boundary (everything without it)
        const int stride = base_aligned ? 32 : kStatThreads;
        for (int i = (body > 0 ? body - a : 0) + tid; i < cnt; i += stride) block[i] = __ldg(src + i);
    };
    const bool issuer = tid < 32 || !base_aligned;
    if (tid == 0)
        for (int r = 0; r < kStatRing; ++r) mbar_init(smem_u32(sm) + 8 * r, 1);
    __syncthreads();
    if (issuer)
        for (int k = lo; k < hi && k < lo + kStatRing; ++k) issue(k);
    __syncthreads();

    float s0 = 0.0f, q0 = 0.0f, s1 = 0.0f, q1 = 0.0f, pivot = 0.0f;
    int gstart = lo, grows = 0;
    for (int k = lo; k < hi; ++k) {
        const PostChunk ck = chunks[k];
        const int cnt = ck.n * dim;
        const float *x = block_of(k, shift_of(ck));
        if (base_aligned) mbar_wait(smem_u32(sm) + 8 * ((k - lo) % kStatRing), static_cast<uint32_t>(((k - lo) / kStatRing) & 1));
        if (tid < per) {
            if (k == gstart) pivot = x[col];
            int i = tid;
            for (; i + per < cnt; i += 2 * per) {
                const float a = x[i] - pivot, b = x[i + per] - pivot;
                s0 += a; q0 = fmaf(a, a, q0);
                s1 += b; q1 = fmaf(b, b, q1);
            }
            if (i < cnt) {
                const float a = x[i] - pivot;
                s0 += a; q0 = fmaf(a, a, q0);
            }
        }
        grows += ck.n;
        const bool group_end = k + 1 == hi || k + 1 == ck.first_chunk + ck.n_chunks || (k + 1 - ck.first_chunk) % kStatGroup == 0;
        if (group_end) {
            s_s[tid] = static_cast<double>(s0) + static_cast<double>(s1);
            s_q[tid] = static_cast<double>(q0) + static_cast<double>(q1);
        }
        __syncthreads();                       // the slot is free (and the group's sums are in shared memory)
        if (issuer && k + kStatRing < hi) issue(k + kStatRing);
        if (!group_end) continue;
        // column d is met by threads d, d + dim, ...: G threads each add a share of them, then thread d adds the G sums
        const int G = nsub < 4 ? nsub : 4;
        if (sub < G) {
            double ts = 0.0, tq = 0.0;
            for (int j = sub; j < nsub; j += G) { ts += s_s[col + j * dim]; tq += s_q[col + j * dim]; }
            s2_s[tid] = ts;
            s2_q[tid] = tq;
        }
        __syncthreads();
        if (tid < dim) {
            double ts = 0.0, tq = 0.0;
            for (int j = 0; j < G; ++j) { ts += s2_s[tid + j * dim]; tq += s2_q[tid + j * dim]; }
            partial[static_cast<int64_t>(gstart) * dim + tid] = StatPartial{ts, tq, static_cast<double>(pivot), static_cast<double>(grows)};
        }
        s0 = q0 = s1 = q1 = 0.0f;
        gstart = k + 1;
        grows = 0;
    }
}

// ---- statistics, second step: one warp per utterance combines the group partials (lane = column) ----
// Group order, everything re-referenced to the utterance's first row P: sum (x - P) = s + n (p - P),
// sum (x - P)^2 = q + 2 (p - P) s + n (p - P)^2.  The result is stored as two-float pairs {mu_hi, mu_lo, inv_hi, inv_lo}: the
// apply kernel then normalises with FP32 instructions only, within 2 ulp of the double evaluation (x - mu_hi is exact or
// correctly rounded at the magnitude of the RESULT, which is what the tolerance is stated on).
__global__ void __launch_bounds__(kPostThreads)
post_finalize_kernel(const PostChunk *__restrict__ chunks, int chunk0, int n_chunks, int dim, int norm_var,
                     const StatPartial *__restrict__ partial, float4 *__restrict__ stats)
{
    const int w = static_cast<int>((static_cast<int64_t>(blockIdx.x) * kPostThreads + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (w >= n_chunks) return;
    const int c = chunk0 + w;
    const PostChunk ck = chunks[c];
    if (c != ck.first_chunk) return;                 // the warp of an utterance's first chunk does the utterance
    const double T = static_cast<double>(ck.f1 - ck.f0);
    for (int col = lane; col < dim; col += 32) {
        const StatPartial *p = partial + static_cast<int64_t>(ck.first_chunk) * dim + col;
        double P = 0.0, ts = 0.0, tq = 0.0;
        for (int j = 0; j < ck.n_chunks; j += kStatGroup) {
            const StatPartial e = p[static_cast<int64_t>(j) * dim];
            if (j == 0) P = e.pivot;
            const double dp = e.pivot - P;
            ts += e.s + e.n * dp;
            tq += e.q + 2.0 * dp * e.s + e.n * dp * dp;
        }
        const double m = ts / T;
        double var = tq / T - m * m;
        if (var < 0.0) var = 0.0;
        const double mu = P + m;
        const double inv = norm_var ? 1.0 / sqrt(var > 1e-20 ? var : 1e-20) : 1.0;
        const float mu_hi = static_cast<float>(mu), inv_hi = static_cast<float>(inv);
        stats[static_cast<int64_t>(ck.utt) * dim + col] =
            make_float4(mu_hi, static_cast<float>(mu - static_cast<double>(mu_hi)), inv_hi,
                        static_cast<float>(inv - static_cast<double>(inv_hi)));
    }
}

// regression sum of one column at row pointer x (row stride `dim` floats), run-time window: sum_k k (x[+k] - x[-k])
__device__ __forceinline__ float regress(const float *x, int dim, int W)
{
    float acc = 0.0f;
    for (int k = 1; k <= W; ++k) acc = fmaf(static_cast<float>(k), x[k * dim] - x[-k * dim], acc);
    return acc;
}

// ---- normalise + delta + delta-delta + stack: one CTA per chunk ----
// W_ > 0: regression window known at compile time (2 = the HTK / Kaldi default), 0: run-time window 1..8.
// The chunk and its halo rows arrive RAW by one bulk copy; the regressions are linear and their weights sum to zero, so they
// are taken on the raw values (delta of normalised = delta of raw times 1 / sigma) and the normalisation moves into the last
// phase, where a thread owns ONE output column and keeps its mean and scale in registers.  The regressions walk DOWN a
// column: a thread keeps the 2 W + 1 values of its window in registers and loads one new value per row.
template <int W_>
__global__ void __launch_bounds__(kPostThreads, MFCC_POST_MINB)
post_apply_kernel(const PostChunk *__restrict__ chunks, const float *feat,
                  const float4 *__restrict__ stats, const PostGeom g, float *out)   // chunks: first chunk of the launch; out may
                                                                                     // be feat when order == 0 (a CTA reads its rows, then writes them)
{
    extern __shared__ __align__(16) float sm[];   // [4: mbarrier][kPostFront + (rows + 2 HX) dim + 4][(rows + 2 HD) dim]
    const int W = W_ ? W_ : g.window;
    const int dim = g.dim, order = g.order, per = g.per;
    const PostChunk ck = chunks[blockIdx.x];
    const int tid = threadIdx.x, n = ck.n;
    const int HX = order * W;                 // halo rows of X on each side (order 2: the delta of the +-W rows needs +-2W)
    const int HD = order == 2 ? W : 0;        // halo rows of D1
    const int sub = div20(tid, g.m_dim), col = tid - sub * dim;
    const uint32_t bar = smem_u32(sm);

    // phase A: the chunk and its halo, raw, by one bulk copy; rows outside the utterance are copies of its edge rows
    const int64_t lo = max(ck.f0, ck.row0 - HX), hi = min(ck.f1, ck.row0 + n + HX);
    const int nb = static_cast<int>(lo - (ck.row0 - HX));       // halo rows missing below
    const int na = static_cast<int>(ck.row0 + n + HX - hi);     // and above
    const int cnt = static_cast<int>(hi - lo) * dim;
    const bool base_aligned = (reinterpret_cast<uintptr_t>(feat) & 15) == 0;
    // the staged block starts nb rows into X: leave room for them in front
    const Staged st = stage_issue(sm + 4, HX * dim + 4, bar, feat + lo * dim, cnt, base_aligned, tid);
    float *X = st.block - nb * dim;            // [n + 2 HX][dim], row r of the chunk at X + (r + HX) dim
    float *D1 = sm + 4 + kPostFront + 4 + (g.rows + 3 * HX) * dim;   // [n + 2 HD][dim]: unscaled first regression (order 2 only)
    // this thread's output column in the last phase: its mean / scale pair, fetched while the copy flies
    const int OD = g.od;
    const int blk = div20(tid, g.m_cw), cl = tid - blk * g.cw;
    stage_wait(st, bar);
    if (nb > 0 || na > 0) {
        const float *first = st.block, *last = st.block + cnt - dim;
        if (tid < per) {
            for (int i = tid; i < nb * dim; i += per) X[i] = first[col];
            for (int i = tid; i < na * dim; i += per) st.block[cnt + i] = last[col];
        }
        __syncthreads();
    }

    // phase B (order 2): the first regression (unscaled sums) at the positions [row0 - W, row0 + n + W) that lie inside the
    // utterance (D1 row q = position row0 - W + q = X row q + W); positions outside take the value of the utterance's edge row
    if (order == 2) {
        const int qlo = static_cast<int>(max(ck.f0 - (ck.row0 - HD), static_cast<int64_t>(0)));
        const int qhi = static_cast<int>(min(ck.f1 - (ck.row0 - HD), static_cast<int64_t>(n + 2 * HD)));
        if (tid < per) {
            const int rpt = div20(qhi - qlo + g.nsub - 1, g.m_nsub);
            const int q0 = qlo + sub * rpt, q1 = min(qhi, q0 + rpt);
            if (q0 < q1) {
                const float *x = X + (q0 + W) * dim + col;
                float *o = D1 + q0 * dim + col;
                if constexpr (W_ == 2) {
                    float m2 = x[-2 * dim], m1 = x[-dim], c0 = x[0], p1 = x[dim];
                    x += 2 * dim;
#pragma unroll 5
                    for (int q = q0; q < q1; ++q) {
                        const float p2 = *x;
                        *o = fmaf(2.0f, p2 - m2, p1 - m1);
                        x += dim;
                        o += dim;
                        m2 = m1; m1 = c0; c0 = p1; p1 = p2;
                    }
                } else {
                    for (int q = q0; q < q1; ++q, x += dim, o += dim) *o = regress(x, dim, W);
                }
            }
        }
        __syncthreads();
        if (qlo > 0 || qhi < n + 2 * HD) {
            const float *first = D1 + qlo * dim, *last = D1 + (qhi - 1) * dim;
            if (tid < per) {
                for (int i = tid; i < qlo * dim; i += per) D1[i] = first[col];
                for (int i = tid; i < (n + 2 * HD - qhi) * dim; i += per) D1[qhi * dim + i] = last[col];
            }
            __syncthreads();
        }
    }

    // phase C: thread = output column (static | delta | delta-delta) x a block of consecutive rows, so that at every step
    // the lanes of a warp write consecutive floats of one (or two) output rows, and the next step continues right behind
    // them.  One loop body for all three parts (a warp spans two or three of them): the regression is evaluated on every
    // lane — the halo makes its loads legal for the static columns too — the static columns select the centred value, and
    // one two-float scale per column finishes both: 1 / sigma for the static part, 1 / (sigma den) and 1 / (sigma den^2) for
    // the regressions (den = 2 sum k^2).
    float *orow = out + ck.row0 * OD;
    if (blk >= g.nblk) return;
    const int rpt = div20(n + g.nblk - 1, g.m_nblk);
    const int r0 = blk * rpt, r1 = min(n, r0 + rpt);
    if (r0 >= r1) return;
    for (int c = cl; c < OD; c += g.cw) {
        const int pc = div20(c, g.m_dim), d = c - pc * dim, part = pc + g.part0;
        const float4 ms = g.cmvn ? __ldg(stats + static_cast<int64_t>(ck.utt) * dim + d) : make_float4(0.0f, 0.0f, 1.0f, 0.0f);
        const float sc = part == 0 ? 1.0f : (part == 1 ? g.inv_den : g.inv_den * g.inv_den);
        const float k_hi = ms.z * sc;
        const float k_lo = fmaf(ms.z, sc, -k_hi) + ms.w * sc;     // (inv_hi + inv_lo) sc = k_hi + k_lo to two-float accuracy
        const float *x = (part == 2 ? D1 + HD * dim : X + HX * dim) + r0 * dim + d;
        float *o = orow + r0 * OD + c;
        const bool stat = part == 0;
        if (order == 0) {
            for (int r = r0; r < r1; ++r, x += dim, o += OD) {
                const float t = (*x - ms.x) - ms.y;
                *o = fmaf(t, k_lo, t * k_hi);
            }
        } else if constexpr (W_ == 2) {
            float m2 = x[-2 * dim], m1 = x[-dim], c0 = x[0], p1 = x[dim];
            x += 2 * dim;
#pragma unroll 5
            for (int r = r0; r < r1; ++r) {
                const float p2 = *x;
                const float v = fmaf(2.0f, p2 - m2, p1 - m1);
                const float t = stat ? (c0 - ms.x) - ms.y : v;
                *o = fmaf(t, k_lo, t * k_hi);
                x += dim;
                o += OD;
                m2 = m1; m1 = c0; c0 = p1; p1 = p2;
            }
        } else {
            for (int r = r0; r < r1; ++r, x += dim, o += OD) {
                const float t = stat ? (x[0] - ms.x) - ms.y : regress(x, dim, W);
                *o = fmaf(t, k_lo, t * k_hi);
            }
        }
    }
}

std::atomic<uint64_t> g_optin2{0}, g_optin0{0}, g_optin_stats{0};

}  // namespace

int post_rows_for(int dim)
{
    // ~4,096 staged elements per chunk (27 KB of shared memory with both halos at window 2): 256 rows of 13 cepstra, 48 rows
    // of an 80-band log-mel matrix
    int rows = (4096 / std::max(dim, 1)) & ~7;
    rows = std::min(256, std::max(32, rows));
    if (const char *e = std::getenv("MFCC_POST_ROWS")) {   // tuning experiments only (tools/gpu_post_rows.sh)
        const int v = std::atoi(e);
        if (v >= 8 && v <= 1024 && post_smem_bytes(dim, v, 8, 2) <= kPostSmemMax) rows = v;
    }
    return rows;
}

// Chunk table of a batch: utterance u's rows [f0, f1) cut into pieces of at most `rows` rows; utt_first[u] = index of its
// first chunk (utt_first[n_utts] = number of chunks), so a range of utterances is a range of chunks.
void post_build_chunks(const std::vector<int64_t> &frame_offsets, int dim, std::vector<PostChunk> &chunks,
                       std::vector<int64_t> &utt_first, int *rows_out)
{
    const int rows = post_rows_for(dim);
    *rows_out = rows;
    chunks.clear();
    const int64_t n_utts = static_cast<int64_t>(frame_offsets.size()) - 1;
    utt_first.assign(static_cast<size_t>(std::max<int64_t>(n_utts, 0)) + 1, 0);
    for (int64_t u = 0; u < n_utts; ++u) {
        utt_first[u] = static_cast<int64_t>(chunks.size());
        const int64_t f0 = frame_offsets[u], f1 = frame_offsets[u + 1];
        if (f1 <= f0) continue;
        const int64_t nc = (f1 - f0 + rows - 1) / rows;
        const int32_t first = static_cast<int32_t>(chunks.size());
        for (int64_t c = 0; c < nc; ++c) {
            PostChunk ck{};
            ck.row0 = f0 + c * rows;
            ck.f0 = f0;
            ck.f1 = f1;
            ck.n = static_cast<int32_t>(std::min<int64_t>(rows, f1 - ck.row0));
            ck.utt = static_cast<int32_t>(u);
            ck.first_chunk = first;
            ck.n_chunks = static_cast<int32_t>(nc);
            chunks.push_back(ck);
        }
    }
    if (n_utts >= 0) utt_first[n_utts] = static_cast<int64_t>(chunks.size());
}

size_t post_smem_bytes(int dim, int rows, int window, int order)
{
    // mbarrier slot, alignment room of the bulk copy, X with its halo (the staged block may start up to one halo further in),
    // 4 floats behind it, D1 with its halo
    const int hx = order * window, hd = order == 2 ? window : 0;
    return sizeof(float) * (4 + kPostFront + 4 + static_cast<size_t>(dim) * ((rows + 3 * hx) + (order == 2 ? rows + 2 * hd : 0)));
}

int launch_post(const PostView &v, const float *d_feat, int dim, int cmvn, int window, int order, int part0, float *d_out,
                cudaStream_t s)
{
    if (v.n_chunks <= 0) return MFCC_OK;
    if (dim > kPostMaxDim || dim > kStatThreads || v.chunks == nullptr || v.chunk0 + v.n_chunks > (1 << 30)) return MFCC_EINVAL;
    PostGeom g{};
    g.dim = dim;
    g.rows = v.rows;
    g.order = order;
    g.window = window;
    g.cmvn = cmvn != MFCC_CMVN_NONE;
    g.part0 = part0;
    g.per = (kPostThreads / dim) * dim;
    g.nsub = g.per / dim;
    g.s_per = (kStatThreads / dim) * dim;
    g.s_nsub = g.s_per / dim;
    g.od = dim * (1 + order - part0);
    g.cw = std::min(g.od, kPostThreads);
    g.nblk = kPostThreads / g.cw;
    g.m_dim = magic20(dim);
    g.m_cw = magic20(g.cw);
    g.m_nblk = magic20(g.nblk);
    g.m_nsub = magic20(g.nsub);
    double den = 0.0;
    for (int k = 1; k <= window; ++k) den += 2.0 * k * k;
    g.inv_den = static_cast<float>(1.0 / den);
    const unsigned grid = static_cast<unsigned>(v.n_chunks);
    if (cmvn != MFCC_CMVN_NONE) {
        const size_t smem_stats = sizeof(float) * (16 + static_cast<size_t>(kStatRing) * (8 + ((static_cast<size_t>(v.rows) * dim + 3) & ~static_cast<size_t>(3))));
        if (smem_stats > kPostSmemMax) return MFCC_EINVAL;
        // (the kernel also holds 8 KB of static shared memory: opt in whenever the sum could pass 48 KB — once per device)
        if (smem_stats > 36 * 1024 && ensure_smem_optin(post_stats_kernel, v.device, kPostSmemMax, g_optin_stats) != MFCC_OK) return MFCC_ECUDA;
        // persistent: MFCC_POST_STAT_CTAS CTAs per SM, at least two chunks each
        const int64_t want = std::min<int64_t>(MFCC_POST_STAT_CTAS * static_cast<int64_t>(std::max(v.sms, 1)), (v.n_chunks + 1) / 2);
        post_stats_kernel<<<static_cast<unsigned>(std::max<int64_t>(want, 1)), kStatThreads, smem_stats, s>>>(
            v.chunks, static_cast<int>(v.chunk0), static_cast<int>(v.n_chunks), d_feat, g, static_cast<StatPartial *>(v.partial));
        post_finalize_kernel<<<static_cast<unsigned>((v.n_chunks * 32 + kPostThreads - 1) / kPostThreads), kPostThreads, 0, s>>>(
            v.chunks, static_cast<int>(v.chunk0), static_cast<int>(v.n_chunks), dim, cmvn == MFCC_CMVN_MEAN_VAR,
            static_cast<const StatPartial *>(v.partial), static_cast<float4 *>(v.stats));
        g_launches.fetch_add(2, std::memory_order_relaxed);
    }
    const size_t smem = post_smem_bytes(dim, v.rows, window, order);
    if (smem > kPostSmemMax) return MFCC_EINVAL;
    if (window == 2) {
        if (smem > 48 * 1024 && ensure_smem_optin(post_apply_kernel<2>, v.device, kPostSmemMax, g_optin2) != MFCC_OK)
            return MFCC_ECUDA;
        post_apply_kernel<2><<<grid, kPostThreads, smem, s>>>(v.chunks + v.chunk0, d_feat, static_cast<const float4 *>(v.stats), g, d_out);
    } else {
        if (smem > 48 * 1024 && ensure_smem_optin(post_apply_kernel<0>, v.device, kPostSmemMax, g_optin0) != MFCC_OK)
            return MFCC_ECUDA;
        post_apply_kernel<0><<<grid, kPostThreads, smem, s>>>(v.chunks + v.chunk0, d_feat, static_cast<const float4 *>(v.stats), g, d_out);
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError() == cudaSuccess ? MFCC_OK : MFCC_ECUDA;
}

}  // namespace mfcc
