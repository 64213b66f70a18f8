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
//   * the three output parts of a row are produced by the same sweep (thread = output column, rows strided), so an output
//     sector is written once, completely, by neighbouring threads — not by three passes with a 3 * dim stride;
//   * the chunk (plus the +-2 * window halo rows the two regressions need, edge rows replicated as the spec says) is
//     staged in shared memory once; the delta of the halo rows is recomputed instead of exchanged;
//   * the grid is one CTA per chunk with 8 CTAs resident per SM (27 KB of shared memory each for 13 cepstra), enough
//     loads in flight to cover the HBM latency; nothing is allocated per call.
// Statistics: chunk sums of (x - pivot) and (x - pivot)^2 in double (pivot = the utterance's first row, which keeps the
// uncentred variance formula free of cancellation), one partial per chunk, combined in chunk order by the LAST chunk of
// the utterance to finish (counter + fence) — deterministic, no floating-point atomics, no extra launch.
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

constexpr int kPostThreads = 256;
constexpr int kPostMaxDim = 256;   // dim <= n_mel + 1 <= 129 for any valid plan
constexpr int kStatGroup = 4;      // chunks one CTA of the statistics kernel walks (consecutive chunks of one utterance are summed as one block)

// Everything the kernels would otherwise derive with integer divisions by run-time values, prepared on the host: a chunk
// gives a thread only ~40 output elements, so a handful of 20-instruction divisions per thread is a third of the work.
struct PostGeom {
    int dim, rows, order, window, cmvn;
    int per;              // (256 / dim) * dim: thread t < per always meets column t % dim
    int nsub;             // per / dim
    int od, cw, nblk;     // output columns, columns per pass of the stacking sweep, row blocks in it
    unsigned m_dim, m_cw, m_nblk, m_nsub;   // t / d == (t * m) >> 20 for t < 4096, d <= 256
    float inv_den;
};
__device__ __forceinline__ int div20(int t, unsigned m) { return static_cast<int>((static_cast<unsigned>(t) * m) >> 20); }
unsigned magic20(int d) { return ((1u << 20) + static_cast<unsigned>(d) - 1u) / static_cast<unsigned>(d); }

// ---- per-utterance statistics: one CTA per group of kStatGroup chunks, the utterance's last CTA to finish combines the partials ----
// Groups are aligned to the utterance (chunks first + 4 k .. first + 4 k + 3), so the summation order — and with it every
// bit of the result — does not depend on which range of the batch a launch covers; the CTAs of the other chunks exit at once.
// Inside a thread the values it meets are summed in f32 about the pivot (|x - pivot| is a few sigma, so the sums keep 7
// digits of a quantity whose mean needs 5); across threads and chunks everything is double, in a fixed order.
// The result is stored as two-float pairs {mu_hi, mu_lo, inv_hi, inv_lo}: the apply kernel then normalises with four FP32
// instructions and no conversions, within 2 ulp of the double evaluation (x - mu_hi is exact or correctly rounded at the
// magnitude of the RESULT, which is what the tolerance is stated on).
__global__ void __launch_bounds__(kPostThreads, 6)
post_stats_kernel(const PostChunk *__restrict__ chunks, int chunk0, const float *__restrict__ feat, const PostGeom g,
                  int norm_var, double2 *__restrict__ partial, float4 *__restrict__ stats, unsigned *__restrict__ count)
{
    __shared__ double s_s[kPostThreads], s_q[kPostThreads], s2_s[kPostThreads], s2_q[kPostThreads];
    __shared__ int s_last;
    const int tid = threadIdx.x, dim = g.dim, per = g.per;
    const int col = tid - div20(tid, g.m_dim) * dim, sub = div20(tid, g.m_dim);
    {
        const int c = chunk0 + static_cast<int>(blockIdx.x);
        const PostChunk ck = chunks[c];
        if ((c - ck.first_chunk) % kStatGroup != 0) return;
        const int e = min(c + kStatGroup, ck.first_chunk + ck.n_chunks);   // run = the chunks c .. e - 1: one contiguous block
        int run_rows = ck.n;
        for (int k = c + 1; k < e; ++k) run_rows += chunks[k].n;
        const int total = run_rows * dim;
        float s0 = 0.0f, q0 = 0.0f, s1 = 0.0f, q1 = 0.0f;
        if (tid < per) {
            const float pivot = __ldg(feat + ck.f0 * dim + col);
            const float *src = feat + ck.row0 * dim;
            int i = tid;
            for (; i < total; i += 8 * per) {     // eight independent loads in flight per thread (past the end: the pivot, which adds 0)
                float v[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) v[k] = i + k * per < total ? __ldg(src + i + k * per) : pivot;
#pragma unroll
                for (int k = 0; k < 8; k += 2) {
                    const float a = v[k] - pivot, b = v[k + 1] - pivot;
                    s0 += a; q0 = fmaf(a, a, q0);
                    s1 += b; q1 = fmaf(b, b, q1);
                }
            }
        }
        s_s[tid] = static_cast<double>(s0) + static_cast<double>(s1);
        s_q[tid] = static_cast<double>(q0) + static_cast<double>(q1);
        __syncthreads();
        // column d is met by threads d, d + dim, ...: G threads each add a share of them, then thread d adds the G sums
        const int G = g.nsub < 4 ? g.nsub : 4;
        if (sub < G) {
            double ts = 0.0, tq = 0.0;
            for (int j = sub; j < g.nsub; j += G) { ts += s_s[col + j * dim]; tq += s_q[col + j * dim]; }
            s2_s[tid] = ts;
            s2_q[tid] = tq;
        }
        __syncthreads();
        if (tid < dim) {
            double ts = 0.0, tq = 0.0;
            for (int k = 0; k < G; ++k) { ts += s2_s[tid + k * dim]; tq += s2_q[tid + k * dim]; }
            partial[static_cast<int64_t>(c) * dim + tid] = make_double2(ts, tq);
            for (int k = c + 1; k < e; ++k) partial[static_cast<int64_t>(k) * dim + tid] = make_double2(0.0, 0.0);
        }
        __syncthreads();
        if (tid == 0) {      // the fence is cumulative: the partials the other threads wrote before the barrier are ordered with it
            __threadfence();
            s_last = atomicAdd(&count[ck.utt], static_cast<unsigned>(e - c)) + static_cast<unsigned>(e - c) == static_cast<unsigned>(ck.n_chunks);
            __threadfence();
        }
        __syncthreads();
        if (s_last) {
            if (tid < dim) {
                double ts = 0.0, tq = 0.0;
                const double2 *p = partial + static_cast<int64_t>(ck.first_chunk) * dim + tid;
                for (int k = 0; k < ck.n_chunks; ++k) {      // chunk order: the result does not depend on which CTA came last
                    const double2 v = __ldcg(p + static_cast<int64_t>(k) * dim);
                    ts += v.x;
                    tq += v.y;
                }
                const double T = static_cast<double>(ck.f1 - ck.f0);
                const double m = ts / T;
                double var = tq / T - m * m;
                if (var < 0.0) var = 0.0;
                const double mu = static_cast<double>(__ldg(feat + ck.f0 * dim + tid)) + m;
                const double inv = norm_var ? 1.0 / sqrt(var > 1e-20 ? var : 1e-20) : 1.0;
                const float mu_hi = static_cast<float>(mu), inv_hi = static_cast<float>(inv);
                stats[static_cast<int64_t>(ck.utt) * dim + tid] =
                    make_float4(mu_hi, static_cast<float>(mu - static_cast<double>(mu_hi)), inv_hi,
                                static_cast<float>(inv - static_cast<double>(inv_hi)));
            }
            if (tid == 0) count[ck.utt] = 0;   // the counters are zero again for the next (stream-ordered) call
        }
    }
}

// regression of one column at row pointer x (row stride `dim` floats), run-time window
__device__ __forceinline__ float regress(const float *x, int dim, int W, float inv_den)
{
    float acc = 0.0f;
    for (int k = 1; k <= W; ++k) acc = fmaf(static_cast<float>(k), x[k * dim] - x[-k * dim], acc);
    return acc * inv_den;
}

// ---- normalise + delta + delta-delta + stack: one CTA per chunk ----
// W_ > 0: regression window known at compile time (2 = the HTK / Kaldi default), 0: run-time window 1..8.
// The regressions walk DOWN a column: a thread keeps the 2 W + 1 values of its window in registers and loads one new
// value per row (W_ = 2: one LDS, two FADD, one FFMA, one FMUL, one select, one store per output element).
template <int W_>
__global__ void __launch_bounds__(kPostThreads, 8)
post_apply_kernel(const PostChunk *__restrict__ chunks, const float *__restrict__ feat,
                  const float4 *__restrict__ stats, const PostGeom g, float *__restrict__ out)   // chunks: first chunk of the launch
{
    extern __shared__ __align__(16) float sm[];
    const int W = W_ ? W_ : g.window;
    const int dim = g.dim, order = g.order, per = g.per;
    const float inv_den = g.inv_den;
    const PostChunk ck = chunks[blockIdx.x];
    const int tid = threadIdx.x, n = ck.n;
    const int HX = order * W;                 // halo rows of X on each side (order 2: the delta of the +-W rows needs +-2W)
    const int HD = order == 2 ? W : 0;        // halo rows of D1
    float *X = sm;                            // [rows + 2 HX][dim]  normalised features, edge rows replicated
    float *D1 = X + (g.rows + 2 * HX) * dim;  // [rows + 2 HD][dim]  first regression (order 2 only)
    const int sub = div20(tid, g.m_dim), col = tid - sub * dim;

    // phase A: the chunk and its halo, one contiguous sweep; rows outside the utterance are copies of its edge rows
    const int64_t lo = max(ck.f0, ck.row0 - HX), hi = min(ck.f1, ck.row0 + n + HX);
    const int nb = static_cast<int>(lo - (ck.row0 - HX));       // halo rows missing below
    const int na = static_cast<int>(ck.row0 + n + HX - hi);     // and above
    const int cnt = static_cast<int>(hi - lo) * dim;
    {
        const float *src = feat + lo * dim;
        float *dst = X + nb * dim;
        if (g.cmvn) {
            if (tid < per) {
                const float4 st = __ldg(stats + static_cast<int64_t>(ck.utt) * dim + col);
                int i = tid;
                for (; i + 3 * per < cnt; i += 4 * per) {       // four independent loads in flight per thread
                    const float a = __ldg(src + i), b = __ldg(src + i + per), c = __ldg(src + i + 2 * per), d = __ldg(src + i + 3 * per);
                    const float ta = (a - st.x) - st.y, tb = (b - st.x) - st.y, tc = (c - st.x) - st.y, td = (d - st.x) - st.y;
                    dst[i] = fmaf(ta, st.w, ta * st.z);
                    dst[i + per] = fmaf(tb, st.w, tb * st.z);
                    dst[i + 2 * per] = fmaf(tc, st.w, tc * st.z);
                    dst[i + 3 * per] = fmaf(td, st.w, td * st.z);
                }
                if (i < cnt) {                                   // up to three left: loaded together as well
                    const bool h1 = i + per < cnt, h2 = i + 2 * per < cnt;
                    const float a = __ldg(src + i), b = h1 ? __ldg(src + i + per) : 0.0f, c = h2 ? __ldg(src + i + 2 * per) : 0.0f;
                    const float ta = (a - st.x) - st.y, tb = (b - st.x) - st.y, tc = (c - st.x) - st.y;
                    dst[i] = fmaf(ta, st.w, ta * st.z);
                    if (h1) dst[i + per] = fmaf(tb, st.w, tb * st.z);
                    if (h2) dst[i + 2 * per] = fmaf(tc, st.w, tc * st.z);
                }
            }
        } else {
            int i = tid;
            for (; i + 3 * kPostThreads < cnt; i += 4 * kPostThreads) {
                const float a = __ldg(src + i), b = __ldg(src + i + kPostThreads);
                const float c = __ldg(src + i + 2 * kPostThreads), d = __ldg(src + i + 3 * kPostThreads);
                dst[i] = a;
                dst[i + kPostThreads] = b;
                dst[i + 2 * kPostThreads] = c;
                dst[i + 3 * kPostThreads] = d;
            }
            if (i < cnt) {
                const bool h1 = i + kPostThreads < cnt, h2 = i + 2 * kPostThreads < cnt;
                const float a = __ldg(src + i), b = h1 ? __ldg(src + i + kPostThreads) : 0.0f;
                const float c = h2 ? __ldg(src + i + 2 * kPostThreads) : 0.0f;
                dst[i] = a;
                if (h1) dst[i + kPostThreads] = b;
                if (h2) dst[i + 2 * kPostThreads] = c;
            }
        }
    }
    __syncthreads();
    if (nb > 0 || na > 0) {
        const float *first = X + nb * dim, *last = X + nb * dim + cnt - dim;
        if (tid < per) {
            for (int i = tid; i < nb * dim; i += per) X[i] = first[col];
            for (int i = tid; i < na * dim; i += per) X[nb * dim + cnt + i] = last[col];
        }
        __syncthreads();
    }

    // phase B (order 2): the first regression at the positions [row0 - W, row0 + n + W) that lie inside the utterance
    // (D1 row q = position row0 - W + q = X row q + W); positions outside take the value of the utterance's edge row
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
                        *o = fmaf(2.0f, p2 - m2, p1 - m1) * inv_den;
                        x += dim;
                        o += dim;
                        m2 = m1; m1 = c0; c0 = p1; p1 = p2;
                    }
                } else {
                    for (int q = q0; q < q1; ++q, x += dim, o += dim) *o = regress(x, dim, W, inv_den);
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
    // lane — the halo makes its loads legal for the static columns too — and the static columns select the centre value.
    const int OD = g.od;
    float *orow = out + ck.row0 * OD;
    if (order == 0) {
        for (int i = tid; i < n * dim; i += kPostThreads) orow[i] = X[i];
        return;
    }
    const int blk = div20(tid, g.m_cw), cl = tid - blk * g.cw;
    if (blk >= g.nblk) return;
    const int rpt = div20(n + g.nblk - 1, g.m_nblk);
    const int r0 = blk * rpt, r1 = min(n, r0 + rpt);
    if (r0 >= r1) return;
    for (int c = cl; c < OD; c += g.cw) {
        const int part = div20(c, g.m_dim), d = c - part * dim;
        const float *x = (part == 2 ? D1 + HD * dim : X + HX * dim) + r0 * dim + d;
        float *o = orow + r0 * OD + c;
        const bool stat = part == 0;
        if constexpr (W_ == 2) {
            float m2 = x[-2 * dim], m1 = x[-dim], c0 = x[0], p1 = x[dim];
            x += 2 * dim;
#pragma unroll 5
            for (int r = r0; r < r1; ++r) {
                const float p2 = *x;
                const float v = fmaf(2.0f, p2 - m2, p1 - m1) * inv_den;
                *o = stat ? c0 : v;
                x += dim;
                o += OD;
                m2 = m1; m1 = c0; c0 = p1; p1 = p2;
            }
        } else {
            for (int r = r0; r < r1; ++r, x += dim, o += OD) *o = stat ? x[0] : regress(x, dim, W, inv_den);
        }
    }
}

std::atomic<uint64_t> g_optin2{0}, g_optin0{0};

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
    const int hx = order * window, hd = order == 2 ? window : 0;
    return sizeof(float) * static_cast<size_t>(dim) * ((rows + 2 * hx) + (order == 2 ? rows + 2 * hd : 0));
}

int launch_post(const PostView &v, const float *d_feat, int dim, int cmvn, int window, int order, float *d_out,
                cudaStream_t s)
{
    if (v.n_chunks <= 0) return MFCC_OK;
    if (dim > kPostMaxDim || v.chunks == nullptr || v.chunk0 + v.n_chunks > (1 << 30)) return MFCC_EINVAL;
    PostGeom g{};
    g.dim = dim;
    g.rows = v.rows;
    g.order = order;
    g.window = window;
    g.cmvn = cmvn != MFCC_CMVN_NONE;
    g.per = (kPostThreads / dim) * dim;
    g.nsub = g.per / dim;
    g.od = dim * (1 + order);
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
        post_stats_kernel<<<grid, kPostThreads, 0, s>>>(
            v.chunks, static_cast<int>(v.chunk0), d_feat, g, cmvn == MFCC_CMVN_MEAN_VAR,
            static_cast<double2 *>(v.partial), static_cast<float4 *>(v.stats), v.count);
        g_launches.fetch_add(1, std::memory_order_relaxed);
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
