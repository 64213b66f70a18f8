// mfcc_fused.cu — the fused 32-frame-tile MFCC kernels (the performance path).
//
// Layout idea (DESIGN.md "Fused kernel"): a CTA owns one tile = up to 32
// consecutive frames of one utterance; LANE = FRAME.  Every shared-memory
// array is indexed [...][lane], so all accesses are bank-conflict free, and
// every constant a butterfly needs (window, twiddle, mel weight, DCT entry) is
// WARP-UNIFORM: the warp index selects the butterfly, the lane only selects
// the frame.  Phases, separated by CTA barriers:
//   S0 stage : PCM tile (read from HBM exactly once) -> f32 -> pre-emphasis -> smem,
//              one pad pair per hop so that lane stride is 2*odd words
//   S1 pass 1: R2 radix-R1 butterflies over the windowed, zero-padded frame
//              (packed as N/2 complex points) + inter-pass twiddle -> workspace
//   S2 pass 2: R1 radix-R2 butterflies, two per thread (k1 and R1-k1) so that
//              the real-FFT split X[k] / X[N/2-k] and |X|^2 happen in registers
//   S3 mel   : per-segment rise/fall sums of the power spectrum (sparse filterbank)
//   S4 log   : E[m] = rise[m] + fall[m+1], ln(max(E, floor))
//   S5 dct   : n_cep x n_mel contraction, store
//
// No reference code corresponds to this (SURVEY.md §8a "Ref file:line = none").
#include <cuda_runtime.h>

#include <cmath>
#include <cstring>
#include <vector>

#include "mfcc_fft.cuh"
#include "mfcc_host.h"

namespace mfcc {

namespace {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int kPad = 2;  // words inserted per hop in the staged tile

// Per-plan tables of the fused kernel, one device blob.
struct FusedTables {
    const float2 *win2;     // [N/2]   (w[2n], w[2n+1]), zero past frame_len
    const float2 *tw;       // [R2][R1] inter-pass twiddles W_{N/2}^{n2*k1}
    const float2 *post;     // [N/2+1] exp(-2 pi i k / N): the split twiddle of bin k
    const float *rise;      // [N/2+1]
    const float *fall;      // [N/2+1]
    const int32_t *bins;    // [n_mel+2]
    const int32_t *seg_lo;  // [kWarps+1] segment range per warp for S3
    const float *dct;       // [n_out][n_mel]
};

struct FusedArgs {
    const Tile *tiles;
    float *out;
    FusedTables t;
    int frame_len, hop, n_mel, n_cep, logmel;
    float preemph, log_floor;
    int staged_words;  // size of the staged/P union region in floats
};

__device__ __forceinline__ float to_f32(int16_t v) { return static_cast<float>(v); }
__device__ __forceinline__ float to_f32(float v) { return v; }

// NFFT = 2 * R1 * R2.  R2 must be 16 (one pass-1 butterfly per half... see mapping below).
template <typename PcmT, int R1, int R2>
__global__ void __launch_bounds__(kThreads, 2) fused_rt_kernel(const PcmT *__restrict__ pcm, FusedArgs a)
{
    constexpr int N2 = R1 * R2;       // complex points
    constexpr int NFFT = 2 * N2;
    constexpr int NB = N2 + 1;        // power bins
    extern __shared__ __align__(16) float smem[];
    float *staged = smem;                                  // [staged_words]  (S0-S1), aliased by
    float *pw = smem;                                      // [NB][32]        (S2-S3)
    float2 *ws = reinterpret_cast<float2 *>(smem + a.staged_words);  // [R1][R2][32] complex
    float *ws_f = reinterpret_cast<float *>(ws);           // S3-S5: er | ef | lg, each [.][32]

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const Tile tile = a.tiles[blockIdx.x];
    const int hop = a.hop, L = a.frame_len, stride = hop + kPad;

    // ---- S0: stage the PCM tile once: y[s] = x[s] - a x[s-1] (0 before the utterance,
    //      0 past its end), f32, at word (i / hop) * (hop + kPad) + i % hop.
    {
        const int T = (tile.n_frames - 1) * hop + L;
        const PcmT *x = pcm + tile.first_sample;
        const int64_t room_lo = tile.first_sample - tile.utt_begin;   // samples available before i = 0
        const int64_t room_hi = tile.utt_end - tile.first_sample;     // samples available from i = 0
        int h = 0, r = threadIdx.x;       // i = h * hop + r, advanced by kThreads each step
        while (r >= hop) { r -= hop; ++h; }
        for (int i = threadIdx.x; i <= T; i += kThreads) {   // word T is read (times a zero window) when L is odd
            float y = 0.0f;
            if (i < room_hi && i < T) {
                const float x0 = to_f32(x[i]);
                const float x1 = (i > -room_lo) ? to_f32(x[i - 1]) : 0.0f;
                y = fmaf(-a.preemph, x1, x0);
            }
            staged[h * stride + r] = y;
            r += kThreads;
            while (r >= hop) { r -= hop; ++h; }
        }
    }
    __syncthreads();

    // ---- S1: pass 1.  Butterfly n2 takes z[n2 + R2 n1], n1 < R1, z[n] = (y[2n], y[2n+1]) * window.
    {
        const float *base = staged + lane * stride;
        constexpr int PER_WARP = R2 / kWarps;
#pragma unroll 1
        for (int j = 0; j < PER_WARP; ++j) {
            const int n2 = warp * PER_WARP + j;
            cplx x[R1];
#pragma unroll
            for (int n1 = 0; n1 < R1; ++n1) {
                const int n = n2 + R2 * n1, i = 2 * n;
                x[n1] = cplx{0.0f, 0.0f};
                if (i < L) {  // warp-uniform
                    const int hh = i / hop, rr = i - hh * hop;
                    const float2 y = *reinterpret_cast<const float2 *>(base + hh * stride + rr);
                    const float2 w = __ldg(a.t.win2 + n);
                    x[n1] = cplx{y.x * w.x, y.y * w.y};
                }
            }
            Dft<R1>::run(x);
            ws[(0 * R2 + n2) * 32 + lane] = make_float2(x[0].re, x[0].im);
#pragma unroll
            for (int k1 = 1; k1 < R1; ++k1) {
                const float2 t = __ldg(a.t.tw + n2 * R1 + k1);
                const cplx v = cmulc(x[k1], t.x, t.y);
                ws[(k1 * R2 + n2) * 32 + lane] = make_float2(v.re, v.im);
            }
        }
    }
    __syncthreads();

    // ---- S2: pass 2, real-FFT split, power.  Work item it handles butterflies
    //      ka = it, kb = R1 - it (it = 0 pairs 0 with R1/2).  R1/2 items over kWarps warps.
    {
        constexpr int ITEMS = R1 / 2;
        for (int it = warp; it < ITEMS; it += kWarps) {
            const int ka = it, kb = it == 0 ? R1 / 2 : R1 - it;
            cplx u[R2], v[R2];
#pragma unroll
            for (int n2 = 0; n2 < R2; ++n2) {
                const float2 t = ws[(ka * R2 + n2) * 32 + lane];
                u[n2] = cplx{t.x, t.y};
            }
#pragma unroll
            for (int n2 = 0; n2 < R2; ++n2) {
                const float2 t = ws[(kb * R2 + n2) * 32 + lane];
                v[n2] = cplx{t.x, t.y};
            }
            Dft<R2>::run(u);  // u[k2] = Z[ka + R1 k2]
            Dft<R2>::run(v);  // v[k2] = Z[kb + R1 k2]
            if (it != 0) {
                // bin k = ka + R1 k2 pairs with N2 - k = kb + R1 (R2 - 1 - k2)
#pragma unroll
                for (int k2 = 0; k2 < R2; ++k2) {
                    const int k = ka + R1 * k2, m = N2 - k;
                    float pk, pm;
                    split_power(u[k2], v[R2 - 1 - k2], __ldg(a.t.post + k), pk, pm);
                    pw[k * 32 + lane] = pk;
                    pw[m * 32 + lane] = pm;
                }
            } else {
                // u: bins R1 k2 pair with N2 - R1 k2 = R1 (R2 - k2); k2 = 0 is DC/Nyquist, k2 = R2/2 self-paired
                {
                    const float dc = u[0].re + u[0].im, ny = u[0].re - u[0].im;
                    pw[0 * 32 + lane] = 4.0f * dc * dc;
                    pw[N2 * 32 + lane] = 4.0f * ny * ny;
                    const cplx z = u[R2 / 2];   // bin N2/2: X = conj(Z)
                    pw[(N2 / 2) * 32 + lane] = 4.0f * fmaf(z.re, z.re, z.im * z.im);
                }
#pragma unroll
                for (int k2 = 1; k2 < R2 / 2; ++k2) {
                    const int k = R1 * k2;
                    float pk, pm;
                    split_power(u[k2], u[R2 - k2], __ldg(a.t.post + k), pk, pm);
                    pw[k * 32 + lane] = pk;
                    pw[(N2 - k) * 32 + lane] = pm;
                }
                // v: bins R1/2 + R1 k2 pair with N2 - that = R1/2 + R1 (R2 - 1 - k2)
#pragma unroll
                for (int k2 = 0; k2 < R2 / 2; ++k2) {
                    const int k = R1 / 2 + R1 * k2;
                    float pk, pm;
                    split_power(v[k2], v[R2 - 1 - k2], __ldg(a.t.post + k), pk, pm);
                    pw[k * 32 + lane] = pk;
                    pw[(N2 - k) * 32 + lane] = pm;
                }
            }
        }
    }
    __syncthreads();

    // ---- S3: sparse mel.  Segment j = [bins[j], bins[j+1]) feeds filter j (rise) and j-1 (fall).
    const int n_seg = a.n_mel + 1;
    float *er = ws_f, *ef = ws_f + n_seg * 32, *lg = ws_f + 2 * n_seg * 32;
    {
        const float scale = 1.0f / (4.0f * static_cast<float>(NFFT));
        const int j0 = __ldg(a.t.seg_lo + warp), j1 = __ldg(a.t.seg_lo + warp + 1);
        for (int j = j0; j < j1; ++j) {
            const int k0 = __ldg(a.t.bins + j), k1 = __ldg(a.t.bins + j + 1);
            float r = 0.0f, f = 0.0f;
            for (int k = k0; k < k1; ++k) {
                const float p = pw[k * 32 + lane];
                r = fmaf(__ldg(a.t.rise + k), p, r);
                f = fmaf(__ldg(a.t.fall + k), p, f);
            }
            er[j * 32 + lane] = r * scale;
            ef[j * 32 + lane] = f * scale;
        }
    }
    __syncthreads();

    // ---- S4: log mel energies.
    const bool live = lane < tile.n_frames;
    for (int m = warp; m < a.n_mel; m += kWarps) {
        const float e = er[m * 32 + lane] + ef[(m + 1) * 32 + lane];
        const float l = logf(fmaxf(e, a.log_floor));
        if (a.logmel) {
            if (live) a.out[(tile.out_row + lane) * a.n_mel + m] = l;
        } else {
            lg[m * 32 + lane] = l;
        }
    }
    if (a.logmel) return;
    __syncthreads();

    // ---- S5: DCT-II rows (lifter folded in).
    for (int k = warp; k < a.n_cep; k += kWarps) {
        const float *d = a.t.dct + k * a.n_mel;
        float c = 0.0f;
        for (int m = 0; m < a.n_mel; ++m) c = fmaf(__ldg(d + m), lg[m * 32 + lane], c);
        if (live) a.out[(tile.out_row + lane) * a.n_cep + k] = c;
    }
}

struct Geometry { int nfft, r1, r2; const char *name; };
constexpr Geometry kGeoms[] = {
    {512, 16, 16, "fused_rt_tile32_r16x16_n512"},
    {256, 8, 16, "fused_rt_tile32_r8x16_n256"},
};

size_t staged_words_for(const mfcc_params &p)
{
    // staged tile: reads reach i = 31*hop + nfft - 1; P: (nfft/2+1) * 32
    const size_t hops = (31ull * p.hop_len + p.nfft) / p.hop_len + 1;
    const size_t staged = hops * (p.hop_len + kPad);
    const size_t pw = (static_cast<size_t>(p.nfft) / 2 + 1) * 32;
    size_t w = staged > pw ? staged : pw;
    return (w + 3) / 4 * 4;  // keep the workspace 16-byte aligned
}

size_t smem_bytes_for(const mfcc_params &p)
{
    return sizeof(float) * (staged_words_for(p) + static_cast<size_t>(p.nfft) * 32);
}

}  // namespace

struct FusedKernel {
    Geometry g;
};

namespace {
const FusedKernel kKernels[] = {{kGeoms[0]}, {kGeoms[1]}};
}

const FusedKernel *find_fused(const mfcc_params &p)
{
    for (const auto &k : kKernels) {
        if (k.g.nfft != p.nfft) continue;
        if (p.hop_len % 2 != 0 || p.hop_len < 2) return nullptr;  // float2 pairs must not straddle a hop block
        if (smem_bytes_for(p) > 227 * 1024) return nullptr;
        // S3-S5 scratch (er | ef | lg) lives in the workspace
        if (static_cast<size_t>(3 * (p.n_mel + 1)) * 32 > static_cast<size_t>(p.nfft) * 32) return nullptr;
        return &k;
    }
    return nullptr;
}

const char *fused_name(const FusedKernel *k) { return k ? k->g.name : ""; }

int fused_prepare(mfcc_plan *plan)
{
    const mfcc_params &p = plan->p;
    const FusedKernel *fk = plan->fused;
    if (fk == nullptr) return MFCC_ENOTSUP;
    const int N = p.nfft, N2 = N / 2, R1 = fk->g.r1, R2 = fk->g.r2, nb = N2 + 1, M = p.n_mel;
    const HostTables &h = plan->host;

    std::vector<float2> win2(N2), tw(static_cast<size_t>(R2) * R1), post(N2 + 1);
    for (int n = 0; n < N2; ++n) {
        const float a0 = 2 * n < p.frame_len ? h.window[2 * n] : 0.0f;
        const float a1 = 2 * n + 1 < p.frame_len ? h.window[2 * n + 1] : 0.0f;
        win2[n] = make_float2(a0, a1);
    }
    for (int n2 = 0; n2 < R2; ++n2)
        for (int k1 = 0; k1 < R1; ++k1) {
            const double ang = -2.0 * M_PI * static_cast<double>(n2) * k1 / N2;
            tw[static_cast<size_t>(n2) * R1 + k1] =
                make_float2(static_cast<float>(std::cos(ang)), static_cast<float>(std::sin(ang)));
        }
    for (int k = 0; k <= N2; ++k) {
        const double ang = -2.0 * M_PI * k / N;
        post[k] = make_float2(static_cast<float>(std::cos(ang)), static_cast<float>(std::sin(ang)));
    }
    // Balanced contiguous split of the M+1 segments over the warps by bin count.
    std::vector<int32_t> seg_lo(kWarps + 1, 0);
    {
        const int total = h.mel_bins[M + 1] - h.mel_bins[0];
        int j = 0;
        for (int w = 0; w < kWarps; ++w) {
            seg_lo[w] = j;
            const double target = static_cast<double>(total) * (w + 1) / kWarps;
            while (j < M + 1 && (h.mel_bins[j + 1] - h.mel_bins[0]) <= target + 1e-9) ++j;
            if (w == kWarps - 1) j = M + 1;
        }
        seg_lo[kWarps] = M + 1;
    }

    auto up = [](size_t v) { return (v + 255) / 256 * 256; };
    size_t off = 0;
    const size_t o_win = off;  off = up(off + sizeof(float2) * win2.size());
    const size_t o_tw = off;   off = up(off + sizeof(float2) * tw.size());
    const size_t o_post = off; off = up(off + sizeof(float2) * post.size());
    const size_t o_rise = off; off = up(off + sizeof(float) * nb);
    const size_t o_fall = off; off = up(off + sizeof(float) * nb);
    const size_t o_bins = off; off = up(off + sizeof(int32_t) * (M + 2));
    const size_t o_seg = off;  off = up(off + sizeof(int32_t) * (kWarps + 1));
    const size_t o_dct = off;  off = up(off + sizeof(float) * h.dct.size());
    std::vector<char> blob(off, 0);
    std::memcpy(blob.data() + o_win, win2.data(), sizeof(float2) * win2.size());
    std::memcpy(blob.data() + o_tw, tw.data(), sizeof(float2) * tw.size());
    std::memcpy(blob.data() + o_post, post.data(), sizeof(float2) * post.size());
    std::memcpy(blob.data() + o_rise, h.rise.data(), sizeof(float) * nb);
    std::memcpy(blob.data() + o_fall, h.fall.data(), sizeof(float) * nb);
    std::memcpy(blob.data() + o_bins, h.mel_bins.data(), sizeof(int32_t) * (M + 2));
    std::memcpy(blob.data() + o_seg, seg_lo.data(), sizeof(int32_t) * (kWarps + 1));
    std::memcpy(blob.data() + o_dct, h.dct.data(), sizeof(float) * h.dct.size());

    struct { void *blob; FusedTables t; } st{nullptr, {}};
    if (cudaMalloc(&st.blob, off) != cudaSuccess) { cudaGetLastError(); return MFCC_ENOMEM; }
    if (cudaMemcpy(st.blob, blob.data(), off, cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaFree(st.blob);
        return MFCC_ECUDA;
    }
    char *b = static_cast<char *>(st.blob);
    st.t.win2 = reinterpret_cast<const float2 *>(b + o_win);
    st.t.tw = reinterpret_cast<const float2 *>(b + o_tw);
    st.t.post = reinterpret_cast<const float2 *>(b + o_post);
    st.t.rise = reinterpret_cast<const float *>(b + o_rise);
    st.t.fall = reinterpret_cast<const float *>(b + o_fall);
    st.t.bins = reinterpret_cast<const int32_t *>(b + o_bins);
    st.t.seg_lo = reinterpret_cast<const int32_t *>(b + o_seg);
    st.t.dct = reinterpret_cast<const float *>(b + o_dct);
    plan->fused_blob = st.blob;
    plan->fused_tables = new FusedTables(st.t);
    return MFCC_OK;
}

void fused_release(mfcc_plan *plan)
{
    if (plan->fused_blob) cudaFree(plan->fused_blob);
    delete static_cast<FusedTables *>(plan->fused_tables);
    plan->fused_blob = nullptr;
    plan->fused_tables = nullptr;
}

template <typename PcmT, int R1, int R2>
static int launch_geom(const mfcc_plan *plan, const Tile *d_tiles, int64_t n_tiles, const PcmT *d_pcm,
                       float *d_out, cudaStream_t stream)
{
    const mfcc_params &p = plan->p;
    const size_t smem = smem_bytes_for(p);
    static thread_local const void *configured = nullptr;
    auto kern = fused_rt_kernel<PcmT, R1, R2>;
    if (configured != reinterpret_cast<const void *>(kern)) {
        // idempotent; cheap enough to redo per thread
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
            cudaGetLastError();
            return MFCC_ECUDA;
        }
        configured = reinterpret_cast<const void *>(kern);
    }
    FusedArgs a;
    a.out = d_out;
    a.t = *static_cast<const FusedTables *>(plan->fused_tables);
    a.frame_len = p.frame_len;
    a.hop = p.hop_len;
    a.n_mel = p.n_mel;
    a.n_cep = p.n_cep;
    a.logmel = p.output == MFCC_OUT_LOGMEL;
    a.preemph = p.preemph;
    a.log_floor = p.log_floor;
    a.staged_words = static_cast<int>(staged_words_for(p));
    int64_t done = 0;
    while (done < n_tiles) {
        const int64_t n = n_tiles - done > (1 << 30) ? (1 << 30) : n_tiles - done;
        a.tiles = d_tiles + done;
        kern<<<static_cast<unsigned>(n), kThreads, smem, stream>>>(d_pcm, a);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        done += n;
    }
    return cudaGetLastError() == cudaSuccess ? MFCC_OK : MFCC_ECUDA;
}

template <typename PcmT>
int launch_fused(const mfcc_plan *plan, const Tile *d_tiles, int64_t n_tiles, const PcmT *d_pcm, int64_t pcm_len,
                 float *d_out, cudaStream_t stream)
{
    if (n_tiles <= 0) return MFCC_OK;
    if (plan->sp_state != nullptr) return sp_launch<PcmT>(plan, d_tiles, n_tiles, d_pcm, pcm_len, d_out, stream);
    if (plan->wide_state != nullptr) return wide_launch<PcmT>(plan, d_tiles, n_tiles, d_pcm, pcm_len, d_out, stream);
    if (plan->ct_state != nullptr) return ct_launch<PcmT>(plan, d_tiles, n_tiles, d_pcm, pcm_len, d_out, stream);
    if (plan->fused == nullptr || plan->fused_tables == nullptr) return MFCC_ENOTSUP;
    switch (plan->fused->g.nfft) {
        case 512: return launch_geom<PcmT, 16, 16>(plan, d_tiles, n_tiles, d_pcm, d_out, stream);
        case 256: return launch_geom<PcmT, 8, 16>(plan, d_tiles, n_tiles, d_pcm, d_out, stream);
        default: return MFCC_ENOTSUP;
    }
}

template int launch_fused<int16_t>(const mfcc_plan *, const Tile *, int64_t, const int16_t *, int64_t, float *,
                                   cudaStream_t);
template int launch_fused<float>(const mfcc_plan *, const Tile *, int64_t, const float *, int64_t, float *,
                                 cudaStream_t);

}  // namespace mfcc
