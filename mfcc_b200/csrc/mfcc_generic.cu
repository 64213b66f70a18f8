// mfcc_generic.cu — the any-geometry kernel and the small post-processing kernels.
//
// generic_radix2: one CTA per tile (<= 32 consecutive frames of one utterance),
// one frame at a time through shared memory: framing + pre-emphasis + window,
// in-place radix-2 complex FFT, power, dense mel rows in ascending-bin order,
// log, DCT.  It exists so that EVERY valid mfcc_params has a CUDA path (there
// is no CPU fallback) and as an independent cross-check of the fused kernels;
// its summation orders are the same as the oracle's.  The fused kernels in
// mfcc_fused.cu are the performance path.
//
// No reference code corresponds to any of this ("Ref file:line
// = none"); stage definitions are DESIGN.md "Spec".
#include <cuda_runtime.h>

#include "mfcc_host.h"

namespace mfcc {

std::atomic<uint64_t> g_launches{0};

namespace {

constexpr int kGenericThreads = 128;

__device__ __forceinline__ int ulaw_expand(unsigned b);
__device__ __forceinline__ int alaw_expand(unsigned b);
__device__ __forceinline__ float load_sample(const int16_t *p, int64_t i, int) { return static_cast<float>(p[i]); }
__device__ __forceinline__ float load_sample(const float *p, int64_t i, int) { return p[i]; }
__device__ __forceinline__ float load_sample(const uint8_t *p, int64_t i, int alaw)   // G.711 code -> linear PCM
{
    return static_cast<float>(alaw ? alaw_expand(p[i]) : ulaw_expand(p[i]));
}

template <typename PcmT>
__global__ void __launch_bounds__(kGenericThreads)
generic_radix2_kernel(const Tile *__restrict__ tiles, const PcmT *__restrict__ pcm,
                      float *__restrict__ out, DevTables tb, int frame_len, int hop, int nfft,
                      int log2n, int n_mel, int n_cep, int logmel, int energy, int alaw, float preemph, float log_floor)
{
    extern __shared__ float smem[];
    float *re = smem;                 // [nfft]
    float *im = re + nfft;            // [nfft]
    float *pw = im + nfft;            // [nfft/2 + 1]
    float *lg = pw + (nfft / 2 + 1);  // [n_mel] log band energies, then the log frame energy
    const int nbins = nfft / 2 + 1;
    const int base_dim = logmel ? n_mel : n_cep;
    const int out_dim = base_dim + (energy == MFCC_ENERGY_APPEND ? 1 : 0);
    const Tile tile = tiles[blockIdx.x];
    const float inv_n = 1.0f / static_cast<float>(nfft);

    for (int f = 0; f < tile.n_frames; ++f) {
        const int64_t s0 = tile.first_sample + static_cast<int64_t>(f) * hop;
        // Framing + pre-emphasis + window, stored bit-reversed for the DIT FFT.
        for (int i = threadIdx.x; i < nfft; i += kGenericThreads) {
            float v = 0.0f;
            const int64_t s = s0 + i;
            if (i < frame_len && s < tile.utt_end) {
                const float x0 = load_sample(pcm, s, alaw);
                const float x1 = s > tile.utt_begin ? load_sample(pcm, s - 1, alaw) : 0.0f;
                v = __fmul_rn(__fsub_rn(x0, __fmul_rn(preemph, x1)), tb.window[i]);
            }
            const int r = static_cast<int>(__brev(static_cast<unsigned>(i)) >> (32 - log2n));
            re[r] = v;
            im[r] = 0.0f;
        }
        __syncthreads();
        for (int len = 2; len <= nfft; len <<= 1) {
            const int half = len >> 1, step = nfft / len;
            for (int b = threadIdx.x; b < nfft / 2; b += kGenericThreads) {
                const int k = b & (half - 1);
                const int a = ((b - k) << 1) + k, c = a + half;
                const float2 w = tb.twiddle[k * step];
                const float xr = __fsub_rn(__fmul_rn(re[c], w.x), __fmul_rn(im[c], w.y));
                const float xi = __fadd_rn(__fmul_rn(re[c], w.y), __fmul_rn(im[c], w.x));
                const float ar = re[a], ai = im[a];
                re[c] = ar - xr; im[c] = ai - xi;
                re[a] = ar + xr; im[a] = ai + xi;
            }
            __syncthreads();
        }
        for (int k = threadIdx.x; k < nbins; k += kGenericThreads)
            pw[k] = __fmul_rn(__fadd_rn(__fmul_rn(re[k], re[k]), __fmul_rn(im[k], im[k])), inv_n);
        __syncthreads();
        for (int m = threadIdx.x; m < n_mel + (energy != MFCC_ENERGY_NONE ? 1 : 0); m += kGenericThreads) {
            if (m == n_mel) {   // frame energy: the one-sided power spectrum summed in ascending-bin order
                float e = 0.0f;
                for (int k = 0; k < nbins; ++k) e = __fadd_rn(e, pw[k]);
                lg[m] = logf(fmaxf(e, log_floor));
                continue;
            }
            const float *w = tb.mel_w + static_cast<size_t>(m) * nbins;
            int k1 = tb.mel_bins[m + 2];
            if (k1 > nbins - 1) k1 = nbins - 1;
            float e = 0.0f;
            for (int k = tb.mel_bins[m]; k <= k1; ++k) e = __fadd_rn(e, __fmul_rn(w[k], pw[k]));
            lg[m] = logf(fmaxf(e, log_floor));
        }
        __syncthreads();
        float *o = out + (tile.out_row + f) * out_dim;
        for (int k = threadIdx.x; k < out_dim; k += kGenericThreads) {
            if (k == base_dim || (k == 0 && energy == MFCC_ENERGY_REPLACE_C0)) {
                o[k] = lg[n_mel];
            } else if (logmel) {
                o[k] = lg[k];
            } else {
                const float *d = tb.dct + static_cast<size_t>(k) * n_mel;
                float c = 0.0f;
                for (int m = 0; m < n_mel; ++m) c = __fadd_rn(c, __fmul_rn(d[m], lg[m]));
                o[k] = c;
            }
        }
        __syncthreads();
    }
}

// ---- §8(f) rank 3: G.711 expansion, 16 codes per thread (uint4 in, 2 x uint4 out) ----
__device__ __forceinline__ int ulaw_expand(unsigned b)
{
    const unsigned u = (~b) & 0xFFu;
    const int mag = static_cast<int>((((u & 0x0Fu) << 3) + 0x84u) << ((u >> 4) & 7u)) - 0x84;
    return (u & 0x80u) ? -mag : mag;
}
__device__ __forceinline__ int alaw_expand(unsigned b)
{
    const unsigned a = b ^ 0x55u;
    const unsigned seg = (a >> 4) & 7u, man = a & 0x0Fu;
    const int mag = seg == 0 ? static_cast<int>((man << 4) + 8u)
                             : static_cast<int>(((man << 4) + 0x108u) << (seg - 1));
    return (a & 0x80u) ? mag : -mag;
}

__global__ void __launch_bounds__(256)
g711_kernel(const uint8_t *__restrict__ src, int64_t n, int alaw, int16_t *__restrict__ dst)
{
    const int64_t i0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 16;
    if (i0 >= n) return;
    const bool vec = (i0 + 16 <= n) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
    if (vec) {
        const uint4 in = *reinterpret_cast<const uint4 *>(src + i0);
        const unsigned w[4] = {in.x, in.y, in.z, in.w};
        unsigned o[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const unsigned b0 = (w[j] >> (16 * h)) & 0xFFu, b1 = (w[j] >> (16 * h + 8)) & 0xFFu;
                const int s0 = alaw ? alaw_expand(b0) : ulaw_expand(b0);
                const int s1 = alaw ? alaw_expand(b1) : ulaw_expand(b1);
                o[2 * j + h] = (static_cast<unsigned>(s0) & 0xFFFFu) | (static_cast<unsigned>(s1) << 16);
            }
        }
        uint4 *d = reinterpret_cast<uint4 *>(dst + i0);
        d[0] = make_uint4(o[0], o[1], o[2], o[3]);
        d[1] = make_uint4(o[4], o[5], o[6], o[7]);
    } else {
        for (int64_t i = i0; i < n && i < i0 + 16; ++i)
            dst[i] = static_cast<int16_t>(alaw ? alaw_expand(src[i]) : ulaw_expand(src[i]));
    }
}

}  // namespace

template <typename PcmT>
int launch_generic(const mfcc_plan *plan, const Tile *d_tiles, int64_t n_tiles, const PcmT *d_pcm,
                   float *d_out, int alaw, cudaStream_t stream)
{
    if (n_tiles <= 0) return MFCC_OK;
    const mfcc_params &p = plan->p;
    int log2n = 0;
    while ((1 << log2n) < p.nfft) ++log2n;
    const size_t smem = sizeof(float) * (2 * static_cast<size_t>(p.nfft) + p.nfft / 2 + 1 + p.n_mel + 1);
    int64_t done = 0;
    while (done < n_tiles) {  // gridDim.x limit is 2^31-1; chunk anyway
        const int64_t n = n_tiles - done > (1 << 30) ? (1 << 30) : n_tiles - done;
        generic_radix2_kernel<PcmT><<<static_cast<unsigned>(n), kGenericThreads, smem, stream>>>(
            d_tiles + done, d_pcm, d_out, plan->dev, p.frame_len, p.hop_len, p.nfft, log2n, p.n_mel,
            p.n_cep, p.output == MFCC_OUT_LOGMEL, p.energy, alaw, p.preemph, p.log_floor);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        done += n;
    }
    return cudaGetLastError() == cudaSuccess ? MFCC_OK : MFCC_ECUDA;
}

template int launch_generic<int16_t>(const mfcc_plan *, const Tile *, int64_t, const int16_t *, float *, int,
                                     cudaStream_t);
template int launch_generic<float>(const mfcc_plan *, const Tile *, int64_t, const float *, float *, int,
                                   cudaStream_t);
template int launch_generic<uint8_t>(const mfcc_plan *, const Tile *, int64_t, const uint8_t *, float *, int,
                                     cudaStream_t);

int launch_g711(const uint8_t *d_src, int64_t n, int alaw, int16_t *d_dst, cudaStream_t s)
{
    if (n <= 0) return MFCC_OK;
    const int64_t threads = (n + 15) / 16, blocks = (threads + 255) / 256;
    g711_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(d_src, n, alaw, d_dst);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError() == cudaSuccess ? MFCC_OK : MFCC_ECUDA;
}

}  // namespace mfcc
