// mfcc_api.cu — the C ABI of libmfcc_b200.so (include/mfcc_b200.h) and the thin
// host layer behind it: plan = validated parameters + device tables; batch =
// offsets -> frame rows -> tile table; compute = kernel launches only.
//
// Convention kept from the reference: int return, 0 ok,
// negative on failure (src/mfcc/main.c:72-76); caller-owned buffers filled in
// place (src/mfcc/main.c:64-66).  Dropped: exit()/assert() on error
// (src/mfcc/main.c:124-127, src/mfcc/codegen.c:166) — a library reports.
#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <new>

#include "mfcc_host.h"

using mfcc::Tile;

namespace {

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// PcmT = int16_t, float or uint8_t (G.711 codes, `alaw` selects the law).  The 2048-point kernel has no G.711 entry
// (telephony codes at 48 kHz do not occur): such a plan takes the generic kernel for them.
template <typename PcmT>
int compute_batch_impl(const mfcc_plan *plan, const mfcc_batch *batch, const PcmT *d_pcm, float *d_out,
                       int64_t tile0, int64_t n_tiles, cudaStream_t stream, int alaw = 0)
{
    if (plan->sp_state != nullptr)
        return mfcc::sp_launch<PcmT>(plan, batch->d_tiles + tile0, n_tiles, d_pcm, batch->total_samples, d_out, alaw, stream);
    if constexpr (sizeof(PcmT) != 1) {
        if (plan->wide_state != nullptr)
            return mfcc::wide_launch<PcmT>(plan, batch->d_tiles + tile0, n_tiles, d_pcm, batch->total_samples, d_out, stream);
    }
    return mfcc::launch_generic<PcmT>(plan, batch->d_tiles + tile0, n_tiles, d_pcm, d_out, alaw, stream);
}

// A batch's tile table (frame starts, kTileInside flags) was computed for one framing: only plans with that framing
// on that device may use it.
bool batch_fits(const mfcc_plan *plan, const mfcc_batch *b)
{
    return plan != nullptr && b != nullptr && b->device == plan->device && b->out_dim == plan->host.out_dim &&
           b->frame_len == plan->p.frame_len && b->hop_len == plan->p.hop_len && b->pad_mode == plan->p.pad_mode;
}

int grow(void **ptr, size_t *have, size_t need)
{
    if (*have >= need) return MFCC_OK;
    if (*ptr) cudaFree(*ptr);
    *ptr = nullptr;
    *have = 0;
    const size_t want = align_up(need + need / 8, 1 << 20);
    if (cudaMalloc(ptr, want) != cudaSuccess) { cudaGetLastError(); return MFCC_ENOMEM; }
    *have = want;
    return MFCC_OK;
}

}  // namespace

extern "C" {

int mfcc_plan_create(const mfcc_params *p, int32_t device, int32_t kernel, mfcc_plan **out)
{
    if (out == nullptr) return MFCC_EINVAL;
    *out = nullptr;
    if (mfcc::validate_params(p) != MFCC_OK) return MFCC_EINVAL;
    if (kernel < MFCC_KERNEL_AUTO || kernel > MFCC_KERNEL_FUSED) return MFCC_EINVAL;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) { cudaGetLastError(); return MFCC_ECUDA; }
    if (device < 0 && cudaGetDevice(&device) != cudaSuccess) return MFCC_ECUDA;
    if (device >= count) return MFCC_EINVAL;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return MFCC_ECUDA;
    if (prop.major != 10) return MFCC_ECUDA;  // the only code in this library is sm_100a SASS

    mfcc_plan *plan = new (std::nothrow) mfcc_plan();
    if (plan == nullptr) return MFCC_ENOMEM;
    plan->p = *p;
    plan->device = device;
    plan->sm_count = prop.multiProcessorCount;
    int rc = mfcc::build_tables(plan->p, plan->host);
    if (rc != MFCC_OK) { delete plan; return rc; }

    // Kernel choice: the fused tile kernel of the geometry (512- / 256-point: mfcc_fused_sp.cu, 2048-point:
    // mfcc_fused_wide.cu) when the plan has one, else the generic kernel.  MFCC_KERNEL_FUSED insists.
    const bool want_fused = kernel != MFCC_KERNEL_GENERIC;
    const char *fused_name = nullptr;
    if (want_fused) {
        fused_name = mfcc::sp_match(plan->p, plan->host);
        if (fused_name == nullptr) fused_name = mfcc::wide_match(plan->p, plan->host);
    }
    if (kernel == MFCC_KERNEL_FUSED && fused_name == nullptr) { delete plan; return MFCC_ENOTSUP; }
    plan->kernel = fused_name != nullptr ? MFCC_KERNEL_FUSED : MFCC_KERNEL_GENERIC;

    DeviceGuard guard(device);
    if (!guard.ok) { delete plan; return MFCC_ECUDA; }

    // One device blob for all generic-path tables.
    const mfcc::HostTables &h = plan->host;
    const int N = p->nfft, nb = h.nbins, M = p->n_mel;
    size_t off = 0;
    const size_t o_window = off;  off = align_up(off + sizeof(float) * N, 256);
    const size_t o_twiddle = off; off = align_up(off + sizeof(float2) * (N / 2), 256);
    const size_t o_melw = off;    off = align_up(off + sizeof(float) * M * nb, 256);
    const size_t o_bins = off;    off = align_up(off + sizeof(int32_t) * (M + 2), 256);
    const size_t o_dct = off;     off = align_up(off + sizeof(float) * h.dct.size(), 256);
    const size_t o_rise = off;    off = align_up(off + sizeof(float) * nb, 256);
    const size_t o_fall = off;    off = align_up(off + sizeof(float) * nb, 256);
    std::vector<char> blob(off, 0);
    std::memcpy(blob.data() + o_window, h.window.data(), sizeof(float) * h.window.size());
    float2 *tw = reinterpret_cast<float2 *>(blob.data() + o_twiddle);
    for (int k = 0; k < N / 2; ++k) tw[k] = make_float2(h.tw_re[k], h.tw_im[k]);
    std::memcpy(blob.data() + o_melw, h.mel_w.data(), sizeof(float) * h.mel_w.size());
    std::memcpy(blob.data() + o_bins, h.mel_bins.data(), sizeof(int32_t) * h.mel_bins.size());
    std::memcpy(blob.data() + o_dct, h.dct.data(), sizeof(float) * h.dct.size());
    std::memcpy(blob.data() + o_rise, h.rise.data(), sizeof(float) * nb);
    std::memcpy(blob.data() + o_fall, h.fall.data(), sizeof(float) * nb);
    if (cudaMalloc(&plan->dev_blob, off) != cudaSuccess) { cudaGetLastError(); delete plan; return MFCC_ENOMEM; }
    if (cudaMemcpy(plan->dev_blob, blob.data(), off, cudaMemcpyHostToDevice) != cudaSuccess) {
        mfcc_plan_destroy(plan);
        return MFCC_ECUDA;
    }
    char *base = static_cast<char *>(plan->dev_blob);
    plan->dev.window = reinterpret_cast<const float *>(base + o_window);
    plan->dev.twiddle = reinterpret_cast<const float2 *>(base + o_twiddle);
    plan->dev.mel_w = reinterpret_cast<const float *>(base + o_melw);
    plan->dev.mel_bins = reinterpret_cast<const int32_t *>(base + o_bins);
    plan->dev.dct = reinterpret_cast<const float *>(base + o_dct);
    plan->dev.rise = reinterpret_cast<const float *>(base + o_rise);
    plan->dev.fall = reinterpret_cast<const float *>(base + o_fall);

    if (plan->kernel == MFCC_KERNEL_FUSED) {
        rc = mfcc::sp_prepare(plan);
        if (rc == MFCC_ENOTSUP) rc = mfcc::wide_prepare(plan);
        if (rc == MFCC_ENOTSUP && kernel == MFCC_KERNEL_AUTO) { plan->kernel = MFCC_KERNEL_GENERIC; rc = MFCC_OK; }
        if (rc != MFCC_OK) { mfcc_plan_destroy(plan); return rc; }
    }
    plan->kernel_name = plan->kernel == MFCC_KERNEL_FUSED ? fused_name : "generic_radix2";
    *out = plan;
    return MFCC_OK;
}

void mfcc_plan_destroy(mfcc_plan *plan)
{
    if (plan == nullptr) return;
    DeviceGuard guard(plan->device);
    for (auto &s : plan->streams)
        if (s) cudaStreamDestroy(s);
    if (plan->h2d_pcm) cudaFree(plan->h2d_pcm);
    if (plan->d2h_out) cudaFree(plan->d2h_out);
    if (plan->d_tiles) cudaFree(plan->d_tiles);
    if (plan->h_tiles) cudaFreeHost(plan->h_tiles);
    if (plan->d_post_out) cudaFree(plan->d_post_out);
    if (plan->d_post_chunks) cudaFree(plan->d_post_chunks);
    if (plan->h_post_chunks) cudaFreeHost(plan->h_post_chunks);
    if (plan->d_post_partial) cudaFree(plan->d_post_partial);
    if (plan->d_post_stats) cudaFree(plan->d_post_stats);
    if (plan->h_stage_pcm) cudaFreeHost(plan->h_stage_pcm);
    if (plan->h_stage_out) cudaFreeHost(plan->h_stage_out);
    if (plan->tiles_ready) cudaEventDestroy(plan->tiles_ready);
    for (cudaEvent_t e : plan->chunk_ready) cudaEventDestroy(e);
    if (plan->dev_blob) cudaFree(plan->dev_blob);
    mfcc::sp_release(plan);
    mfcc::wide_release(plan);
    delete plan;
}

int mfcc_plan_params(const mfcc_plan *plan, mfcc_params *out)
{
    if (plan == nullptr || out == nullptr) return MFCC_EINVAL;
    *out = plan->p;
    return MFCC_OK;
}

const char *mfcc_plan_kernel_name(const mfcc_plan *plan)
{
    return plan ? plan->kernel_name.c_str() : "";
}

int64_t mfcc_plan_window(const mfcc_plan *plan, float *dst)
{
    if (plan == nullptr) return MFCC_EINVAL;
    if (dst) std::memcpy(dst, plan->host.window.data(), sizeof(float) * plan->host.window.size());
    return static_cast<int64_t>(plan->host.window.size());
}

int64_t mfcc_plan_mel_bins(const mfcc_plan *plan, int32_t *dst)
{
    if (plan == nullptr) return MFCC_EINVAL;
    if (dst) std::memcpy(dst, plan->host.mel_bins.data(), sizeof(int32_t) * plan->host.mel_bins.size());
    return static_cast<int64_t>(plan->host.mel_bins.size());
}

int64_t mfcc_plan_mel_weights(const mfcc_plan *plan, float *dst)
{
    if (plan == nullptr) return MFCC_EINVAL;
    if (dst) std::memcpy(dst, plan->host.mel_w.data(), sizeof(float) * plan->host.mel_w.size());
    return static_cast<int64_t>(plan->host.mel_w.size());
}

int64_t mfcc_plan_dct(const mfcc_plan *plan, float *dst)
{
    if (plan == nullptr) return MFCC_EINVAL;
    if (dst) std::memcpy(dst, plan->host.dct.data(), sizeof(float) * plan->host.dct.size());
    return static_cast<int64_t>(plan->host.dct.size());
}

// Host half of a batch: offsets -> frame rows -> tile table.  No CUDA calls.
// frames of an utterance of n samples under a plan's (already validated) framing: mfcc_num_frames without the re-validation,
// for the per-utterance loops of the host path (16,384 utterances per call in the ragged telephony batch)
static inline int64_t frames_of(const mfcc_params &p, int64_t n)
{
    const int64_t L = p.frame_len, H = p.hop_len;
    if (p.pad_mode == MFCC_PAD_NONE) return n < L ? 0 : 1 + (n - L) / H;
    return n <= 0 ? 0 : (n <= L ? 1 : 1 + (n - L + H - 1) / H);
}

// Host half of a batch, first step: offsets -> frame rows -> first tile index per utterance.  No tiles yet, no CUDA calls.
// h_lead (may be null): h_lead[u] != 0 marks utterance u as a PIECE of a longer recording whose first sample is history — it
// only serves as the pre-emphasis predecessor of the piece's first frame; the frames start one sample later.
static int batch_build_meta(const mfcc_plan *plan, const int64_t *h_offsets, int64_t n_utts, mfcc_batch **out,
                            const uint8_t *h_lead = nullptr)
{
    if (out == nullptr) return MFCC_EINVAL;
    *out = nullptr;
    if (plan == nullptr || n_utts < 0 || (n_utts > 0 && h_offsets == nullptr)) return MFCC_EINVAL;
    mfcc_batch *b = new (std::nothrow) mfcc_batch();
    if (b == nullptr) return MFCC_ENOMEM;
    b->device = plan->device;
    b->n_utts = n_utts;
    b->out_dim = plan->host.out_dim;
    const mfcc_params &p = plan->p;
    b->frame_len = p.frame_len;
    b->hop_len = p.hop_len;
    b->pad_mode = p.pad_mode;
    try {
        b->offsets.assign(n_utts + 1, 0);
        b->frame_offsets.assign(n_utts + 1, 0);
        b->utt_first_tile.assign(n_utts + 1, 0);
        if (h_lead != nullptr) b->lead.assign(h_lead, h_lead + n_utts);
        if (n_utts > 0) std::memcpy(b->offsets.data(), h_offsets, sizeof(int64_t) * (n_utts + 1));
        for (int64_t u = 0; u < n_utts; ++u) {
            const int64_t begin = b->offsets[u], end = b->offsets[u + 1];
            if (begin < 0 || end < begin) { delete b; return MFCC_EINVAL; }
            const int64_t lead = (h_lead != nullptr && h_lead[u] != 0 && end > begin) ? 1 : 0;
            const int64_t nf = frames_of(p, end - begin - lead);
            b->frame_offsets[u + 1] = b->frame_offsets[u] + nf;
            b->utt_first_tile[u + 1] = b->utt_first_tile[u] + (nf + mfcc::kTileFrames - 1) / mfcc::kTileFrames;
        }
    } catch (const std::bad_alloc &) {
        delete b;
        return MFCC_ENOMEM;
    }
    b->total_frames = b->frame_offsets[n_utts];
    b->total_samples = n_utts > 0 ? b->offsets[n_utts] : 0;
    *out = b;
    return MFCC_OK;
}

// Second step: the tiles of utterances [u0, u1), written to dst[0 .. utt_first_tile[u1] - utt_first_tile[u0]).  The host path
// calls this per chunk of utterances, straight into its pinned staging buffer, while the earlier chunks are on the wire.
static void batch_build_tiles(const mfcc_params &p, const mfcc_batch &b, int64_t u0, int64_t u1, Tile *dst)
{
    for (int64_t u = u0; u < u1; ++u) {
        const int64_t begin = b.offsets[u], end = b.offsets[u + 1];
        const int64_t lead = (!b.lead.empty() && b.lead[u] != 0 && end > begin) ? 1 : 0;
        const int64_t nf = b.frame_offsets[u + 1] - b.frame_offsets[u];
        for (int64_t f = 0; f < nf; f += mfcc::kTileFrames) {
            Tile t;
            t.utt_begin = begin;
            t.utt_end = end;
            t.first_sample = begin + lead + f * p.hop_len;
            t.out_row = b.frame_offsets[u] + f;
            t.n_frames = static_cast<int32_t>(std::min<int64_t>(mfcc::kTileFrames, nf - f));
            t.flags = mfcc::tile_flags(p, t, b.total_samples);
            *dst++ = t;
        }
    }
}

// Both steps at once (mfcc_batch_create): the tiles land in b->tiles.
static int batch_build_host(const mfcc_plan *plan, const int64_t *h_offsets, int64_t n_utts, mfcc_batch **out,
                            const uint8_t *h_lead = nullptr)
{
    const int rc = batch_build_meta(plan, h_offsets, n_utts, out, h_lead);
    if (rc != MFCC_OK) return rc;
    mfcc_batch *b = *out;
    try {
        b->tiles.resize(static_cast<size_t>(b->utt_first_tile[n_utts]));
    } catch (const std::bad_alloc &) {
        delete b;
        *out = nullptr;
        return MFCC_ENOMEM;
    }
    batch_build_tiles(plan->p, *b, 0, n_utts, b->tiles.data());
    return MFCC_OK;
}

int mfcc_batch_create(const mfcc_plan *plan, const int64_t *h_offsets, int64_t n_utts, mfcc_batch **out)
{
    return mfcc_batch_create_lead(plan, h_offsets, nullptr, n_utts, out);
}

int mfcc_batch_create_lead(const mfcc_plan *plan, const int64_t *h_offsets, const uint8_t *h_lead, int64_t n_utts,
                           mfcc_batch **out)
{
    mfcc_batch *b = nullptr;
    const int rc = batch_build_host(plan, h_offsets, n_utts, &b, h_lead);
    if (out) *out = nullptr;
    if (rc != MFCC_OK) return rc;
    DeviceGuard guard(plan->device);
    if (!guard.ok) { delete b; return MFCC_ECUDA; }
    const size_t tb = sizeof(Tile) * std::max<size_t>(b->tiles.size(), 1);
    const size_t fb = sizeof(int64_t) * (n_utts + 1);
    if (cudaMalloc(&b->d_tiles, tb) != cudaSuccess || cudaMalloc(&b->d_frame_offsets, fb) != cudaSuccess) {
        cudaGetLastError();
        mfcc_batch_destroy(b);
        return MFCC_ENOMEM;
    }
    bool ok = true;
    if (!b->tiles.empty())
        ok = cudaMemcpy(b->d_tiles, b->tiles.data(), sizeof(Tile) * b->tiles.size(),
                        cudaMemcpyHostToDevice) == cudaSuccess;
    ok = ok && cudaMemcpy(b->d_frame_offsets, b->frame_offsets.data(), fb, cudaMemcpyHostToDevice) ==
                   cudaSuccess;
    if (!ok) { mfcc_batch_destroy(b); return MFCC_ECUDA; }
    // post-processing tables (mfcc_post_batch): chunk table + statistics scratch, a few MB at most
    mfcc::post_build_chunks(b->frame_offsets, b->out_dim, b->post_chunks, b->utt_first_post_chunk, &b->post_rows);
    if (!b->post_chunks.empty()) {
        const size_t nc = b->post_chunks.size(), dim = static_cast<size_t>(b->out_dim);
        const size_t nu = static_cast<size_t>(n_utts);
        if (cudaMalloc(&b->d_post_chunks, nc * sizeof(mfcc::PostChunk)) != cudaSuccess ||
            cudaMalloc(&b->d_post_partial, nc * dim * mfcc::kPostPartialBytes) != cudaSuccess ||
            cudaMalloc(&b->d_post_stats, nu * dim * sizeof(float4)) != cudaSuccess) {
            cudaGetLastError();
            mfcc_batch_destroy(b);
            return MFCC_ENOMEM;
        }
        if (cudaMemcpy(b->d_post_chunks, b->post_chunks.data(), nc * sizeof(mfcc::PostChunk), cudaMemcpyHostToDevice) != cudaSuccess) {
            mfcc_batch_destroy(b);
            return MFCC_ECUDA;
        }
    }
    *out = b;
    return MFCC_OK;
}

void mfcc_batch_destroy(mfcc_batch *b)
{
    if (b == nullptr) return;
    DeviceGuard guard(b->device);
    if (b->d_tiles && !b->tiles_borrowed) cudaFree(b->d_tiles);
    if (b->d_frame_offsets) cudaFree(b->d_frame_offsets);
    if (b->d_post_chunks) cudaFree(b->d_post_chunks);
    if (b->d_post_partial) cudaFree(b->d_post_partial);
    if (b->d_post_stats) cudaFree(b->d_post_stats);
    delete b;
}

int64_t mfcc_batch_total_frames(const mfcc_batch *b) { return b ? b->total_frames : MFCC_EINVAL; }
int64_t mfcc_batch_total_samples(const mfcc_batch *b) { return b ? b->total_samples : MFCC_EINVAL; }

int mfcc_batch_frame_offsets(const mfcc_batch *b, int64_t *dst)
{
    if (b == nullptr || dst == nullptr) return MFCC_EINVAL;
    std::memcpy(dst, b->frame_offsets.data(), sizeof(int64_t) * b->frame_offsets.size());
    return MFCC_OK;
}

int mfcc_compute_batch(const mfcc_plan *plan, const mfcc_batch *batch, const int16_t *d_pcm, float *d_out,
                       void *cuda_stream)
{
    if (!batch_fits(plan, batch)) return MFCC_EINVAL;
    if (batch->total_frames == 0) return MFCC_OK;
    if (d_pcm == nullptr || d_out == nullptr) return MFCC_EINVAL;
    DeviceGuard guard(plan->device);
    if (!guard.ok) return MFCC_ECUDA;
    return compute_batch_impl<int16_t>(plan, batch, d_pcm, d_out, 0, static_cast<int64_t>(batch->tiles.size()),
                                       static_cast<cudaStream_t>(cuda_stream));
}

int mfcc_compute_batch_f32(const mfcc_plan *plan, const mfcc_batch *batch, const float *d_pcm, float *d_out,
                           void *cuda_stream)
{
    if (!batch_fits(plan, batch)) return MFCC_EINVAL;
    if (batch->total_frames == 0) return MFCC_OK;
    if (d_pcm == nullptr || d_out == nullptr) return MFCC_EINVAL;
    DeviceGuard guard(plan->device);
    if (!guard.ok) return MFCC_ECUDA;
    return compute_batch_impl<float>(plan, batch, d_pcm, d_out, 0, static_cast<int64_t>(batch->tiles.size()),
                                     static_cast<cudaStream_t>(cuda_stream));
}

// End-to-end: chunks of whole utterances flow H2D -> kernel -> D2H on three
// streams so that the copy of chunk i+1 overlaps the kernel of chunk i and the
// read-back of chunk i-1.
// End to end with HOST buffers.  Order of work (on the plan's four streams):
//   1. validate the offsets and size everything (no CUDA work yet), grow the plan's device buffers if needed;
//   2. queue EVERY host->device PCM chunk, alternating between TWO copy streams, an event after each — the DMA
//      starts at once and two copies are always in flight (measured on B200: one copy stream reaches 46 GB/s,
//      concurrent copies 50 GB/s);
//   3. build the tile table on the host while the first chunks fly, upload it;
//   4. per chunk, on alternating compute streams: wait for the chunk's event, one kernel launch over the
//      chunk's tiles, device->host copy of its feature rows (the other DMA direction).  Result copies never sit
//      in front of an input copy in any stream — with H2D, kernel and D2H of a chunk in ONE stream the next
//      input copy of that stream waits for the result copy, which costs 7 % on the ragged 8 kHz batch.
// Chunks are 32 MiB of PCM, tapering geometrically over the last third of the batch so that the work left
// after the last H2D byte (one kernel + one D2H of the LAST chunk) is small.
}  // extern "C"

namespace {

// post != nullptr: every chunk of utterances also goes through the fused post-processing kernels (CMVN + regressions,
// mfcc_post.cu) on its compute stream and the STACKED rows are what travels back — the normalisation is per utterance and
// a chunk holds whole utterances, so nothing crosses chunks.
struct PostOpts { int cmvn, window, order; };

template <typename PcmT>
int compute_host_impl(mfcc_plan *plan, const PcmT *h_pcm, int alaw, const int64_t *h_offsets, int64_t n_utts,
                      float *h_out, int64_t *h_frame_offsets, const PostOpts *post = nullptr)
{
    if (plan == nullptr || n_utts < 0 || (n_utts > 0 && h_offsets == nullptr)) return MFCC_EINVAL;
    const mfcc_params &p = plan->p;
    const int od = plan->host.out_dim;
    // MFCC_TRACE_HOST=1: host-clock milestones of this call on stderr (where the time of the host path goes)
    static const bool trace = std::getenv("MFCC_TRACE_HOST") != nullptr;
    const auto t_entry = std::chrono::steady_clock::now();
    double t_mark[6] = {0, 0, 0, 0, 0, 0};
    auto mark = [&](int i) { if (trace) t_mark[i] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_entry).count(); };
    // 1. sizes: the host half of the batch without its tiles (offsets checked, frame rows, tile index per utterance)
    mfcc_batch *batch = nullptr;
    {
        const int rc0 = batch_build_meta(plan, h_offsets, n_utts, &batch);
        if (rc0 != MFCC_OK) return rc0;
    }
    struct BatchOwner { mfcc_batch *b; ~BatchOwner() { if (b) mfcc_batch_destroy(b); } } owner{batch};
    const int64_t total_frames = batch->total_frames, n_tiles = batch->utt_first_tile[n_utts];
    const int post_rows = post ? mfcc::post_rows_for(od) : 1;
    const int od_out = post ? od * (1 + post->order) : od;     // floats per row that travels back
    int64_t n_post_chunks = 0;
    if (post != nullptr)
        for (int64_t u = 0; u < n_utts; ++u)
            n_post_chunks += (batch->frame_offsets[u + 1] - batch->frame_offsets[u] + post_rows - 1) / post_rows;
    const int64_t total_samples = batch->total_samples;
    if (h_frame_offsets) mfcc_batch_frame_offsets(batch, h_frame_offsets);
    if (total_frames == 0) return MFCC_OK;
    if (h_pcm == nullptr || h_out == nullptr) return MFCC_EINVAL;

    std::lock_guard<std::mutex> lock(plan->host_mutex);
    DeviceGuard guard(plan->device);
    if (!guard.ok) return MFCC_ECUDA;
    const size_t tile_bytes = sizeof(Tile) * static_cast<size_t>(n_tiles);
    int rc = grow(&plan->h2d_pcm, &plan->h2d_pcm_bytes, sizeof(PcmT) * static_cast<size_t>(total_samples));
    if (rc == MFCC_OK)
        rc = grow(&plan->d2h_out, &plan->d2h_out_bytes, sizeof(float) * static_cast<size_t>(total_frames) * od);
    if (rc == MFCC_OK) rc = grow(&plan->d_tiles, &plan->d_tiles_bytes, tile_bytes);
    const size_t pc_bytes = sizeof(mfcc::PostChunk) * static_cast<size_t>(n_post_chunks);
    if (post != nullptr) {
        const size_t nu = static_cast<size_t>(n_utts), d = static_cast<size_t>(od);
        if (rc == MFCC_OK) rc = grow(&plan->d_post_out, &plan->d_post_out_bytes, sizeof(float) * static_cast<size_t>(total_frames) * od_out);
        if (rc == MFCC_OK) rc = grow(&plan->d_post_chunks, &plan->d_post_chunks_bytes, pc_bytes);
        if (rc == MFCC_OK) rc = grow(&plan->d_post_partial, &plan->d_post_partial_bytes, static_cast<size_t>(n_post_chunks) * d * mfcc::kPostPartialBytes);
        if (rc == MFCC_OK) rc = grow(&plan->d_post_stats, &plan->d_post_stats_bytes, nu * d * sizeof(float4));
        if (rc == MFCC_OK && plan->h_post_chunks_bytes < pc_bytes) {
            if (plan->h_post_chunks) cudaFreeHost(plan->h_post_chunks);
            plan->h_post_chunks = nullptr;
            plan->h_post_chunks_bytes = 0;
            const size_t want = align_up(pc_bytes + pc_bytes / 8, 1 << 16);
            if (cudaHostAlloc(&plan->h_post_chunks, want, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); rc = MFCC_ENOMEM; }
            else plan->h_post_chunks_bytes = want;
        }
    }
    if (rc == MFCC_OK && plan->h_tiles_bytes < tile_bytes) {   // pinned staging of the tile table
        if (plan->h_tiles) cudaFreeHost(plan->h_tiles);
        plan->h_tiles = nullptr;
        plan->h_tiles_bytes = 0;
        const size_t want = align_up(tile_bytes + tile_bytes / 8, 1 << 16);
        if (cudaHostAlloc(&plan->h_tiles, want, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); rc = MFCC_ENOMEM; }
        else plan->h_tiles_bytes = want;
    }
    for (auto &s : plan->streams)
        if (rc == MFCC_OK && s == nullptr && cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess)
            rc = MFCC_ECUDA;
    if (rc == MFCC_OK && plan->tiles_ready == nullptr &&
        cudaEventCreateWithFlags(&plan->tiles_ready, cudaEventDisableTiming) != cudaSuccess)
        rc = MFCC_ECUDA;
    if (rc != MFCC_OK) return rc;

    // 2. chunk plan (utterance boundaries) and the whole H2D queue
    std::vector<int64_t> cut{0};
    try {
        const int64_t full = (32ll << 20) / static_cast<int64_t>(sizeof(PcmT)), least = full / 16;   // 32 MiB .. 2 MiB of PCM
        int64_t u0 = 0;
        while (u0 < n_utts) {
            const int64_t left = total_samples - h_offsets[u0];
            const int64_t want = std::min(full, std::max(least, left / 3));
            int64_t u1 = u0 + 1;
            while (u1 < n_utts && h_offsets[u1 + 1] - h_offsets[u0] <= want) ++u1;
            cut.push_back(u1);
            u0 = u1;
        }
    } catch (const std::bad_alloc &) { return MFCC_ENOMEM; }
    const size_t n_chunks = cut.size() - 1;
    while (plan->chunk_ready.size() < n_chunks) {
        cudaEvent_t e = nullptr;
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return MFCC_ECUDA; }
        plan->chunk_ready.push_back(e);
    }
    PcmT *d_pcm = static_cast<PcmT *>(plan->h2d_pcm);
    float *d_out = static_cast<float *>(plan->d2h_out);
    bool ok = true;
    // chunk c on the wire: first the tiles of its utterances (built here, straight into the pinned staging table, while the
    // earlier chunks fly), then its PCM, on the same copy stream — so the chunk's event covers both.  Building the whole
    // table up front left the link idle: 1.0 ms for the 96 k tiles of the ragged telephony batch against 0.6 ms that the
    // first 32 MiB chunk is on the wire.
    Tile *h_tiles = static_cast<Tile *>(plan->h_tiles), *d_tiles = static_cast<Tile *>(plan->d_tiles);
    batch->d_tiles = d_tiles;
    batch->tiles_borrowed = true;
    // (With the post-processing kernels in the pipeline the table goes up in ONE piece behind the first chunk, together
    // with the chunk table of those kernels: 153 M frames/s on configs[1] against 151 M with per-chunk pieces.)
    const bool piecewise = post == nullptr;
    auto queue_h2d = [&](size_t c) {
        cudaStream_t copy = plan->streams[c & 1];
        const int64_t u0 = cut[c], u1 = cut[c + 1];
        const int64_t t0 = batch->utt_first_tile[u0], t1 = batch->utt_first_tile[u1];
        if (piecewise && t1 > t0) {
            batch_build_tiles(p, *batch, u0, u1, h_tiles + t0);
            ok = ok && cudaMemcpyAsync(d_tiles + t0, h_tiles + t0, sizeof(Tile) * static_cast<size_t>(t1 - t0), cudaMemcpyHostToDevice,
                                       copy) == cudaSuccess;
        }
        const int64_t s0 = h_offsets[u0], s1 = h_offsets[u1];
        if (s1 > s0)
            ok = ok && cudaMemcpyAsync(d_pcm + s0, h_pcm + s0, sizeof(PcmT) * (s1 - s0), cudaMemcpyHostToDevice,
                                       copy) == cudaSuccess;
        ok = ok && cudaEventRecord(plan->chunk_ready[c], copy) == cudaSuccess;
    };
    mark(0);
    queue_h2d(0);
    mark(1);

    // 3. post-processing only: the whole tile table and the chunk table of the post kernels, behind the first chunk (copies of
    // one direction are served in issue order: whatever is queued behind every PCM chunk holds the kernels back until the last
    // sample has arrived)
    if (post != nullptr) {
        if (!piecewise) {
            batch_build_tiles(p, *batch, 0, n_utts, h_tiles);
            ok = ok && cudaMemcpyAsync(d_tiles, h_tiles, tile_bytes, cudaMemcpyHostToDevice, plan->streams[2]) == cudaSuccess;
        }
        mfcc::post_build_chunks(batch->frame_offsets, od, batch->post_chunks, batch->utt_first_post_chunk, &batch->post_rows);
        ok = ok && static_cast<int64_t>(batch->post_chunks.size()) == n_post_chunks;
        if (ok) std::memcpy(plan->h_post_chunks, batch->post_chunks.data(), pc_bytes);
        ok = ok && cudaMemcpyAsync(plan->d_post_chunks, plan->h_post_chunks, pc_bytes, cudaMemcpyHostToDevice, plan->streams[2]) == cudaSuccess;
        ok = ok && cudaEventRecord(plan->tiles_ready, plan->streams[2]) == cudaSuccess;
        ok = ok && cudaStreamWaitEvent(plan->streams[3], plan->tiles_ready, 0) == cudaSuccess;
    }
    mark(2);
    // 4. kernels and result copies of a chunk, gated on the chunk's event
    auto queue_compute = [&](size_t c) {
        if (!ok) return;
        cudaStream_t s = plan->streams[2 + (c & 1)];
        const int64_t u0 = cut[c], u1 = cut[c + 1];
        const int64_t f0 = batch->frame_offsets[u0], f1 = batch->frame_offsets[u1];
        const int64_t t0 = batch->utt_first_tile[u0], t1 = batch->utt_first_tile[u1];
        ok = cudaStreamWaitEvent(s, plan->chunk_ready[c], 0) == cudaSuccess;
        if (ok && t1 > t0)
            ok = compute_batch_impl<PcmT>(plan, batch, d_pcm, d_out, t0, t1 - t0, s, alaw) == MFCC_OK;
        const float *d_rows = d_out;
        if (ok && post != nullptr && f1 > f0) {
            const int64_t c0 = batch->utt_first_post_chunk[u0], c1 = batch->utt_first_post_chunk[u1];
            const mfcc::PostView view{static_cast<const mfcc::PostChunk *>(plan->d_post_chunks), c0, c1 - c0, batch->post_rows,
                                      plan->d_post_partial, plan->d_post_stats, plan->device, plan->sm_count};
            ok = mfcc::launch_post(view, d_out, od, post->cmvn, post->order > 0 ? post->window : 1, post->order, 0,
                                   static_cast<float *>(plan->d_post_out), s) == MFCC_OK;
            d_rows = static_cast<const float *>(plan->d_post_out);
        }
        if (ok && f1 > f0)
            ok = cudaMemcpyAsync(h_out + f0 * od_out, d_rows + f0 * od_out, sizeof(float) * (f1 - f0) * od_out,
                                 cudaMemcpyDeviceToHost, s) == cudaSuccess;
    };
    // The queue is PACED: chunk c goes on the wire only when chunk c - 2 has landed, and the kernels + read-back of chunk
    // c - 1 are queued right behind it, so two input copies are in flight at any time and nothing else sits in the queues.
    // Queueing everything up front (round 1) makes the outcome depend on WHEN the host gets there: with the faster table build
    // of this round the read-back of the post-processing path lost its overlap with the input stream (9.3 ms per configs[1]
    // call instead of 7.0; a 0.4 ms busy-wait before the queueing restored it).  Measured on one box, M frames/s for
    // configs[1] plain / with post, ragged 8 kHz int16 / G.711 — unpaced: 162 / 110 / 304 / 472; depth 1: 156 / 151 / 291 /
    // 467; depth 2: 162 / 152 / 309 / 508; depth 3: 162 / 147 / 304 / 489; depth 4: 162 / 126 / 302 / 486.
    // (MFCC_HOST_DEPTH overrides the depth for such experiments; 0 = unpaced.)
    static const int depth = std::getenv("MFCC_HOST_DEPTH") ? std::atoi(std::getenv("MFCC_HOST_DEPTH")) : 2;
    for (size_t c = 1; c < n_chunks; ++c) {
        if (depth > 0 && c >= static_cast<size_t>(depth)) ok = ok && cudaEventSynchronize(plan->chunk_ready[c - depth]) == cudaSuccess;
        queue_h2d(c);
        if (depth > 0) queue_compute(c - 1);
    }
    mark(3);
    if (depth > 0) queue_compute(n_chunks - 1);
    else
        for (size_t c = 0; c < n_chunks; ++c) queue_compute(c);
    mark(4);
    for (auto &s : plan->streams)
        if (s && cudaStreamSynchronize(s) != cudaSuccess) ok = false;
    mark(5);
    if (trace)
        std::fprintf(stderr, "mfcc host path: %lld utts %lld tiles %zu chunks | sized %.3f first-h2d-queued %.3f post-table %.3f all-h2d-queued %.3f all-launched %.3f done %.3f ms\n",
                     static_cast<long long>(n_utts), static_cast<long long>(n_tiles), n_chunks, t_mark[0], t_mark[1], t_mark[2], t_mark[3], t_mark[4], t_mark[5]);
    if (!ok) { cudaGetLastError(); return MFCC_ECUDA; }
    return MFCC_OK;
}

}  // namespace

extern "C" {

int mfcc_compute_host(mfcc_plan *plan, const int16_t *h_pcm, const int64_t *h_offsets, int64_t n_utts,
                      float *h_out, int64_t *h_frame_offsets)
{
    return compute_host_impl<int16_t>(plan, h_pcm, 0, h_offsets, n_utts, h_out, h_frame_offsets);
}

// PCM in, the stacked static | delta | delta-delta rows out: the post-processing kernels run per chunk of utterances between
// the transform kernel and the read-back.
int mfcc_compute_host_post(mfcc_plan *plan, const int16_t *h_pcm, const int64_t *h_offsets, int64_t n_utts, int32_t cmvn,
                           int32_t delta_window, int32_t delta_order, float *h_out, int64_t *h_frame_offsets)
{
    if (cmvn < MFCC_CMVN_NONE || cmvn > MFCC_CMVN_MEAN_VAR || delta_order < 0 || delta_order > 2) return MFCC_EINVAL;
    if (delta_order > 0 && (delta_window < 1 || delta_window > 8)) return MFCC_EINVAL;
    const PostOpts post{cmvn, delta_window, delta_order};
    return compute_host_impl<int16_t>(plan, h_pcm, 0, h_offsets, n_utts, h_out, h_frame_offsets, &post);
}

// G.711 bytes end to end: 1 byte per sample over PCIe and from HBM, expanded inside the fused kernel's staging.
int mfcc_compute_host_g711(mfcc_plan *plan, const uint8_t *h_codes, int32_t alaw, const int64_t *h_offsets,
                           int64_t n_utts, float *h_out, int64_t *h_frame_offsets)
{
    return compute_host_impl<uint8_t>(plan, h_codes, alaw != 0, h_offsets, n_utts, h_out, h_frame_offsets);
}

int mfcc_compute_batch_g711(const mfcc_plan *plan, const mfcc_batch *batch, const uint8_t *d_codes, int32_t alaw,
                            float *d_out, void *cuda_stream)
{
    if (!batch_fits(plan, batch)) return MFCC_EINVAL;
    if (batch->total_frames == 0) return MFCC_OK;
    if (d_codes == nullptr || d_out == nullptr) return MFCC_EINVAL;
    DeviceGuard guard(plan->device);
    if (!guard.ok) return MFCC_ECUDA;
    return compute_batch_impl<uint8_t>(plan, batch, d_codes, d_out, 0, static_cast<int64_t>(batch->tiles.size()),
                                       static_cast<cudaStream_t>(cuda_stream), alaw != 0);
}

// ---- streaming ----
}  // extern "C"

struct mfcc_stream {
    mfcc_plan *plan = nullptr;
    std::vector<int16_t> buf;   // [history sample if any] + samples from the next frame's start on
    bool has_history = false;   // buf[0] is the sample before the next frame's start
    int64_t fed = 0;            // samples fed since the last reset
    int64_t emitted = 0;        // frames returned since the last reset
    int64_t skip = 0;           // incoming samples to drop first (hop_len > frame_len: gaps between frames)
};

namespace {

// Frames of buf that are due: complete frames, or at flush the zero-padded tail the offline count implies.
int64_t stream_due(const mfcc_stream *st, int64_t n_new, bool at_flush)
{
    const mfcc_params &p = st->plan->p;
    const int64_t total = st->fed + n_new;
    mfcc_params q = p;
    q.pad_mode = at_flush ? p.pad_mode : MFCC_PAD_NONE;
    return mfcc_num_frames(&q, total) - st->emitted;
}

int grow_pinned(void **ptr, size_t *have, size_t need)
{
    if (*have >= need) return MFCC_OK;
    if (*ptr) cudaFreeHost(*ptr);
    *ptr = nullptr;
    *have = 0;
    const size_t want = align_up(need + need / 4, 1 << 16);
    if (cudaHostAlloc(ptr, want, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return MFCC_ENOMEM; }
    *have = want;
    return MFCC_OK;
}

// One launch for any number of live streams of ONE plan.  Stream i contributes due[i] frames of its buffer (whose first
// sample is pre-emphasis history when has_history is set; zero fill past the end, as MFCC_PAD_ZERO_TAIL does).  The
// buffers are packed into ONE pinned staging array (each starting on a 16-byte boundary, so the tiles stay eligible for
// the bulk-copy path), the tile table into another: two H2D copies, one kernel, one D2H into pinned memory, then the
// rows are handed out to the callers' buffers.  All on the plan's stream 0, under the plan's mutex.
int streams_run(mfcc_plan *plan, mfcc_stream *const *sts, int64_t n_streams, const int64_t *due, float *const *outs)
{
    const mfcc_params &p = plan->p;
    const int od = plan->host.out_dim;
    std::vector<int64_t> base(n_streams + 1, 0), row0(n_streams + 1, 0);
    int64_t n_tiles = 0;
    for (int64_t i = 0; i < n_streams; ++i) {
        const int64_t len = due[i] > 0 ? static_cast<int64_t>(sts[i]->buf.size()) : 0;
        base[i + 1] = base[i] + ((len + 7) & ~static_cast<int64_t>(7));
        row0[i + 1] = row0[i] + std::max<int64_t>(due[i], 0);
        n_tiles += (std::max<int64_t>(due[i], 0) + mfcc::kTileFrames - 1) / mfcc::kTileFrames;
    }
    const int64_t total_samples = base[n_streams] + mfcc::kTileSpanSlack + 16, total_rows = row0[n_streams];
    if (total_rows == 0) return MFCC_OK;

    std::lock_guard<std::mutex> lock(plan->host_mutex);
    DeviceGuard guard(plan->device);
    if (!guard.ok) return MFCC_ECUDA;
    int rc = grow(&plan->h2d_pcm, &plan->h2d_pcm_bytes, sizeof(int16_t) * static_cast<size_t>(total_samples));
    if (rc == MFCC_OK) rc = grow(&plan->d2h_out, &plan->d2h_out_bytes, sizeof(float) * static_cast<size_t>(total_rows) * od);
    if (rc == MFCC_OK) rc = grow(&plan->d_tiles, &plan->d_tiles_bytes, sizeof(Tile) * static_cast<size_t>(n_tiles));
    if (rc == MFCC_OK) rc = grow_pinned(&plan->h_tiles, &plan->h_tiles_bytes, sizeof(Tile) * static_cast<size_t>(n_tiles));
    if (rc == MFCC_OK) rc = grow_pinned(&plan->h_stage_pcm, &plan->h_stage_pcm_bytes, sizeof(int16_t) * static_cast<size_t>(total_samples));
    if (rc == MFCC_OK) rc = grow_pinned(&plan->h_stage_out, &plan->h_stage_out_bytes, sizeof(float) * static_cast<size_t>(total_rows) * od);
    if (rc == MFCC_OK && plan->streams[0] == nullptr &&
        cudaStreamCreateWithFlags(&plan->streams[0], cudaStreamNonBlocking) != cudaSuccess)
        rc = MFCC_ECUDA;
    if (rc != MFCC_OK) return rc;

    int16_t *stage = static_cast<int16_t *>(plan->h_stage_pcm);
    Tile *tiles = static_cast<Tile *>(plan->h_tiles);
    int64_t t = 0;
    for (int64_t i = 0; i < n_streams; ++i) {
        if (due[i] <= 0) continue;
        const mfcc_stream *st = sts[i];
        const int64_t len = static_cast<int64_t>(st->buf.size());
        std::memcpy(stage + base[i], st->buf.data(), sizeof(int16_t) * len);
        if (base[i + 1] > base[i] + len) std::memset(stage + base[i] + len, 0, sizeof(int16_t) * (base[i + 1] - base[i] - len));
        const int lead = st->has_history ? 1 : 0;
        for (int64_t f = 0; f < due[i]; f += mfcc::kTileFrames) {
            Tile &tl = tiles[t++];
            tl.utt_begin = base[i];
            tl.utt_end = base[i] + len;
            tl.first_sample = base[i] + lead + f * p.hop_len;
            tl.out_row = row0[i] + f;
            tl.n_frames = static_cast<int32_t>(std::min<int64_t>(mfcc::kTileFrames, due[i] - f));
            tl.flags = mfcc::tile_flags(p, tl, total_samples);
        }
    }
    std::memset(stage + base[n_streams], 0, sizeof(int16_t) * (total_samples - base[n_streams]));
    cudaStream_t s = plan->streams[0];
    mfcc_batch b;
    b.device = plan->device;
    b.out_dim = od;
    b.frame_len = p.frame_len;
    b.hop_len = p.hop_len;
    b.pad_mode = p.pad_mode;
    b.total_samples = total_samples;
    b.total_frames = total_rows;
    b.d_tiles = static_cast<Tile *>(plan->d_tiles);
    b.tiles_borrowed = true;
    bool ok = cudaMemcpyAsync(plan->d_tiles, tiles, sizeof(Tile) * static_cast<size_t>(n_tiles), cudaMemcpyHostToDevice, s) == cudaSuccess;
    ok = ok && cudaMemcpyAsync(plan->h2d_pcm, stage, sizeof(int16_t) * static_cast<size_t>(total_samples), cudaMemcpyHostToDevice, s) == cudaSuccess;
    ok = ok && compute_batch_impl<int16_t>(plan, &b, static_cast<const int16_t *>(plan->h2d_pcm),
                                           static_cast<float *>(plan->d2h_out), 0, n_tiles, s) == MFCC_OK;
    ok = ok && cudaMemcpyAsync(plan->h_stage_out, plan->d2h_out, sizeof(float) * static_cast<size_t>(total_rows) * od,
                               cudaMemcpyDeviceToHost, s) == cudaSuccess;
    ok = ok && cudaStreamSynchronize(s) == cudaSuccess;
    if (!ok) { cudaGetLastError(); return MFCC_ECUDA; }
    const float *res = static_cast<const float *>(plan->h_stage_out);
    for (int64_t i = 0; i < n_streams; ++i)
        if (due[i] > 0) std::memcpy(outs[i], res + row0[i] * od, sizeof(float) * static_cast<size_t>(due[i]) * od);
    return MFCC_OK;
}

int stream_run(mfcc_stream *st, int64_t n_frames, float *out)
{
    mfcc_stream *one[1] = {st};
    float *outs[1] = {out};
    return streams_run(st->plan, one, 1, &n_frames, outs);
}

// Drop what the emitted frames consumed: keep one history sample and everything from the next frame's start.
void stream_advance(mfcc_stream *st, int64_t n_frames)
{
    if (n_frames <= 0) return;
    const int lead = st->has_history ? 1 : 0;
    const int64_t next = lead + n_frames * st->plan->p.hop_len;   // index in buf of the next frame's start
    const int64_t len = static_cast<int64_t>(st->buf.size());
    if (next - 1 >= len) {          // the next frame's history sample has not arrived yet (hop > frame)
        st->skip = next - 1 - len;
        st->buf.clear();
        st->has_history = true;     // the first sample kept after the skip is that history sample
    } else {
        st->buf.erase(st->buf.begin(), st->buf.begin() + (next - 1));
        st->has_history = true;
    }
    st->emitted += n_frames;
}

}  // namespace

extern "C" {

int mfcc_stream_create(mfcc_plan *plan, mfcc_stream **out)
{
    if (out == nullptr) return MFCC_EINVAL;
    *out = nullptr;
    if (plan == nullptr) return MFCC_EINVAL;
    mfcc_stream *st = new (std::nothrow) mfcc_stream();
    if (st == nullptr) return MFCC_ENOMEM;
    st->plan = plan;
    *out = st;
    return MFCC_OK;
}

void mfcc_stream_destroy(mfcc_stream *st) { delete st; }

int64_t mfcc_stream_pending(const mfcc_stream *st, int64_t n_new, int32_t at_flush)
{
    if (st == nullptr || n_new < 0) return MFCC_EINVAL;
    return stream_due(st, n_new, at_flush != 0);
}

int mfcc_stream_feed(mfcc_stream *st, const int16_t *pcm, int64_t n, float *out, int64_t max_frames,
                     int64_t *n_frames)
{
    if (st == nullptr || n < 0 || (n > 0 && pcm == nullptr)) return MFCC_EINVAL;
    const int64_t due = stream_due(st, n, false);
    if (due > max_frames || (due > 0 && out == nullptr)) return MFCC_EINVAL;
    const int64_t drop = std::min(st->skip, n);
    try {
        st->buf.insert(st->buf.end(), pcm + drop, pcm + n);
    } catch (const std::bad_alloc &) {
        return MFCC_ENOMEM;
    }
    st->skip -= drop;
    st->fed += n;
    if (n_frames) *n_frames = 0;
    if (due <= 0) return MFCC_OK;
    const int rc = stream_run(st, due, out);
    if (rc != MFCC_OK) return rc;
    stream_advance(st, due);
    if (n_frames) *n_frames = due;
    return MFCC_OK;
}

int mfcc_stream_feed_many(mfcc_stream *const *streams, int64_t n_streams, const int16_t *const *pcm, const int64_t *n,
                          float *const *out, const int64_t *max_frames, int64_t *n_frames)
{
    if (n_streams < 0 || (n_streams > 0 && (streams == nullptr || pcm == nullptr || n == nullptr || out == nullptr ||
                                            max_frames == nullptr))) return MFCC_EINVAL;
    if (n_streams == 0) return MFCC_OK;
    // validate everything before touching any stream: the call either feeds all of them or none
    std::vector<int64_t> due(n_streams);
    for (int64_t i = 0; i < n_streams; ++i) {
        mfcc_stream *st = streams[i];
        if (st == nullptr || st->plan != streams[0]->plan || n[i] < 0 || (n[i] > 0 && pcm[i] == nullptr)) return MFCC_EINVAL;
        for (int64_t k = 0; k < i && n_streams <= 64; ++k)
            if (streams[k] == st) return MFCC_EINVAL;   // (a stream twice in one call; checked for small calls)
        due[i] = stream_due(st, n[i], false);
        if (due[i] > max_frames[i] || (due[i] > 0 && out[i] == nullptr)) return MFCC_EINVAL;
    }
    try {
        for (int64_t i = 0; i < n_streams; ++i) {
            mfcc_stream *st = streams[i];
            const int64_t drop = std::min(st->skip, n[i]);
            st->buf.insert(st->buf.end(), pcm[i] + drop, pcm[i] + n[i]);
            st->skip -= drop;
            st->fed += n[i];
            if (n_frames) n_frames[i] = 0;
        }
    } catch (const std::bad_alloc &) {
        return MFCC_ENOMEM;
    }
    const int rc = streams_run(streams[0]->plan, streams, n_streams, due.data(), out);
    if (rc != MFCC_OK) return rc;
    for (int64_t i = 0; i < n_streams; ++i) {
        if (due[i] <= 0) continue;
        stream_advance(streams[i], due[i]);
        if (n_frames) n_frames[i] = due[i];
    }
    return MFCC_OK;
}

int mfcc_stream_flush(mfcc_stream *st, float *out, int64_t max_frames, int64_t *n_frames)
{
    if (st == nullptr) return MFCC_EINVAL;
    const int64_t due = stream_due(st, 0, true);
    if (due > max_frames || (due > 0 && out == nullptr)) return MFCC_EINVAL;
    if (n_frames) *n_frames = 0;
    int rc = MFCC_OK;
    if (due > 0) {
        rc = stream_run(st, due, out);
        if (rc == MFCC_OK && n_frames) *n_frames = due;
    }
    st->buf.clear();
    st->has_history = false;
    st->fed = 0;
    st->emitted = 0;
    st->skip = 0;
    return rc;
}

int mfcc_compute(mfcc_plan *plan, const int16_t *pcm, int64_t n_samples, float *out, int64_t *n_frames)
{
    if (plan == nullptr || n_samples < 0) return MFCC_EINVAL;
    const int64_t offsets[2] = {0, n_samples};
    int64_t fo[2] = {0, 0};
    const int rc = mfcc_compute_host(plan, pcm, offsets, 1, out, fo);
    if (rc == MFCC_OK && n_frames) *n_frames = fo[1];
    return rc;
}

int mfcc_cmvn_batch(const mfcc_plan *plan, const mfcc_batch *batch, float *d_feat, int32_t norm_var,
                    void *cuda_stream)
{
    if (!batch_fits(plan, batch)) return MFCC_EINVAL;
    if (batch->total_frames == 0) return MFCC_OK;
    if (d_feat == nullptr) return MFCC_EINVAL;
    DeviceGuard guard(plan->device);
    if (!guard.ok) return MFCC_ECUDA;
    // the fused kernels with no regression: a CTA reads its rows into shared memory, then writes them back normalised
    const mfcc::PostView view{batch->d_post_chunks, 0, static_cast<int64_t>(batch->post_chunks.size()), batch->post_rows,
                              batch->d_post_partial, batch->d_post_stats, batch->device, plan->sm_count};
    return mfcc::launch_post(view, d_feat, batch->out_dim, norm_var != 0 ? MFCC_CMVN_MEAN_VAR : MFCC_CMVN_MEAN, 1, 0, 0, d_feat,
                             static_cast<cudaStream_t>(cuda_stream));
}

int mfcc_delta_batch(const mfcc_plan *plan, const mfcc_batch *batch, const float *d_feat, int32_t window,
                     float *d_delta, void *cuda_stream)
{
    if (!batch_fits(plan, batch)) return MFCC_EINVAL;
    if (window < 1 || window > 8) return MFCC_EINVAL;
    if (batch->total_frames == 0) return MFCC_OK;
    if (d_feat == nullptr || d_delta == nullptr) return MFCC_EINVAL;
    DeviceGuard guard(plan->device);
    if (!guard.ok) return MFCC_ECUDA;
    // the fused kernels without normalisation, writing the regression part only
    const mfcc::PostView view{batch->d_post_chunks, 0, static_cast<int64_t>(batch->post_chunks.size()), batch->post_rows,
                              batch->d_post_partial, batch->d_post_stats, batch->device, plan->sm_count};
    return mfcc::launch_post(view, d_feat, batch->out_dim, MFCC_CMVN_NONE, window, 1, 1, d_delta,
                             static_cast<cudaStream_t>(cuda_stream));
}

int mfcc_post_batch(const mfcc_plan *plan, const mfcc_batch *batch, const float *d_feat, int32_t cmvn,
                    int32_t delta_window, int32_t delta_order, float *d_out, void *cuda_stream)
{
    if (!batch_fits(plan, batch)) return MFCC_EINVAL;
    if (cmvn < MFCC_CMVN_NONE || cmvn > MFCC_CMVN_MEAN_VAR || delta_order < 0 || delta_order > 2) return MFCC_EINVAL;
    if (delta_order > 0 && (delta_window < 1 || delta_window > 8)) return MFCC_EINVAL;
    if (batch->total_frames == 0) return MFCC_OK;
    if (d_feat == nullptr || d_out == nullptr) return MFCC_EINVAL;
    {   // the input is read through the read-only path while the output is written: the two must not overlap
        const char *a = reinterpret_cast<const char *>(d_feat), *o = reinterpret_cast<const char *>(d_out);
        const size_t ab = sizeof(float) * static_cast<size_t>(batch->total_frames) * batch->out_dim;
        if (a < o + ab * (1 + delta_order) && o < a + ab) return MFCC_EINVAL;
    }
    DeviceGuard guard(plan->device);
    if (!guard.ok) return MFCC_ECUDA;
    const mfcc::PostView view{batch->d_post_chunks, 0, static_cast<int64_t>(batch->post_chunks.size()), batch->post_rows,
                              batch->d_post_partial, batch->d_post_stats, batch->device, plan->sm_count};
    return mfcc::launch_post(view, d_feat, batch->out_dim, cmvn, delta_order > 0 ? delta_window : 1, delta_order, 0, d_out,
                             static_cast<cudaStream_t>(cuda_stream));
}

int mfcc_decode_g711(const uint8_t *d_src, int64_t n, int32_t alaw, int16_t *d_dst, void *cuda_stream)
{
    if (n < 0 || (n > 0 && (d_src == nullptr || d_dst == nullptr))) return MFCC_EINVAL;
    return mfcc::launch_g711(d_src, n, alaw != 0, d_dst, static_cast<cudaStream_t>(cuda_stream));
}

int mfcc_host_alloc(void **ptr, int64_t bytes)
{
    if (ptr == nullptr || bytes <= 0) return MFCC_EINVAL;
    if (cudaHostAlloc(ptr, static_cast<size_t>(bytes), cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        *ptr = nullptr;
        return MFCC_ENOMEM;
    }
    return MFCC_OK;
}

int mfcc_host_free(void *ptr)
{
    if (ptr == nullptr) return MFCC_OK;
    return cudaFreeHost(ptr) == cudaSuccess ? MFCC_OK : MFCC_ECUDA;
}

uint64_t mfcc_launch_count(void) { return mfcc::g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
