// mfcc_host.h — internal types shared by the host layer and the kernels of
// libmfcc_b200.so.  Not installed; the public surface is include/mfcc_b200.h.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/mfcc_b200.h"

namespace mfcc {

constexpr int kTileFrames = 32;  // frames per tile: one frame per lane in the fused kernels

// One tile = up to kTileFrames consecutive frames of ONE utterance.
// All sample positions index the concatenated PCM array of the batch.
struct Tile {
    int64_t utt_begin;     // first sample of the utterance (pre-emphasis boundary: x[-1] = 0)
    int64_t utt_end;       // one past the last sample (zero beyond it under MFCC_PAD_ZERO_TAIL)
    int64_t first_sample;  // start of the tile's frame 0
    int64_t out_row;       // output row of the tile's frame 0
    int32_t n_frames;      // 1..kTileFrames
    int32_t flags;         // kTileInside | ...
};

// Tile::flags bit 0: every frame of the tile lies inside its utterance (no zero fill) and the span a fused
// kernel stages for it — from the 8-sample boundary below first_sample to the last frame's end plus
// kTileSpanSlack samples of transform slack and rounding — lies inside the batch's PCM array.  Such a tile can
// be fetched by bulk copy / 16-byte loads whatever the alignment of its utterance.
constexpr int32_t kTileInside = 1;
constexpr int kTileSpanSlack = 64;
inline int32_t tile_flags(const mfcc_params &p, const Tile &t, int64_t total_samples)
{
    const int64_t last_end = t.first_sample + static_cast<int64_t>(t.n_frames - 1) * p.hop_len + p.frame_len;
    const int64_t o = t.first_sample & ~static_cast<int64_t>(7);
    return (last_end <= t.utt_end && o + ((last_end - o + kTileSpanSlack + 7) & ~static_cast<int64_t>(7)) <= total_samples)
               ? kTileInside : 0;
}

// Post-processing (mfcc_post.cu): one chunk = up to mfcc_batch::post_rows consecutive feature rows of ONE utterance, i.e. one
// contiguous block of the feature matrix and of the stacked output matrix.
struct PostChunk {
    int64_t row0;          // first row of the chunk
    int64_t f0, f1;        // row range of its utterance (regression indices are clamped to it)
    int32_t n;             // rows in the chunk
    int32_t utt;
    int32_t first_chunk;   // chunk range of the utterance: the statistics are combined in this order
    int32_t n_chunks;
};
constexpr size_t kPostSmemMax = 96 * 1024;

// Host-side tables, evaluated in double and rounded once to f32 (DESIGN.md "Tables").
struct HostTables {
    int nbins = 0;                   // nfft/2 + 1
    int out_dim = 0;                 // n_cep or n_mel
    std::vector<float> window;       // [frame_len]
    std::vector<int32_t> mel_bins;   // [n_mel + 2]
    std::vector<float> mel_w;        // [n_mel][nbins] dense
    std::vector<float> dct;          // [n_cep][n_mel], lifter folded in
    std::vector<float> tw_re, tw_im; // [nfft/2] exp(-2 pi i k / nfft)
    // Segment form of the filterbank used by the fused kernels: bin k in segment
    // j = [bins[j], bins[j+1]) feeds filter j with rise[k] and filter j-1 with fall[k].
    std::vector<float> rise, fall;   // [nbins]
};

int build_tables(const mfcc_params &p, HostTables &t);
int validate_params(const mfcc_params *p);

// Device-side view handed to kernels by value.
struct DevTables {
    const float *window;    // [nfft], zero past frame_len
    const float2 *twiddle;  // [nfft/2]
    const float *mel_w;     // [n_mel][nbins]
    const int32_t *mel_bins;// [n_mel + 2]
    const float *dct;       // [n_cep][n_mel]
    const float *rise;      // [nbins]
    const float *fall;      // [nbins]
};

}  // namespace mfcc

struct mfcc_plan {
    mfcc_params p{};
    int device = 0;
    int kernel = MFCC_KERNEL_GENERIC;   // resolved: GENERIC or FUSED
    int sm_count = 0;
    std::string kernel_name;
    mfcc::HostTables host;
    mfcc::DevTables dev{};
    void *dev_blob = nullptr;           // one allocation backing every DevTables pointer
    void *sp_state = nullptr;           // 512- / 256-point fused tile kernel state (mfcc_fused_sp.cu)
    void *wide_state = nullptr;         // large-transform fused kernel state (mfcc_fused_wide.cu)
    // host-buffer entry points (mfcc_compute_host, mfcc_stream_*): state grown on demand and reused across calls,
    // guarded by host_mutex — one such call at a time per plan
    std::mutex host_mutex;
    void *h2d_pcm = nullptr;   size_t h2d_pcm_bytes = 0;
    void *d2h_out = nullptr;   size_t d2h_out_bytes = 0;
    void *d_tiles = nullptr;   size_t d_tiles_bytes = 0;   // device tile table of the call in flight
    void *h_tiles = nullptr;   size_t h_tiles_bytes = 0;   // its pinned staging copy
    void *h_stage_pcm = nullptr; size_t h_stage_pcm_bytes = 0;   // streaming: pinned staging of the packed stream buffers
    void *h_stage_out = nullptr; size_t h_stage_out_bytes = 0;   // streaming: pinned landing area of the feature rows
    // mfcc_compute_host_post: stacked output rows, chunk table (device + pinned staging) and statistics scratch
    void *d_post_out = nullptr;      size_t d_post_out_bytes = 0;
    void *d_post_chunks = nullptr;   size_t d_post_chunks_bytes = 0;
    void *h_post_chunks = nullptr;   size_t h_post_chunks_bytes = 0;
    void *d_post_partial = nullptr;  size_t d_post_partial_bytes = 0;
    void *d_post_stats = nullptr;    size_t d_post_stats_bytes = 0;

    cudaEvent_t tiles_ready = nullptr;
    std::vector<cudaEvent_t> chunk_ready;   // one per H2D chunk of the call in flight (reused across calls)
    cudaStream_t streams[4] = {nullptr, nullptr, nullptr, nullptr};   // compute_host: two H2D queues, two compute + D2H queues
};

struct mfcc_batch {
    int device = 0;
    int64_t n_utts = 0;
    int64_t total_frames = 0;
    int64_t total_samples = 0;
    int out_dim = 0;
    int frame_len = 0, hop_len = 0, pad_mode = 0;   // framing of the creating plan: the tile table is only valid for it
    std::vector<int64_t> offsets;        // [n_utts + 1]
    std::vector<int64_t> frame_offsets;  // [n_utts + 1]
    std::vector<mfcc::Tile> tiles;       // host copy
    std::vector<int64_t> utt_first_tile; // [n_utts + 1] tile index range per utterance
    std::vector<uint8_t> lead;           // [n_utts] or empty: the utterance starts with one history sample (mfcc_batch_create_lead)
    mfcc::Tile *d_tiles = nullptr;
    bool tiles_borrowed = false;         // d_tiles belongs to the plan (mfcc_compute_host)
    int64_t *d_frame_offsets = nullptr;
    // post-processing (mfcc_post_batch): chunk table and the statistics scratch, allocated with the batch so that the
    // call itself allocates nothing; calls on one batch must be stream-ordered (they share the scratch)
    std::vector<mfcc::PostChunk> post_chunks;
    std::vector<int64_t> utt_first_post_chunk;   // [n_utts + 1]
    int post_rows = 0;
    mfcc::PostChunk *d_post_chunks = nullptr;
    void *d_post_partial = nullptr;      // [chunks][out_dim] x kPostPartialBytes: per-group sums about the group's first row
    void *d_post_stats = nullptr;        // [n_utts][out_dim] float4 {mean hi, lo, 1 / sigma hi, lo}
};

namespace mfcc {

extern std::atomic<uint64_t> g_launches;

// Function attributes belong to a device: the opt-in to more than 48 KB of dynamic shared memory is made once per
// (kernel instantiation, device).  `done` is a per-instantiation bit mask over device ordinals (ordinals >= 64 opt in
// on every launch).
template <typename Kern>
inline int ensure_smem_optin(Kern kern, int device, size_t bytes, std::atomic<uint64_t> &done)
{
    const uint64_t bit = device >= 0 && device < 64 ? 1ull << device : 0;
    if (bit != 0 && (done.load(std::memory_order_acquire) & bit) != 0) return MFCC_OK;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)) != cudaSuccess) {
        cudaGetLastError();
        return MFCC_ECUDA;
    }
    if (bit != 0) done.fetch_or(bit, std::memory_order_release);
    return MFCC_OK;
}

// Kernel launchers.  `tile0`/`n_tiles` select a range of the batch's tile table;
// d_pcm / d_out are the bases of the WHOLE batch arrays.
// PcmT: int16_t, float (scaled to the int16 range) or uint8_t (G.711 codes; `alaw` selects the law, ignored otherwise)
template <typename PcmT>
int launch_generic(const mfcc_plan *plan, const Tile *d_tiles, int64_t n_tiles, const PcmT *d_pcm,
                   float *d_out, int alaw, cudaStream_t stream);

// Fused tile kernel for the 512- and 256-point geometries (mfcc_fused_sp.cu): name if the plan has one, else nullptr.
const char *sp_match(const mfcc_params &p, const HostTables &h);
int sp_prepare(mfcc_plan *plan);
void sp_release(mfcc_plan *plan);
template <typename PcmT>
int sp_launch(const mfcc_plan *plan, const Tile *d_tiles, int64_t n_tiles, const PcmT *d_pcm, int64_t pcm_len,
              float *d_out, int alaw, cudaStream_t stream);
// Large-transform variant (mfcc_fused_wide.cu): 2048-point frames, 8 frames per tile, 4 items per warp.
const char *wide_match(const mfcc_params &p, const HostTables &h);
int wide_prepare(mfcc_plan *plan);
void wide_release(mfcc_plan *plan);
template <typename PcmT>
int wide_launch(const mfcc_plan *plan, const Tile *d_tiles, int64_t n_tiles, const PcmT *d_pcm, int64_t pcm_len,
                float *d_out, cudaStream_t stream);
// what a launch of the post-processing kernels works on: the chunks [chunk0, chunk0 + n_chunks) of a device chunk table and
// the statistics scratch that goes with the table (indexed by global chunk / utterance number)
struct PostView {
    const PostChunk *chunks;
    int64_t chunk0, n_chunks;
    int rows;            // rows per chunk the table was cut with
    void *partial;       // [chunks][dim] x 32 bytes (kPostPartialBytes)
    void *stats;         // [utterances][dim] float4
    int device;
    int sms;             // SMs of the device (grid of the persistent statistics kernel)
};
constexpr size_t kPostPartialBytes = 32;
int post_rows_for(int dim);
void post_build_chunks(const std::vector<int64_t> &frame_offsets, int dim, std::vector<PostChunk> &chunks,
                       std::vector<int64_t> &utt_first, int *rows_out);
size_t post_smem_bytes(int dim, int rows, int window, int order);
// part0 = 0: out rows are static | delta | delta-delta (1 + order parts); part0 = 1: the regressions only (order parts)
int launch_post(const PostView &v, const float *d_feat, int dim, int cmvn, int window, int order, int part0, float *d_out,
                cudaStream_t s);
int launch_g711(const uint8_t *d_src, int64_t n, int alaw, int16_t *d_dst, cudaStream_t s);

}  // namespace mfcc
