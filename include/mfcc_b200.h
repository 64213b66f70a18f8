/*
 * mfcc_b200.h — C ABI of the B200-native MFCC front end (libmfcc_b200.so).
 *
 * Boundary note.  The nominal reference, simotin13/mfcc, is a toy C compiler
 * ("mf C Compiler"); it has no MFCC API to bind to.  Its
 * only public entry points are
 *     int tokenize(char*, unsigned, Vector*)                    src/mfcc/lex.h:75
 *     int parse_tokens(Vector*, Vector*, Program*)              src/mfcc/parser.h:17
 *     int generate_binary(char*, Vector*, Program*, BuildTargetType)
 *                                                               src/mfcc/codegen.h:18
 * and the header the north star imagines holds the MFCC parameters,
 * src/mfcc/mfcc.h:1-22, only defines STRING_MAX / CODE_LEN_MAX / BuildTargetType.
 * What IS kept from the reference is its convention:
 *   - plain C, plain pointers and sizes, caller-owned buffers filled in place
 *     (src/mfcc/main.c:64-66);
 *   - `int` return, 0 = ok, negative = failure (src/mfcc/main.c:72-76,78-82,
 *     src/mfcc/lex.c:165,197,320) — but a library never exit()s or assert()s.
 * The parameter list is the one BASELINE.json's north_star names: sample rate,
 * frame length / hop, pre-emphasis, window, FFT size, mel-band count, cepstral
 * count.  Every self-chosen convention is a field so a real reference can be
 * matched later.
 *
 * There is NO CPU fallback behind any entry point that computes: every
 * compute call runs hand-written sm_100a kernels or fails with MFCC_ECUDA.
 */
#ifndef MFCC_B200_H_
#define MFCC_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- error codes (0 = ok, negative = failure; cf. src/mfcc/main.c:72-76) ---- */
#define MFCC_OK        0
#define MFCC_EINVAL   (-1)  /* bad parameter / NULL pointer / inconsistent offsets */
#define MFCC_ENOMEM   (-2)  /* host or device allocation failed */
#define MFCC_ECUDA    (-3)  /* CUDA runtime error or no sm_100 device */
#define MFCC_ENOTSUP  (-4)  /* valid request this build does not implement */

/* ---- enumerations ---- */
#define MFCC_WINDOW_RECT     0
#define MFCC_WINDOW_HAMMING  1   /* 0.54 - 0.46 cos(2 pi n / (L-1)) */
#define MFCC_WINDOW_HANN     2   /* 0.5  - 0.5  cos(2 pi n / (L-1)) */

#define MFCC_PAD_NONE        0   /* n_frames = n < L ? 0 : 1 + (n - L) / hop  (truncate) */
#define MFCC_PAD_ZERO_TAIL   1   /* n_frames = n <= 0 ? 0 : 1 + ceil(max(n - L, 0) / hop), zeros past the end */

#define MFCC_OUT_CEPSTRA     0   /* [frames][n_cep]  DCT-II of log mel energies */
#define MFCC_OUT_LOGMEL      1   /* [frames][n_mel]  log mel energies ("fbank") */

#define MFCC_ENERGY_NONE        0
#define MFCC_ENERGY_REPLACE_C0  1   /* cepstra only: c[0] = ln(max(E, log_floor)) instead of the DCT's c0 */
#define MFCC_ENERGY_APPEND      2   /* one more column after the n_cep / n_mel ones: ln(max(E, log_floor)) */
/* E = sum of the frame's one-sided power spectrum P[0 .. nfft/2] (the "appendEnergy" convention of the common Python
 * front ends: total energy of the windowed, pre-emphasised frame as the filterbank sees it). */

#define MFCC_KERNEL_AUTO     0   /* fused tile kernel when the geometry has one, else generic */
#define MFCC_KERNEL_GENERIC  1   /* one-frame-at-a-time shared-memory radix-2 kernel (any geometry) */
#define MFCC_KERNEL_FUSED    2   /* the fused tile kernel; plan creation fails with MFCC_ENOTSUP if the geometry has none */

/* ---- parameters (all conventions explicit; see DESIGN.md "Spec") ---- */
typedef struct mfcc_params {
    int32_t sample_rate;  /* Hz */
    int32_t frame_len;    /* samples per frame (25 ms -> 400 @ 16 kHz) */
    int32_t hop_len;      /* samples between frame starts (10 ms -> 160) */
    int32_t nfft;         /* power of two, frame_len <= nfft <= 4096 */
    int32_t n_mel;        /* mel bands, 1..128 */
    int32_t n_cep;        /* cepstra kept, 1..n_mel */
    float   preemph;      /* y[n] = x[n] - preemph * x[n-1] over the whole utterance, y[0] = x[0] */
    int32_t window;       /* MFCC_WINDOW_* */
    float   f_lo;         /* lowest mel edge, Hz */
    float   f_hi;         /* highest mel edge, Hz; <= 0 means sample_rate / 2 */
    float   log_floor;    /* L[m] = ln(max(E[m], log_floor)) */
    int32_t lifter;       /* 0 = none; Q > 0: c[k] *= 1 + (Q/2) sin(pi k / Q) */
    int32_t pad_mode;     /* MFCC_PAD_* */
    int32_t output;       /* MFCC_OUT_* */
    int32_t energy;       /* MFCC_ENERGY_* */
} mfcc_params;

/* Threading.  A plan's parameters and device tables never change after creation, and mfcc_compute_batch(_f32, _g711)
 * touch nothing else: they are reentrant (any number of host threads, distinct streams).  mfcc_post_batch,
 * mfcc_cmvn_batch and mfcc_delta_batch additionally use the statistics scratch that belongs to the BATCH: reentrant across
 * batches, stream-ordered on one batch.  The calls that take HOST buffers (mfcc_compute_host, mfcc_compute, mfcc_compute_host_g711 and every
 * mfcc_stream_* / mfcc_stream_feed_many call on the plan) share plan-owned staging buffers, streams and events that
 * grow on demand; they serialise on a mutex inside the plan, so they are safe to call from several threads but run
 * one at a time per plan.  Use one plan per thread for concurrent host-buffer traffic. */
typedef struct mfcc_plan  mfcc_plan;   /* owns device tables (immutable) and the host-path staging state (mutex-guarded) */
typedef struct mfcc_batch mfcc_batch;  /* the shape of one batch: offsets -> frame rows -> tiles; tied to the framing
                                          (frame_len, hop_len, pad_mode, out_dim) of the plan that created it */

/* Fill *p with the repo defaults for a sample rate: 25 ms frame, 10 ms hop,
 * nfft = next power of two, 26 mel, 13 cepstra, preemph 0.97, Hamming,
 * f_lo 0, f_hi sr/2, floor 1e-10, no lifter, no padding, cepstra out, no energy term. */
int mfcc_params_init(mfcc_params *p, int32_t sample_rate);

/* 0 if the parameter set is usable, MFCC_EINVAL otherwise.  Pure host code. */
int mfcc_params_validate(const mfcc_params *p);

/* Frames an utterance of n_samples yields under p->pad_mode (pure host code;
 * bit-exact framing contract).  Negative on bad parameters. */
int64_t mfcc_num_frames(const mfcc_params *p, int64_t n_samples);

/* Floats per output frame: n_cep or n_mel depending on p->output, plus one under MFCC_ENERGY_APPEND. */
int32_t mfcc_out_dim(const mfcc_params *p);

/* Build window / mel / DCT / twiddle tables in double, round once to f32 and
 * upload them to CUDA device `device` (-1 = current).  `kernel` is MFCC_KERNEL_*. */
int mfcc_plan_create(const mfcc_params *p, int32_t device, int32_t kernel, mfcc_plan **out);
void mfcc_plan_destroy(mfcc_plan *plan);
int mfcc_plan_params(const mfcc_plan *plan, mfcc_params *out);
/* Name of the kernel family the plan launches ("fused_r16x16_n512", "generic_radix2", ...). */
const char *mfcc_plan_kernel_name(const mfcc_plan *plan);

/* Host copies of the plan's f32 tables, for tests and for callers that want to
 * reproduce the arithmetic.  Each returns the element count (or negative);
 * pass NULL to query the count.
 *   window : frame_len               mel_bins : n_mel + 2 (int32 bin edges)
 *   mel_w  : n_mel * (nfft/2 + 1)    dct      : n_out_cep * n_mel (lifter folded in) */
int64_t mfcc_plan_window(const mfcc_plan *plan, float *dst);
int64_t mfcc_plan_mel_bins(const mfcc_plan *plan, int32_t *dst);
int64_t mfcc_plan_mel_weights(const mfcc_plan *plan, float *dst);
int64_t mfcc_plan_dct(const mfcc_plan *plan, float *dst);

/* Describe a batch: utterance u is pcm[h_offsets[u] .. h_offsets[u+1]) in one
 * concatenated int16 array (h_offsets is HOST memory, n_utts + 1 entries,
 * non-decreasing).  Computes frame rows and the tile table and uploads it. */
int mfcc_batch_create(const mfcc_plan *plan, const int64_t *h_offsets, int64_t n_utts,
                      mfcc_batch **out);
/* The same for PIECES of longer recordings (a 10-minute stream cut across GPUs, a file processed in slices): h_lead[u] != 0
 * says the first sample of utterance u is HISTORY — the sample right before the piece in its recording.  It only serves as
 * the pre-emphasis predecessor of the piece's first frame; the piece's frames start at h_offsets[u] + 1.  With pieces cut at
 * frame boundaries (mfcc_piece_span) the rows of the pieces, one after the other, are the rows of the whole recording bit
 * for bit.  h_lead == NULL: mfcc_batch_create. */
int mfcc_batch_create_lead(const mfcc_plan *plan, const int64_t *h_offsets, const uint8_t *h_lead, int64_t n_utts,
                           mfcc_batch **out);
/* Sample span [*begin, *end) of the piece that holds frames [f0, f1) of a recording of n_samples samples, and whether it
 * starts with one history sample (*lead = 1 whenever f0 > 0).  Pure host code. */
int mfcc_piece_span(const mfcc_params *p, int64_t n_samples, int64_t f0, int64_t f1, int64_t *begin, int64_t *end,
                    int32_t *lead);
void mfcc_batch_destroy(mfcc_batch *batch);
int64_t mfcc_batch_total_frames(const mfcc_batch *batch);
int64_t mfcc_batch_total_samples(const mfcc_batch *batch);
/* frame_offsets[u] = first output row of utterance u; n_utts + 1 entries. */
int mfcc_batch_frame_offsets(const mfcc_batch *batch, int64_t *h_frame_offsets);

/* THE HOT ENTRY.  Device pointers, asynchronous on `cuda_stream` (a
 * cudaStream_t passed as void*, NULL = default stream):
 *   d_pcm : int16, total_samples elements, device memory
 *   d_out : f32, total_frames * out_dim elements, device memory, row = frame
 * Allocates nothing; launches only kernels.  Reentrant for distinct streams. */
int mfcc_compute_batch(const mfcc_plan *plan, const mfcc_batch *batch,
                       const int16_t *d_pcm, float *d_out, void *cuda_stream);
/* Same for f32 PCM already scaled to int16 range (value = sample * 32768). */
int mfcc_compute_batch_f32(const mfcc_plan *plan, const mfcc_batch *batch,
                           const float *d_pcm, float *d_out, void *cuda_stream);

/* Same for G.711 codes (1 byte per sample, alaw = 0: mu-law, != 0: A-law): the codes are expanded inside the fused
 * kernel's staging, so telephony audio moves 1 byte per sample over PCIe and from HBM and never exists as int16
 * (SURVEY.md §8f rank 3).  Values equal mfcc_decode_g711 followed by mfcc_compute_batch, bit for bit. */
int mfcc_compute_batch_g711(const mfcc_plan *plan, const mfcc_batch *batch, const uint8_t *d_codes,
                            int32_t alaw, float *d_out, void *cuda_stream);

/* End-to-end call with HOST buffers: H2D copy of the PCM, the kernels, D2H
 * copy of the features, pipelined in chunks over internal streams; returns
 * after the features are in h_out.  Pinned host buffers (mfcc_host_alloc)
 * make the copies asynchronous.  h_frame_offsets may be NULL. */
int mfcc_compute_host(mfcc_plan *plan, const int16_t *h_pcm, const int64_t *h_offsets,
                      int64_t n_utts, float *h_out, int64_t *h_frame_offsets);

/* mfcc_compute_host for G.711 codes in host memory. */
int mfcc_compute_host_g711(mfcc_plan *plan, const uint8_t *h_codes, int32_t alaw, const int64_t *h_offsets,
                           int64_t n_utts, float *h_out, int64_t *h_frame_offsets);

/* Single-clip convenience (the shape a C caller of an embedded MFCC routine
 * expects): n samples in, *n_frames rows of out_dim floats out. */
int mfcc_compute(mfcc_plan *plan, const int16_t *pcm, int64_t n_samples,
                 float *out, int64_t *n_frames);

/* Post-processing on the feature matrix on the device, one step at a time (the same kernels as
 * mfcc_post_batch below, which does all of it in one pass and is the one to use when the stacked matrix is wanted):
 * mfcc_cmvn_batch : per-utterance cepstral mean (norm_var != 0: and variance) normalisation of d_feat IN PLACE;
 * mfcc_delta_batch: delta regression over +-window frames (HTK formula, edge frames replicated) of d_feat written to
 *                   d_delta (same shape, must not alias d_feat; both required).  Delta-delta = a second call on the
 *                   first call's output. */
int mfcc_cmvn_batch(const mfcc_plan *plan, const mfcc_batch *batch, float *d_feat,
                    int32_t norm_var, void *cuda_stream);
int mfcc_delta_batch(const mfcc_plan *plan, const mfcc_batch *batch, const float *d_feat,
                     int32_t window, float *d_delta, void *cuda_stream);

/* The same post-processing FUSED (mfcc_b200/csrc/mfcc_post.cu): per-utterance CMVN, then delta and
 * delta-delta regression, written once as the stacked matrix acoustic models consume,
 *     d_out [total_frames][out_dim * (1 + delta_order)] = static | delta | delta-delta      (e.g. 13 -> 39 columns),
 * equal to mfcc_cmvn_batch followed by mfcc_delta_batch (twice) and a concatenation.  cmvn: MFCC_CMVN_*; delta_order
 * 0, 1 or 2; delta_window 1..8 (ignored when delta_order == 0).  Asynchronous on the stream, allocates nothing (the
 * statistics scratch belongs to the batch, so calls on ONE batch must be stream-ordered); one launch, three with CMVN (statistics, finalise, apply).
 * d_out must not overlap d_feat.  HBM-bound: reads out_dim floats per frame (twice with CMVN), writes the stacked row. */
enum { MFCC_CMVN_NONE = 0, MFCC_CMVN_MEAN = 1, MFCC_CMVN_MEAN_VAR = 2 };
int mfcc_post_batch(const mfcc_plan *plan, const mfcc_batch *batch, const float *d_feat, int32_t cmvn,
                    int32_t delta_window, int32_t delta_order, float *d_out, void *cuda_stream);

/* mfcc_compute_host with the fused post-processing between the transform kernel and the read-back: int16 PCM in host
 * memory in, h_out [total_frames][out_dim * (1 + delta_order)] (static | delta | delta-delta, per-utterance CMVN) out.
 * Same pipeline, same serialisation per plan; the link carries out_dim * (1 + delta_order) floats per frame back. */
int mfcc_compute_host_post(mfcc_plan *plan, const int16_t *h_pcm, const int64_t *h_offsets, int64_t n_utts, int32_t cmvn,
                           int32_t delta_window, int32_t delta_order, float *h_out, int64_t *h_frame_offsets);

/* Input format widening: G.711 mu-law / A-law bytes to
 * int16 PCM on the device, elementwise, async on the stream. */
int mfcc_decode_g711(const uint8_t *d_src, int64_t n, int32_t alaw, int16_t *d_dst,
                     void *cuda_stream);

/* RIFF/WAVE header parsing on the host.  `data` is the whole file (or at least its header and
 * as much of the data chunk as is present) in host memory; nothing is copied.  On MFCC_OK, info says which device
 * entry takes the samples (MFCC_WAV_PCM16 -> mfcc_compute_batch, MFCC_WAV_MULAW / MFCC_WAV_ALAW ->
 * mfcc_compute_batch_g711, MFCC_WAV_F32 -> mfcc_compute_batch_f32 after scaling by 32768) and where they are:
 * data_offset bytes into the file, n_frames sample frames of `channels` interleaved samples.  A data chunk whose
 * size field is 0 / 0xFFFFFFFF (streamed) or larger than the file (truncated) is taken to run to the end of `data`.
 * MFCC_EINVAL: not a RIFF/WAVE file or malformed; MFCC_ENOTSUP: a sample format without a device entry
 * (8 / 24 / 32-bit integer PCM, ADPCM, f64). */
#define MFCC_WAV_PCM16  1
#define MFCC_WAV_MULAW  2
#define MFCC_WAV_ALAW   3
#define MFCC_WAV_F32    4
typedef struct mfcc_wav_info {
    int32_t format;           /* MFCC_WAV_* */
    int32_t channels;
    int32_t sample_rate;
    int32_t bits_per_sample;
    int64_t data_offset;      /* byte offset of the first sample in the file */
    int64_t data_bytes;       /* n_frames * channels * bytes per sample */
    int64_t n_frames;         /* samples per channel */
} mfcc_wav_info;
int mfcc_wav_parse(const void *data, int64_t bytes, mfcc_wav_info *info);

/* Streaming / online front end.  One mfcc_stream is ONE audio stream fed in
 * chunks of any size (host int16 PCM).  Row t it returns is bit-identical to row t of mfcc_compute over
 * the concatenation of everything fed so far: the object carries the frame_len - hop_len unconsumed
 * samples plus one sample of pre-emphasis history between calls, nothing else.
 *   mfcc_stream_feed  : append n samples, write every frame that is now complete to `out`
 *                       (capacity max_frames rows of out_dim floats; MFCC_EINVAL if too small — size it with
 *                       mfcc_stream_pending), *n_frames = rows written.  One H2D, one kernel, one D2H.
 *   mfcc_stream_flush : end of stream.  Under MFCC_PAD_ZERO_TAIL emits the zero-padded tail frame(s);
 *                       under MFCC_PAD_NONE emits nothing.  Resets the stream for reuse.
 *   mfcc_stream_pending: frames a feed of n_new more samples would return (n_new = 0 with
 *                       at_flush != 0: frames a flush would return). */
typedef struct mfcc_stream mfcc_stream;
int mfcc_stream_create(mfcc_plan *plan, mfcc_stream **out);
void mfcc_stream_destroy(mfcc_stream *stream);
int64_t mfcc_stream_pending(const mfcc_stream *stream, int64_t n_new, int32_t at_flush);
int mfcc_stream_feed(mfcc_stream *stream, const int16_t *pcm, int64_t n, float *out, int64_t max_frames,
                     int64_t *n_frames);
int mfcc_stream_flush(mfcc_stream *stream, float *out, int64_t max_frames, int64_t *n_frames);
/* Serving form: n_streams live streams of ONE plan fed in one call — stream i gets n[i] samples from pcm[i] and writes
 * its complete frames to out[i] (capacity max_frames[i] rows; NULL allowed when nothing is due), n_frames[i] = rows
 * written.  The streams' pending samples are packed into one pinned staging array: ONE host->device copy, ONE kernel
 * launch over all streams' tiles, ONE device->host copy, whatever n_streams is.  Row for row identical to feeding the
 * streams one by one.  The call is all-or-nothing: on MFCC_EINVAL no stream has been fed. */
int mfcc_stream_feed_many(mfcc_stream *const *streams, int64_t n_streams, const int16_t *const *pcm, const int64_t *n,
                          float *const *out, const int64_t *max_frames, int64_t *n_frames);

/* Pinned host memory helpers for the end-to-end path. */
int mfcc_host_alloc(void **ptr, int64_t bytes);
int mfcc_host_free(void *ptr);

/* Kernel launches issued by this library in this process (bench.py's gpu_launches). */
uint64_t mfcc_launch_count(void);

const char *mfcc_strerror(int err);
const char *mfcc_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MFCC_B200_H_ */
