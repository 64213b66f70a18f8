/*
 * mfcc_oracle.h — prototypes of the CPU oracle (TEST INFRASTRUCTURE; see
 * mfcc_oracle.c for the scope and the "parity unpinned" statement).
 */
#ifndef MFCC_ORACLE_H_
#define MFCC_ORACLE_H_
#include <stdint.h>
#include "../include/mfcc_b200.h"
#ifdef __cplusplus
extern "C" {
#endif
int     oracle_params_validate(const mfcc_params *p);
int64_t oracle_num_frames(const mfcc_params *p, int64_t n);
int     oracle_window_f64(const mfcc_params *p, double *w);
int     oracle_mel_bins(const mfcc_params *p, int *bins);
int     oracle_mel_weights_f64(const mfcc_params *p, double *W);
int     oracle_dct_f64(const mfcc_params *p, double *D);
int64_t oracle_mfcc_f32(const mfcc_params *p, const int16_t *pcm, int64_t n, float *out);
int64_t oracle_mfcc_f64(const mfcc_params *p, const int16_t *pcm, int64_t n, double *out);
int     oracle_stages_f32(const mfcc_params *p, const int16_t *pcm, int64_t n, int64_t frame,
                          float *framed, float *power, float *mel, float *out);
int     oracle_stages_f64(const mfcc_params *p, const int16_t *pcm, int64_t n, int64_t frame,
                          double *framed, double *power, double *mel, double *out);
int64_t oracle_mfcc_batch_f32(const mfcc_params *p, const int16_t *pcm, const int64_t *offsets,
                              int64_t n_utts, float *out, int64_t *frame_offsets, int nthreads);
/* the CPU baseline of bench.py (mfcc_cpu_fast.c): same contract as oracle_mfcc_batch_f32, real-input FFT, -O3 / AVX2 */
int64_t oracle_fast_mfcc_batch_f32(const mfcc_params *p, const int16_t *pcm, const int64_t *offsets,
                                   int64_t n_utts, float *out, int64_t *frame_offsets, int nthreads);
int     oracle_cmvn_f32(float *feat, const int64_t *frame_offsets, int64_t n_utts, int dim,
                        int norm_var);
int     oracle_delta_f32(const float *feat, const int64_t *frame_offsets, int64_t n_utts, int dim,
                         int window, float *delta);
int16_t oracle_ulaw_decode(uint8_t b);
int16_t oracle_alaw_decode(uint8_t b);
int     oracle_decode_g711(const uint8_t *src, int64_t n, int alaw, int16_t *dst);
#ifdef __cplusplus
}
#endif
#endif
