/* A plain C99 caller of libmfcc_b200.so — the binding INTEGRATION.md §3 describes, kept compilable.
 * Host-only calls are exercised everywhere; the device part runs only when a plan can be created
 * (MFCC_ECUDA without an sm_100 GPU: the library has no CPU fallback). */
#include <stdio.h>
#include <stdlib.h>
#include "mfcc_b200.h"

int main(void)
{
    mfcc_params p;
    if (mfcc_params_init(&p, 16000) != MFCC_OK) return 1;
    if (p.frame_len != 400 || p.hop_len != 160 || p.nfft != 512 || p.n_mel != 26 || p.n_cep != 13) return 2;
    if (mfcc_num_frames(&p, 160000) != 998 || mfcc_num_frames(&p, 399) != 0 || mfcc_out_dim(&p) != 13) return 3;
    p.nfft = 500;                                   /* not a power of two */
    if (mfcc_params_validate(&p) != MFCC_EINVAL) return 4;
    p.nfft = 512;
    printf("%s | %s\n", mfcc_version(), mfcc_strerror(MFCC_ECUDA));

    mfcc_plan *plan = NULL;
    int rc = mfcc_plan_create(&p, 0, MFCC_KERNEL_AUTO, &plan);
    if (rc == MFCC_ECUDA) { printf("no device: %s\n", mfcc_strerror(rc)); return 0; }
    if (rc != MFCC_OK) return 5;

    enum { N = 16000 };
    int16_t *pcm = NULL;
    float *out = NULL;
    int64_t nf = mfcc_num_frames(&p, N), got = 0;
    if (mfcc_host_alloc((void **)&pcm, sizeof(int16_t) * N) != MFCC_OK) return 6;
    if (mfcc_host_alloc((void **)&out, sizeof(float) * (size_t)nf * 13) != MFCC_OK) return 7;
    for (int i = 0; i < N; ++i) pcm[i] = (int16_t)((i * 7919) % 2001 - 1000);
    rc = mfcc_compute(plan, pcm, N, out, &got);
    printf("mfcc_compute: %s, %lld frames, c0[0] = %f, kernel %s\n", mfcc_strerror(rc), (long long)got, out[0],
           mfcc_plan_kernel_name(plan));
    if (rc != MFCC_OK || got != nf) return 8;

    /* the same clip with per-utterance CMVN + delta + delta-delta (39 columns back): the mean-normalised static part
     * sums to zero over the utterance, column by column */
    {
        const int64_t offsets[2] = {0, N};
        int64_t fo[2] = {0, 0};
        float *out3 = NULL;
        double sum0 = 0.0;
        if (mfcc_host_alloc((void **)&out3, sizeof(float) * (size_t)nf * 39) != MFCC_OK) return 9;
        rc = mfcc_compute_host_post(plan, pcm, offsets, 1, MFCC_CMVN_MEAN, 2, 2, out3, fo);
        for (int64_t t = 0; rc == MFCC_OK && t < nf; ++t) sum0 += out3[t * 39];
        printf("mfcc_compute_host_post: %s, %lld frames x 39, sum of normalised c0 = %.3g\n", mfcc_strerror(rc),
               (long long)fo[1], sum0);
        mfcc_host_free(out3);
        if (rc != MFCC_OK || fo[1] != nf || sum0 > 1e-2 || sum0 < -1e-2) return 10;
    }
    mfcc_host_free(pcm);
    mfcc_host_free(out);
    mfcc_plan_destroy(plan);
    return 0;
}
