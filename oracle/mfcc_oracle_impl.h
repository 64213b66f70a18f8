/*
 * mfcc_oracle_impl.h — body of the CPU oracle, included twice by
 * mfcc_oracle.c with REAL = float (suffix _f32) and REAL = double (_f64).
 *
 * TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (see mfcc_oracle.c header).
 *
 * Every stage is the textbook definition written as the most obvious loop;
 * the order of floating-point operations written here IS the parity target of
 * the CUDA path (tolerances in tests/, not bit-exactness, for the float math;
 * framing indices are bit-exact).  Reference citations: there is no MFCC code
 * in /root/reference to follow (SURVEY.md §0.2, §8a "Ref file:line = none");
 * each function cites the SURVEY.md §8(a) row that defines it instead.
 */

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUFFIX)

typedef struct {
    int nfft, nbins, n_mel, n_out, frame_len;
    REAL *window;   /* [frame_len] */
    REAL *tw_re;    /* [nfft/2] cos(2 pi k / nfft) */
    REAL *tw_im;    /* [nfft/2] -sin(2 pi k / nfft) */
    REAL *melw;     /* [n_mel][nbins] dense triangular weights */
    int  *bins;     /* [n_mel + 2] */
    REAL *dct;      /* [n_cep][n_mel] orthonormal DCT-II rows, lifter folded in */
    REAL *re, *im, *pw, *loge; /* scratch */
} FN(tables);

/* SURVEY.md §8(a) rows "Hamming window", "Mel filterbank", "DCT-II":
 * every table is evaluated in double and rounded once to REAL. */
static int FN(tables_init)(FN(tables) *t, const mfcc_params *p)
{
    memset(t, 0, sizeof(*t));
    t->nfft = p->nfft;
    t->nbins = p->nfft / 2 + 1;
    t->n_mel = p->n_mel;
    t->n_out = p->n_cep;
    t->frame_len = p->frame_len;
    t->window = (REAL *)malloc(sizeof(REAL) * (size_t)p->frame_len);
    t->tw_re = (REAL *)malloc(sizeof(REAL) * (size_t)(p->nfft / 2 + 1));
    t->tw_im = (REAL *)malloc(sizeof(REAL) * (size_t)(p->nfft / 2 + 1));
    t->melw = (REAL *)calloc((size_t)p->n_mel * (size_t)t->nbins, sizeof(REAL));
    t->bins = (int *)malloc(sizeof(int) * (size_t)(p->n_mel + 2));
    t->dct = (REAL *)malloc(sizeof(REAL) * (size_t)p->n_cep * (size_t)p->n_mel);
    t->re = (REAL *)malloc(sizeof(REAL) * (size_t)p->nfft);
    t->im = (REAL *)malloc(sizeof(REAL) * (size_t)p->nfft);
    t->pw = (REAL *)malloc(sizeof(REAL) * (size_t)t->nbins);
    t->loge = (REAL *)malloc(sizeof(REAL) * (size_t)p->n_mel);
    if (!t->window || !t->tw_re || !t->tw_im || !t->melw || !t->bins || !t->dct || !t->re ||
        !t->im || !t->pw || !t->loge)
        return MFCC_ENOMEM;

    double *wd = (double *)malloc(sizeof(double) * (size_t)p->frame_len);
    double *md = (double *)malloc(sizeof(double) * (size_t)p->n_mel * (size_t)t->nbins);
    double *dd = (double *)malloc(sizeof(double) * (size_t)p->n_cep * (size_t)p->n_mel);
    if (!wd || !md || !dd) { free(wd); free(md); free(dd); return MFCC_ENOMEM; }
    oracle_window_f64(p, wd);
    oracle_mel_bins(p, t->bins);
    oracle_mel_weights_f64(p, md);
    oracle_dct_f64(p, dd);
    for (int n = 0; n < p->frame_len; ++n) t->window[n] = (REAL)wd[n];
    for (size_t i = 0; i < (size_t)p->n_mel * (size_t)t->nbins; ++i) t->melw[i] = (REAL)md[i];
    for (size_t i = 0; i < (size_t)p->n_cep * (size_t)p->n_mel; ++i) t->dct[i] = (REAL)dd[i];
    for (int k = 0; k < p->nfft / 2; ++k) {
        double a = 2.0 * M_PI * (double)k / (double)p->nfft;
        t->tw_re[k] = (REAL)cos(a);
        t->tw_im[k] = (REAL)(-sin(a));
    }
    free(wd); free(md); free(dd);
    return 0;
}

static void FN(tables_free)(FN(tables) *t)
{
    free(t->window); free(t->tw_re); free(t->tw_im); free(t->melw); free(t->bins);
    free(t->dct); free(t->re); free(t->im); free(t->pw); free(t->loge);
    memset(t, 0, sizeof(*t));
}

/* SURVEY.md §8(a) row "Real FFT": X[k] = sum_n x[n] exp(-2 pi i k n / N).
 * In-place iterative radix-2 decimation-in-time on a full complex buffer
 * (imaginary input zero) — deliberately the plainest possible formulation. */
static void FN(fft_inplace)(const FN(tables) *t, REAL *re, REAL *im)
{
    const int n = t->nfft;
    for (int i = 1, j = 0; i < n; ++i) {          /* bit-reversal permutation */
        int bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) {
            REAL tr = re[i]; re[i] = re[j]; re[j] = tr;
            REAL ti = im[i]; im[i] = im[j]; im[j] = ti;
        }
    }
    for (int len = 2; len <= n; len <<= 1) {
        const int half = len >> 1, step = n / len;
        for (int base = 0; base < n; base += len) {
            for (int k = 0; k < half; ++k) {
                const REAL wr = t->tw_re[k * step], wi = t->tw_im[k * step];
                const int a = base + k, b = a + half;
                const REAL xr = re[b] * wr - im[b] * wi;
                const REAL xi = re[b] * wi + im[b] * wr;
                re[b] = re[a] - xr; im[b] = im[a] - xi;
                re[a] = re[a] + xr; im[a] = im[a] + xi;
            }
        }
    }
}

/* One frame, all stages (SURVEY.md §8(a), rows top to bottom).  `pcm` is the
 * whole utterance of n samples, `t0` the frame start.  Optional stage taps. */
static void FN(one_frame)(const mfcc_params *p, FN(tables) *t, const int16_t *pcm, int64_t n,
                          int64_t t0, REAL *out, REAL *tap_frame, REAL *tap_power,
                          REAL *tap_mel)
{
    const REAL a = (REAL)p->preemph;
    /* Framing + int16->REAL + pre-emphasis + window + zero-pad. */
    for (int i = 0; i < p->nfft; ++i) {
        REAL v = 0;
        const int64_t s = t0 + i;
        if (i < p->frame_len && s < n) {        /* s >= n only under MFCC_PAD_ZERO_TAIL */
            const REAL x0 = (REAL)pcm[s];
            const REAL x1 = (s > 0) ? (REAL)pcm[s - 1] : (REAL)0;
            const REAL y = x0 - a * x1;         /* y[0] = x[0] */
            v = y * t->window[i];
        }
        t->re[i] = v;
        t->im[i] = 0;
        if (tap_frame) tap_frame[i] = v;
    }
    FN(fft_inplace)(t, t->re, t->im);
    /* Power spectrum, scale 1/NFFT. */
    const REAL inv_n = (REAL)(1.0 / (double)p->nfft);
    for (int k = 0; k < t->nbins; ++k) {
        t->pw[k] = (t->re[k] * t->re[k] + t->im[k] * t->im[k]) * inv_n;
        if (tap_power) tap_power[k] = t->pw[k];
    }
    /* Mel energies: ascending-k dot product over each triangle's support. */
    const REAL flo = (REAL)p->log_floor;
    for (int m = 0; m < p->n_mel; ++m) {
        REAL e = 0;
        const REAL *w = t->melw + (size_t)m * (size_t)t->nbins;
        int k0 = t->bins[m], k1 = t->bins[m + 2];
        if (k1 > t->nbins - 1) k1 = t->nbins - 1;
        for (int k = k0; k <= k1; ++k) e += w[k] * t->pw[k];
        if (tap_mel) tap_mel[m] = e;
        t->loge[m] = (REAL)LOGFN(e > flo ? e : flo);
    }
    /* Frame energy term (MFCC_ENERGY_*): E = sum of the one-sided power spectrum, ascending k. */
    REAL log_energy = 0;
    if (p->energy != MFCC_ENERGY_NONE) {
        REAL e = 0;
        for (int k = 0; k < t->nbins; ++k) e += t->pw[k];
        log_energy = (REAL)LOGFN(e > flo ? e : flo);
    }
    if (p->output == MFCC_OUT_LOGMEL) {
        for (int m = 0; m < p->n_mel; ++m) out[m] = t->loge[m];
        if (p->energy == MFCC_ENERGY_APPEND) out[p->n_mel] = log_energy;
        return;
    }
    /* DCT-II (orthonormal; lifter already folded into the rows). */
    for (int k = 0; k < p->n_cep; ++k) {
        REAL c = 0;
        const REAL *d = t->dct + (size_t)k * (size_t)p->n_mel;
        for (int m = 0; m < p->n_mel; ++m) c += d[m] * t->loge[m];
        out[k] = c;
    }
    if (p->energy == MFCC_ENERGY_REPLACE_C0) out[0] = log_energy;
    if (p->energy == MFCC_ENERGY_APPEND) out[p->n_cep] = log_energy;
}

/* Whole utterance: returns the frame count (>= 0) or a negative error.
 * out is [n_frames][out_dim]. */
int64_t FN(oracle_mfcc)(const mfcc_params *p, const int16_t *pcm, int64_t n, REAL *out)
{
    if (oracle_params_validate(p) != 0 || n < 0 || (n > 0 && !pcm)) return MFCC_EINVAL;
    const int64_t nf = oracle_num_frames(p, n);
    if (nf <= 0) return nf;
    if (!out) return MFCC_EINVAL;
    FN(tables) t;
    int rc = FN(tables_init)(&t, p);
    if (rc != 0) { FN(tables_free)(&t); return rc; }
    const int od = ORACLE_OUT_DIM(p);
    for (int64_t f = 0; f < nf; ++f)
        FN(one_frame)(p, &t, pcm, n, f * (int64_t)p->hop_len, out + f * od, NULL, NULL, NULL);
    FN(tables_free)(&t);
    return nf;
}

/* Stage taps of a single frame, for the known-answer tests:
 * windowed frame [nfft], power [nfft/2+1], mel energies [n_mel], output [out_dim]. */
int FN(oracle_stages)(const mfcc_params *p, const int16_t *pcm, int64_t n, int64_t frame,
                      REAL *framed, REAL *power, REAL *mel, REAL *out)
{
    if (oracle_params_validate(p) != 0) return MFCC_EINVAL;
    const int64_t nf = oracle_num_frames(p, n);
    if (frame < 0 || frame >= nf) return MFCC_EINVAL;
    FN(tables) t;
    int rc = FN(tables_init)(&t, p);
    if (rc != 0) { FN(tables_free)(&t); return rc; }
    const int od = ORACLE_OUT_DIM(p);
    REAL *tmp = (REAL *)malloc(sizeof(REAL) * (size_t)od);
    FN(one_frame)(p, &t, pcm, n, frame * (int64_t)p->hop_len, out ? out : tmp, framed, power, mel);
    free(tmp);
    FN(tables_free)(&t);
    return 0;
}

#undef FN
#undef CAT
#undef CAT_
