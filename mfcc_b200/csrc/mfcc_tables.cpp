// mfcc_tables.cpp — parameter validation, framing arithmetic and the host
// table builder of libmfcc_b200.so.
//
// Nothing here follows reference code: simotin13/mfcc has no MFCC tables
// (SURVEY.md §0.2 — no <math.h> anywhere in src/mfcc).  The conventions are
// SURVEY.md §8(a) "PROPOSED", restated in DESIGN.md "Spec".  Every table is
// evaluated in double and rounded ONCE to f32, which is what the kernels read.
#include <cmath>
#include <cstring>

#include "mfcc_host.h"

namespace mfcc {

int validate_params(const mfcc_params *p)
{
    if (p == nullptr) return MFCC_EINVAL;
    if (p->sample_rate <= 0 || p->frame_len <= 0 || p->hop_len <= 0) return MFCC_EINVAL;
    const bool pow2 = p->nfft > 0 && (p->nfft & (p->nfft - 1)) == 0;
    if (!pow2 || p->nfft < 8 || p->nfft > 4096 || p->frame_len > p->nfft) return MFCC_EINVAL;
    if (p->n_mel < 1 || p->n_mel > 128 || p->n_cep < 1 || p->n_cep > p->n_mel) return MFCC_EINVAL;
    switch (p->window) {
        case MFCC_WINDOW_RECT: case MFCC_WINDOW_HAMMING: case MFCC_WINDOW_HANN: break;
        default: return MFCC_EINVAL;
    }
    if (p->pad_mode != MFCC_PAD_NONE && p->pad_mode != MFCC_PAD_ZERO_TAIL) return MFCC_EINVAL;
    if (p->output != MFCC_OUT_CEPSTRA && p->output != MFCC_OUT_LOGMEL) return MFCC_EINVAL;
    if (p->energy < MFCC_ENERGY_NONE || p->energy > MFCC_ENERGY_APPEND) return MFCC_EINVAL;
    if (p->energy == MFCC_ENERGY_REPLACE_C0 && p->output != MFCC_OUT_CEPSTRA) return MFCC_EINVAL;
    if (p->lifter < 0) return MFCC_EINVAL;
    if (!(p->log_floor > 0.0f) || !std::isfinite(p->log_floor)) return MFCC_EINVAL;
    if (!(p->preemph >= 0.0f && p->preemph <= 1.0f)) return MFCC_EINVAL;
    const double nyq = 0.5 * static_cast<double>(p->sample_rate);
    const double hi = p->f_hi > 0.0f ? static_cast<double>(p->f_hi) : nyq;
    if (!(p->f_lo >= 0.0f) || !(static_cast<double>(p->f_lo) < hi) || hi > nyq) return MFCC_EINVAL;
    return MFCC_OK;
}

static inline double mel_of_hz(double hz) { return 2595.0 * std::log10(1.0 + hz / 700.0); }
static inline double hz_of_mel(double mel) { return 700.0 * (std::pow(10.0, mel / 2595.0) - 1.0); }

int build_tables(const mfcc_params &p, HostTables &t)
{
    if (validate_params(&p) != MFCC_OK) return MFCC_EINVAL;
    const int L = p.frame_len, N = p.nfft, M = p.n_mel, nb = N / 2 + 1;
    t.nbins = nb;
    t.out_dim = (p.output == MFCC_OUT_LOGMEL ? M : p.n_cep) + (p.energy == MFCC_ENERGY_APPEND ? 1 : 0);

    // Window.
    t.window.resize(L);
    const double a0 = p.window == MFCC_WINDOW_HAMMING ? 0.54 : 0.5;
    for (int n = 0; n < L; ++n) {
        double w = 1.0;
        if (p.window != MFCC_WINDOW_RECT && L > 1)
            w = a0 - (1.0 - a0) * std::cos(2.0 * M_PI * n / static_cast<double>(L - 1));
        t.window[n] = static_cast<float>(w);
    }

    // Mel edge bins: M + 2 points equally spaced in mel, floor((N+1) f / sr).
    t.mel_bins.resize(M + 2);
    const double hi_hz = p.f_hi > 0.0f ? p.f_hi : 0.5 * p.sample_rate;
    const double m_lo = mel_of_hz(p.f_lo), m_hi = mel_of_hz(hi_hz);
    for (int i = 0; i < M + 2; ++i) {
        const double hz = hz_of_mel(m_lo + (m_hi - m_lo) * i / static_cast<double>(M + 1));
        long b = static_cast<long>(std::floor((N + 1) * hz / static_cast<double>(p.sample_rate)));
        if (b < 0) b = 0;
        if (b > N / 2) b = N / 2;
        t.mel_bins[i] = static_cast<int32_t>(b);
    }

    // Dense triangles and their segment form.  For bin k in segment j =
    // [bins[j], bins[j+1]) the rising weight belongs to filter j and the
    // falling weight (1 - rise) to filter j-1; bins outside every segment get 0.
    t.mel_w.assign(static_cast<size_t>(M) * nb, 0.0f);
    t.rise.assign(nb, 0.0f);
    t.fall.assign(nb, 0.0f);
    for (int j = 0; j <= M; ++j) {
        const int lo = t.mel_bins[j], hi = t.mel_bins[j + 1];
        for (int k = lo; k < hi; ++k) {
            const double up = static_cast<double>(k - lo) / static_cast<double>(hi - lo);
            const double dn = static_cast<double>(hi - k) / static_cast<double>(hi - lo);
            if (j < M) {
                t.mel_w[static_cast<size_t>(j) * nb + k] = static_cast<float>(up);
                t.rise[k] = static_cast<float>(up);
            }
            if (j > 0) {
                t.mel_w[static_cast<size_t>(j - 1) * nb + k] = static_cast<float>(dn);
                t.fall[k] = static_cast<float>(dn);
            }
        }
    }

    // Orthonormal DCT-II rows with the lifter folded in.
    t.dct.resize(static_cast<size_t>(p.n_cep) * M);
    for (int k = 0; k < p.n_cep; ++k) {
        const double scale = std::sqrt((k == 0 ? 1.0 : 2.0) / M);
        const double lift =
            p.lifter > 0 ? 1.0 + 0.5 * p.lifter * std::sin(M_PI * k / static_cast<double>(p.lifter)) : 1.0;
        for (int m = 0; m < M; ++m)
            t.dct[static_cast<size_t>(k) * M + m] =
                static_cast<float>(lift * scale * std::cos(M_PI * k * (m + 0.5) / M));
    }

    // Radix-2 twiddles for the generic kernel.
    t.tw_re.resize(N / 2);
    t.tw_im.resize(N / 2);
    for (int k = 0; k < N / 2; ++k) {
        const double ang = 2.0 * M_PI * k / static_cast<double>(N);
        t.tw_re[k] = static_cast<float>(std::cos(ang));
        t.tw_im[k] = static_cast<float>(-std::sin(ang));
    }
    return MFCC_OK;
}

}  // namespace mfcc

extern "C" {

int mfcc_params_init(mfcc_params *p, int32_t sample_rate)
{
    if (p == nullptr || sample_rate <= 0) return MFCC_EINVAL;
    std::memset(p, 0, sizeof(*p));
    p->sample_rate = sample_rate;
    p->frame_len = static_cast<int32_t>((static_cast<int64_t>(sample_rate) * 25) / 1000);
    p->hop_len = static_cast<int32_t>((static_cast<int64_t>(sample_rate) * 10) / 1000);
    if (p->frame_len < 1) p->frame_len = 1;
    if (p->hop_len < 1) p->hop_len = 1;
    int n = 8;
    while (n < p->frame_len) n <<= 1;
    p->nfft = n;
    p->n_mel = 26;
    p->n_cep = 13;
    p->preemph = 0.97f;
    p->window = MFCC_WINDOW_HAMMING;
    p->f_lo = 0.0f;
    p->f_hi = 0.0f;
    p->log_floor = 1e-10f;
    p->lifter = 0;
    p->pad_mode = MFCC_PAD_NONE;
    p->output = MFCC_OUT_CEPSTRA;
    p->energy = MFCC_ENERGY_NONE;
    return mfcc::validate_params(p);
}

int mfcc_params_validate(const mfcc_params *p) { return mfcc::validate_params(p); }

int64_t mfcc_num_frames(const mfcc_params *p, int64_t n)
{
    if (mfcc::validate_params(p) != MFCC_OK || n < 0) return MFCC_EINVAL;
    const int64_t L = p->frame_len, H = p->hop_len;
    if (p->pad_mode == MFCC_PAD_NONE) return n < L ? 0 : 1 + (n - L) / H;
    if (n == 0) return 0;
    return n <= L ? 1 : 1 + (n - L + H - 1) / H;
}

int mfcc_piece_span(const mfcc_params *p, int64_t n, int64_t f0, int64_t f1, int64_t *begin, int64_t *end, int32_t *lead)
{
    const int64_t total = mfcc_num_frames(p, n);
    if (total < 0 || f0 < 0 || f1 < f0 || f1 > total || begin == nullptr || end == nullptr || lead == nullptr) return MFCC_EINVAL;
    const int64_t L = p->frame_len, H = p->hop_len;
    *lead = f0 > 0 ? 1 : 0;
    if (f1 == f0) {                 // no frames: an empty span
        *begin = *end = 0;
        *lead = 0;
        return MFCC_OK;
    }
    *begin = f0 * H - *lead;
    // every piece but the last ends with its last frame; the last one keeps the rest of the recording (what a zero-padded
    // last frame needs under MFCC_PAD_ZERO_TAIL; under MFCC_PAD_NONE the samples past the last frame are never read)
    *end = f1 == total ? n : (f1 - 1) * H + L;
    return MFCC_OK;
}

int32_t mfcc_out_dim(const mfcc_params *p)
{
    if (mfcc::validate_params(p) != MFCC_OK) return MFCC_EINVAL;
    return (p->output == MFCC_OUT_LOGMEL ? p->n_mel : p->n_cep) + (p->energy == MFCC_ENERGY_APPEND ? 1 : 0);
}

const char *mfcc_strerror(int err)
{
    switch (err) {
        case MFCC_OK: return "ok";
        case MFCC_EINVAL: return "invalid argument";
        case MFCC_ENOMEM: return "out of memory";
        case MFCC_ECUDA: return "CUDA error (or no sm_100 device)";
        case MFCC_ENOTSUP: return "not supported by this build";
        default: return "unknown error";
    }
}

const char *mfcc_version(void) { return "mfcc_b200 0.1.0 (sm_100a)"; }

}  // extern "C"
