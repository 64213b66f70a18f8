// mfcc_wav.cpp — RIFF/WAVE header parsing on the host ("WAV header parsing on host").
// Pure host code, no CUDA: finds the `fmt ` and `data` chunks, classifies the sample format into the four the
// library has a device entry for (int16 PCM -> mfcc_compute_batch, G.711 mu-law / A-law -> mfcc_compute_batch_g711,
// IEEE f32 -> mfcc_compute_batch_f32) and reports where the samples are.  Nothing is copied or converted here.
//
// The reference has no audio I/O of any kind (its only file access is the fopen/fread of the C source it compiles,
// /root/reference/src/mfcc/main.c:117-135); the error convention is the library's (int, 0 ok, negative failure).
#include <cstdint>
#include <cstring>

#include "../../include/mfcc_b200.h"

namespace {

inline uint32_t le32(const unsigned char *p) { return p[0] | (p[1] << 8) | (p[2] << 16) | (static_cast<uint32_t>(p[3]) << 24); }
inline uint16_t le16(const unsigned char *p) { return static_cast<uint16_t>(p[0] | (p[1] << 8)); }

}  // namespace

extern "C" int mfcc_wav_parse(const void *data, int64_t bytes, mfcc_wav_info *info)
{
    if (data == nullptr || info == nullptr || bytes < 0) return MFCC_EINVAL;
    std::memset(info, 0, sizeof(*info));
    const unsigned char *b = static_cast<const unsigned char *>(data);
    if (bytes < 12 || std::memcmp(b, "RIFF", 4) != 0 || std::memcmp(b + 8, "WAVE", 4) != 0) return MFCC_EINVAL;
    bool have_fmt = false;
    uint16_t tag = 0, channels = 0, bits = 0, block = 0;
    uint32_t rate = 0;
    int64_t pos = 12;
    while (pos + 8 <= bytes) {
        const unsigned char *ck = b + pos;
        const int64_t size = le32(ck + 4);
        const int64_t body = pos + 8;
        if (std::memcmp(ck, "fmt ", 4) == 0) {
            if (size < 16 || body + 16 > bytes) return MFCC_EINVAL;
            tag = le16(b + body);
            channels = le16(b + body + 2);
            rate = le32(b + body + 4);
            block = le16(b + body + 12);
            bits = le16(b + body + 14);
            if (tag == 0xFFFE) {   // WAVE_FORMAT_EXTENSIBLE: the real tag is the first two bytes of the sub-format GUID
                if (size < 40 || body + 40 > bytes) return MFCC_EINVAL;
                tag = le16(b + body + 24);
            }
            have_fmt = true;
        } else if (std::memcmp(ck, "data", 4) == 0) {
            if (!have_fmt) return MFCC_EINVAL;   // the format must be known before the samples
            int64_t n = size;
            // streamed files leave the size at 0 or 0xFFFFFFFF, truncated files claim more than is there: take what exists
            if (n == 0 || n == 0xFFFFFFFFll || body + n > bytes) n = bytes - body;
            int32_t format = 0, bytes_per_sample = 0;
            if (tag == 1 && bits == 16) { format = MFCC_WAV_PCM16; bytes_per_sample = 2; }
            else if (tag == 7 && bits == 8) { format = MFCC_WAV_MULAW; bytes_per_sample = 1; }
            else if (tag == 6 && bits == 8) { format = MFCC_WAV_ALAW; bytes_per_sample = 1; }
            else if (tag == 3 && bits == 32) { format = MFCC_WAV_F32; bytes_per_sample = 4; }
            else return MFCC_ENOTSUP;           // 8 / 24 / 32-bit integer PCM, ADPCM, f64, ...
            if (channels == 0 || rate == 0 || block != channels * bytes_per_sample) return MFCC_EINVAL;
            info->format = format;
            info->channels = channels;
            info->sample_rate = static_cast<int32_t>(rate);
            info->bits_per_sample = bits;
            info->data_offset = body;
            info->n_frames = n / block;
            info->data_bytes = info->n_frames * block;
            return MFCC_OK;
        }
        pos = body + size + (size & 1);         // chunks are word-aligned
    }
    return MFCC_EINVAL;                         // no data chunk
}
