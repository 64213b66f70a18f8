"""ctypes mirror of ``mfcc_params`` (include/mfcc_b200.h) and the named configs.

The struct is the parameter list BASELINE.json's north_star names (sample rate,
frame length/hop, pre-emphasis, window, FFT size, mel-band count, cepstral
count); the reference has no such struct (src/mfcc/mfcc.h:1-22 only holds
STRING_MAX / CODE_LEN_MAX / BuildTargetType — SURVEY.md §0.2).
"""
from __future__ import annotations

import ctypes as C

MFCC_OK, MFCC_EINVAL, MFCC_ENOMEM, MFCC_ECUDA, MFCC_ENOTSUP = 0, -1, -2, -3, -4
WINDOW_RECT, WINDOW_HAMMING, WINDOW_HANN = 0, 1, 2
PAD_NONE, PAD_ZERO_TAIL = 0, 1
OUT_CEPSTRA, OUT_LOGMEL = 0, 1
KERNEL_AUTO, KERNEL_GENERIC, KERNEL_FUSED = 0, 1, 2
ENERGY_NONE, ENERGY_REPLACE_C0, ENERGY_APPEND = 0, 1, 2


class MfccParams(C.Structure):
    _fields_ = [
        ("sample_rate", C.c_int32),
        ("frame_len", C.c_int32),
        ("hop_len", C.c_int32),
        ("nfft", C.c_int32),
        ("n_mel", C.c_int32),
        ("n_cep", C.c_int32),
        ("preemph", C.c_float),
        ("window", C.c_int32),
        ("f_lo", C.c_float),
        ("f_hi", C.c_float),
        ("log_floor", C.c_float),
        ("lifter", C.c_int32),
        ("pad_mode", C.c_int32),
        ("output", C.c_int32),
        ("energy", C.c_int32),
    ]

    def copy(self, **kw) -> "MfccParams":
        q = MfccParams.from_buffer_copy(bytes(self))
        for k, v in kw.items():
            if not hasattr(q, k):
                raise AttributeError(k)
            setattr(q, k, v)
        return q

    @property
    def out_dim(self) -> int:
        return (self.n_mel if self.output == OUT_LOGMEL else self.n_cep) + (1 if self.energy == ENERGY_APPEND else 0)

    def as_dict(self) -> dict:
        return {k: getattr(self, k) for k, _ in self._fields_}


def make_params(sample_rate=16000, frame_len=400, hop_len=160, nfft=512, n_mel=26, n_cep=13,
                preemph=0.97, window=WINDOW_HAMMING, f_lo=0.0, f_hi=0.0, log_floor=1e-10,
                lifter=0, pad_mode=PAD_NONE, output=OUT_CEPSTRA, energy=ENERGY_NONE) -> MfccParams:
    return MfccParams(sample_rate, frame_len, hop_len, nfft, n_mel, n_cep, preemph, window,
                      f_lo, f_hi, log_floor, lifter, pad_mode, output, energy)


def config_a() -> MfccParams:
    """BASELINE.json configs[0,1,4]: 16 kHz, 25 ms / 10 ms, 512-pt FFT, 26 mel, 13 MFCC."""
    return make_params()


def config_b() -> MfccParams:
    """BASELINE.json configs[2]: 8 kHz telephony, 256-pt FFT, 20 mel, 13 MFCC."""
    return make_params(sample_rate=8000, frame_len=200, hop_len=80, nfft=256, n_mel=20, n_cep=13)


def config_c() -> MfccParams:
    """BASELINE.json configs[3]: 48 kHz wideband, 2048-pt FFT, 80 mel, 40 cepstra."""
    return make_params(sample_rate=48000, frame_len=1200, hop_len=480, nfft=2048, n_mel=80,
                       n_cep=40)


CONFIGS = {"A": config_a, "B": config_b, "C": config_c}
