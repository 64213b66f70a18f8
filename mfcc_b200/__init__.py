"""mfcc_b200 — B200-native MFCC front end (host-side mirror of include/mfcc_b200.h).

Importing this package never touches the GPU or the shared library; the
library is loaded on first use by :mod:`mfcc_b200.api` and fails loudly when
it is missing (there is no CPU fallback).
"""
from .params import (MfccParams, make_params, config_a, config_b, config_c, CONFIGS,  # noqa: F401
                     WINDOW_RECT, WINDOW_HAMMING, WINDOW_HANN, PAD_NONE, PAD_ZERO_TAIL,
                     OUT_CEPSTRA, OUT_LOGMEL, KERNEL_AUTO, KERNEL_GENERIC, KERNEL_FUSED,
                     ENERGY_NONE, ENERGY_REPLACE_C0, ENERGY_APPEND)

__version__ = "0.1.0"
