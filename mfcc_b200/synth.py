"""Seeded synthetic PCM for every BASELINE.json config.

The reference ships no audio (test/test.c:1-5 is a hello-world program), so the
inputs are defined here once and used by tests, fixtures and bench.py alike.
Plain numpy on the host; bench.py copies to the device before timing.
"""
from __future__ import annotations

import numpy as np


def clip_config1(seconds: float = 1.0, sr: int = 16000, seed: int = 0) -> np.ndarray:
    """Config 1 clip: 440 Hz + 1 kHz + 3 kHz at -12 dBFS total + Gaussian noise, sigma 100 LSB."""
    n = int(round(seconds * sr))
    t = np.arange(n) / sr
    amp = 32768.0 * 10 ** (-12 / 20) / 3.0
    x = amp * (np.sin(2 * np.pi * 440 * t) + np.sin(2 * np.pi * 1000 * t) + np.sin(2 * np.pi * 3000 * t))
    x += np.random.default_rng(seed).normal(0.0, 100.0, n)
    return np.clip(np.rint(x), -32768, 32767).astype(np.int16)


def noise_utterance(n: int, seed: int, sigma: float = 3000.0) -> np.ndarray:
    """Configs 2/4/5 utterance: Gaussian sigma LSB, clipped to int16, one seed per utterance."""
    x = np.random.default_rng(seed).normal(0.0, sigma, n)
    return np.clip(np.rint(x), -32768, 32767).astype(np.int16)


def fixed_batch(n_utts: int, n_samples: int, seed0: int = 1000, sigma: float = 3000.0):
    """Config 2 shape: n_utts equal-length utterances, concatenated.  Returns (pcm, offsets)."""
    pcm = np.empty(n_utts * n_samples, np.int16)
    for u in range(n_utts):
        pcm[u * n_samples:(u + 1) * n_samples] = noise_utterance(n_samples, seed0 + u, sigma)
    offsets = np.arange(n_utts + 1, dtype=np.int64) * n_samples
    return pcm, offsets


def fast_fixed_batch(n_utts: int, n_samples: int, seed: int = 1000, sigma: float = 3000.0):
    """Same distribution as fixed_batch from ONE generator stream (for the full-size
    bench buffers, where 1,024 separate generators cost seconds, not for fixtures)."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal(n_utts * n_samples, dtype=np.float32)
    x *= np.float32(sigma)
    np.rint(x, out=x)
    np.clip(x, -32768, 32767, out=x)
    offsets = np.arange(n_utts + 1, dtype=np.int64) * n_samples
    return x.astype(np.int16), offsets


def ragged_batch(n_utts: int, min_len: int, max_len: int, seed: int = 3, sigma: float = 3000.0):
    """Config 3 shape: lengths uniform in [min_len, max_len], concatenated with offsets."""
    rng = np.random.default_rng(seed)
    lens = rng.integers(min_len, max_len + 1, n_utts).astype(np.int64)
    offsets = np.zeros(n_utts + 1, np.int64)
    np.cumsum(lens, out=offsets[1:])
    x = rng.standard_normal(int(offsets[-1]), dtype=np.float32) * np.float32(sigma)
    pcm = np.clip(np.rint(x), -32768, 32767).astype(np.int16)
    return pcm, offsets


HOSTILE_KINDS = ("silence_gaps", "dc_offset", "low_noise", "fullscale_tone", "clicks")


def hostile_clip(kind: str, n: int, sr: int, seed: int = 7) -> np.ndarray:
    """Inputs chosen to stress the places where a fast kernel and a plain loop can part ways:
    the log floor (digital silence inside speech-level noise), pre-emphasis against a DC offset, the low end of the
    int16 range (sigma = 3 LSB), FFT leakage under a full-scale tone, and flat spectra (single-sample clicks)."""
    rng = np.random.default_rng(seed)
    if kind == "silence_gaps":
        x = rng.normal(0.0, 3000.0, n)
        for lo, hi in ((n // 7, n // 7 + n // 5), (n // 2, n // 2 + n // 9), (n - n // 20, n)):
            x[lo:hi] = 0.0
    elif kind == "dc_offset":
        x = rng.normal(0.0, 1000.0, n) + 2000.0
    elif kind == "low_noise":
        x = rng.normal(0.0, 3.0, n)
    elif kind == "fullscale_tone":
        x = 32767.0 * np.sin(2 * np.pi * 1000.0 * np.arange(n) / sr)
    elif kind == "clicks":
        x = np.zeros(n)
        x[rng.integers(0, n, max(n // 900, 3))] = rng.choice([-20000.0, 15000.0, 32767.0], max(n // 900, 3))
    else:
        raise ValueError(kind)
    return np.clip(np.rint(x), -32768, 32767).astype(np.int16)
