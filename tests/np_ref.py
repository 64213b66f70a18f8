"""Independent numpy/scipy float64 restatement of the MFCC spec (DESIGN.md "Spec").

It shares NO code with oracle/mfcc_oracle.c: vectorised framing, numpy.fft.rfft,
scipy.fft.dct.  Used to (a) check the oracle's double build to ~1e-9 and (b)
generate the fixtures in tests/golden/ (tests/golden/make_golden.py).

The reference (simotin13/mfcc) has no MFCC code to restate;
the conventions are SURVEY.md §8(a) "PROPOSED".
"""
from __future__ import annotations

import numpy as np
import scipy.fft


def num_frames(p, n: int) -> int:
    if p.pad_mode == 0:
        return 0 if n < p.frame_len else 1 + (n - p.frame_len) // p.hop_len
    if n == 0:
        return 0
    if n <= p.frame_len:
        return 1
    return 1 + -(-(n - p.frame_len) // p.hop_len)


def window(p) -> np.ndarray:
    L = p.frame_len
    if p.window == 0 or L == 1:
        return np.ones(L)
    n = np.arange(L)
    c = np.cos(2 * np.pi * n / (L - 1))
    return 0.54 - 0.46 * c if p.window == 1 else 0.5 - 0.5 * c


def mel_bins(p) -> np.ndarray:
    fhi = p.f_hi if p.f_hi > 0 else p.sample_rate / 2
    mel = lambda f: 2595.0 * np.log10(1.0 + f / 700.0)
    pts = np.linspace(mel(float(p.f_lo)), mel(float(fhi)), p.n_mel + 2)
    hz = 700.0 * (10.0 ** (pts / 2595.0) - 1.0)
    return np.clip(np.floor((p.nfft + 1) * hz / p.sample_rate), 0, p.nfft // 2).astype(np.int64)


def mel_weights(p) -> np.ndarray:
    b = mel_bins(p)
    W = np.zeros((p.n_mel, p.nfft // 2 + 1))
    for m in range(p.n_mel):
        lo, ce, hi = b[m], b[m + 1], b[m + 2]
        if ce > lo:
            W[m, lo:ce] = (np.arange(lo, ce) - lo) / (ce - lo)
        if hi > ce:
            W[m, ce:hi] = (hi - np.arange(ce, hi)) / (hi - ce)
    return W


def mfcc(p, pcm: np.ndarray) -> np.ndarray:
    x = np.asarray(pcm, np.float64)
    n = x.size
    nf = num_frames(p, n)
    if nf == 0:
        return np.zeros((0, (p.n_mel if p.output == 1 else p.n_cep) + (getattr(p, "energy", 0) == 2)))
    y = x.copy()
    y[1:] -= float(np.float32(p.preemph)) * x[:-1]
    need = (nf - 1) * p.hop_len + p.frame_len
    if need > n:
        y = np.concatenate([y, np.zeros(need - n)])
    idx = np.arange(p.frame_len)[None, :] + p.hop_len * np.arange(nf)[:, None]
    frames = y[idx] * window(p)[None, :]
    spec = np.fft.rfft(frames, n=p.nfft, axis=1)
    power = (spec.real ** 2 + spec.imag ** 2) / p.nfft
    E = power @ mel_weights(p).T
    floor = float(np.float32(p.log_floor))
    L = np.log(np.maximum(E, floor))
    energy = getattr(p, "energy", 0)
    log_e = np.log(np.maximum(power.sum(axis=1), floor))[:, None]      # total of the one-sided power spectrum
    if p.output == 1:
        return np.concatenate([L, log_e], axis=1) if energy == 2 else L
    c = scipy.fft.dct(L, type=2, norm="ortho", axis=1)[:, : p.n_cep]
    if p.lifter > 0:
        k = np.arange(p.n_cep)
        c = c * (1.0 + 0.5 * p.lifter * np.sin(np.pi * k / p.lifter))[None, :]
    if energy == 1:
        c[:, :1] = log_e
    elif energy == 2:
        c = np.concatenate([c, log_e], axis=1)
    return c
