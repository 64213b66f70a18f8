"""ctypes loader for the CPU oracle (oracle/libmfcc_oracle.so).

TEST INFRASTRUCTURE.  Importable only from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  PARITY UNPINNED: see the
header of oracle/mfcc_oracle.c.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from mfcc_b200.params import MfccParams

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libmfcc_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("mfcc_oracle.c", "mfcc_oracle_impl.h", "mfcc_oracle.h", "mfcc_cpu_fast.c")]
    srcs.append(os.path.join(_HERE, "..", "include", "mfcc_b200.h"))
    stale = force or not os.path.exists(_SO) or any(
        os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if stale:
        subprocess.run(["make", "-C", _HERE, "-B", "libmfcc_oracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return _SO


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        P = C.POINTER(MfccParams)
        i64, vp = C.c_int64, C.c_void_p
        L.oracle_params_validate.argtypes = [P]
        L.oracle_num_frames.argtypes = [P, i64]
        L.oracle_num_frames.restype = i64
        for f in ("oracle_window_f64", "oracle_mel_bins", "oracle_mel_weights_f64", "oracle_dct_f64"):
            getattr(L, f).argtypes = [P, vp]
        for f in ("oracle_mfcc_f32", "oracle_mfcc_f64"):
            getattr(L, f).argtypes = [P, vp, i64, vp]
            getattr(L, f).restype = i64
        for f in ("oracle_stages_f32", "oracle_stages_f64"):
            getattr(L, f).argtypes = [P, vp, i64, i64, vp, vp, vp, vp]
        for f in ("oracle_mfcc_batch_f32", "oracle_fast_mfcc_batch_f32"):
            getattr(L, f).argtypes = [P, vp, vp, i64, vp, vp, C.c_int]
            getattr(L, f).restype = i64
        L.oracle_cmvn_f32.argtypes = [vp, vp, i64, C.c_int, C.c_int]
        L.oracle_delta_f32.argtypes = [vp, vp, i64, C.c_int, C.c_int, vp]
        L.oracle_decode_g711.argtypes = [vp, i64, C.c_int, vp]
        _lib = L
    return _lib


def _ptr(a: np.ndarray) -> int:
    return a.ctypes.data


def num_frames(p: MfccParams, n: int) -> int:
    return int(lib().oracle_num_frames(C.byref(p), n))


def window(p: MfccParams) -> np.ndarray:
    w = np.empty(p.frame_len, np.float64)
    assert lib().oracle_window_f64(C.byref(p), _ptr(w)) == 0
    return w


def mel_bins(p: MfccParams) -> np.ndarray:
    b = np.empty(p.n_mel + 2, np.int32)
    assert lib().oracle_mel_bins(C.byref(p), _ptr(b)) == 0
    return b


def mel_weights(p: MfccParams) -> np.ndarray:
    W = np.empty((p.n_mel, p.nfft // 2 + 1), np.float64)
    assert lib().oracle_mel_weights_f64(C.byref(p), _ptr(W)) == 0
    return W


def dct(p: MfccParams) -> np.ndarray:
    D = np.empty((p.n_cep, p.n_mel), np.float64)
    assert lib().oracle_dct_f64(C.byref(p), _ptr(D)) == 0
    return D


def mfcc(p: MfccParams, pcm: np.ndarray, dtype=np.float32) -> np.ndarray:
    pcm = np.ascontiguousarray(pcm, np.int16)
    nf = num_frames(p, pcm.size)
    if nf < 0:
        raise ValueError(f"oracle: bad parameters ({nf})")
    out = np.empty((nf, p.out_dim), dtype)
    fn = lib().oracle_mfcc_f32 if dtype == np.float32 else lib().oracle_mfcc_f64
    rc = fn(C.byref(p), _ptr(pcm), pcm.size, _ptr(out))
    if rc != nf:
        raise RuntimeError(f"oracle_mfcc returned {rc}, expected {nf}")
    return out


def stages(p: MfccParams, pcm: np.ndarray, frame: int, dtype=np.float64):
    pcm = np.ascontiguousarray(pcm, np.int16)
    fr = np.empty(p.nfft, dtype)
    pw = np.empty(p.nfft // 2 + 1, dtype)
    mel = np.empty(p.n_mel, dtype)
    out = np.empty(p.out_dim, dtype)
    fn = lib().oracle_stages_f32 if dtype == np.float32 else lib().oracle_stages_f64
    rc = fn(C.byref(p), _ptr(pcm), pcm.size, frame, _ptr(fr), _ptr(pw), _ptr(mel), _ptr(out))
    if rc != 0:
        raise ValueError(f"oracle_stages: {rc}")
    return fr, pw, mel, out


def mfcc_batch(p: MfccParams, pcm: np.ndarray, offsets: np.ndarray, nthreads: int = 1, fast: bool = False):
    """Returns (features [total_frames, out_dim] f32, frame_offsets [B+1] i64).  ``fast``: the CPU BASELINE build of the
    same spec (mfcc_cpu_fast.c: real-input FFT, -O3, AVX2 clones) instead of the plain parity oracle."""
    pcm = np.ascontiguousarray(pcm, np.int16)
    offsets = np.ascontiguousarray(offsets, np.int64)
    B = offsets.size - 1
    fo = np.empty(B + 1, np.int64)
    fn = lib().oracle_fast_mfcc_batch_f32 if fast else lib().oracle_mfcc_batch_f32
    total = fn(C.byref(p), _ptr(pcm), _ptr(offsets), B, None, _ptr(fo), 1)
    if total < 0:
        raise ValueError(f"oracle_mfcc_batch: {total}")
    out = np.empty((total, p.out_dim), np.float32)
    rc = fn(C.byref(p), _ptr(pcm), _ptr(offsets), B, _ptr(out), _ptr(fo), nthreads)
    if rc != total:
        raise RuntimeError(f"oracle_mfcc_batch: {rc}")
    return out, fo


def cmvn(feat: np.ndarray, frame_offsets: np.ndarray, norm_var: bool) -> np.ndarray:
    f = np.array(feat, np.float32, copy=True, order="C")
    fo = np.ascontiguousarray(frame_offsets, np.int64)
    assert lib().oracle_cmvn_f32(_ptr(f), _ptr(fo), fo.size - 1, f.shape[1], int(norm_var)) == 0
    return f


def post(feat: np.ndarray, frame_offsets: np.ndarray, cmvn_mode: int = 1, window: int = 2, order: int = 2) -> np.ndarray:
    """The stacked matrix static | delta | delta-delta the fused post-processing kernel writes, by composition of the
    oracle's own functions: cmvn (mode 0 none, 1 mean, 2 mean and variance), delta, delta of delta."""
    x = np.ascontiguousarray(feat, np.float32) if cmvn_mode == 0 else cmvn(feat, frame_offsets, cmvn_mode == 2)
    parts = [x]
    for _ in range(order):
        parts.append(delta(parts[-1], frame_offsets, window))
    return np.concatenate(parts, axis=1)


def delta(feat: np.ndarray, frame_offsets: np.ndarray, window: int = 2) -> np.ndarray:
    f = np.ascontiguousarray(feat, np.float32)
    fo = np.ascontiguousarray(frame_offsets, np.int64)
    d = np.empty_like(f)
    assert lib().oracle_delta_f32(_ptr(f), _ptr(fo), fo.size - 1, f.shape[1], window, _ptr(d)) == 0
    return d


def decode_g711(src: np.ndarray, alaw: bool) -> np.ndarray:
    s = np.ascontiguousarray(src, np.uint8)
    d = np.empty(s.size, np.int16)
    assert lib().oracle_decode_g711(_ptr(s), s.size, int(alaw), _ptr(d)) == 0
    return d
