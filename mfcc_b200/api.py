"""Host-side mirror of the C ABI (include/mfcc_b200.h) over ctypes.

PyTorch is plumbing here: it owns device memory and streams; every number is
produced by libmfcc_b200.so's own sm_100a kernels.  There is no CPU fallback —
:func:`load` raises if the shared library is missing, and the library itself
returns MFCC_ECUDA when no sm_100 device is present.

The reference exposes no operator/plugin API for this path (its entry points
are tokenize / parse_tokens / generate_binary, src/mfcc/lex.h:75,
src/mfcc/parser.h:17, src/mfcc/codegen.h:18); what is mirrored is its calling
convention: int status codes, caller-owned output buffers.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

from .params import (MfccParams, KERNEL_AUTO, MFCC_OK, MFCC_EINVAL)  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
# MFCC_B200_LIB selects another BUILD of the same library (the poison build libmfcc_b200_poison.so of the tests);
# it is never a fallback: whatever it names must exist and export the whole ABI.
LIB_PATH = os.environ.get("MFCC_B200_LIB") or os.path.join(_HERE, "libmfcc_b200.so")
_lib: Optional[C.CDLL] = None

# name -> (restype, argtypes); every symbol include/mfcc_b200.h declares.
_P = C.POINTER(MfccParams)
_vp, _i64, _i32 = C.c_void_p, C.c_int64, C.c_int32
ABI = {
    "mfcc_params_init": (C.c_int, [_P, _i32]),
    "mfcc_params_validate": (C.c_int, [_P]),
    "mfcc_num_frames": (_i64, [_P, _i64]),
    "mfcc_out_dim": (_i32, [_P]),
    "mfcc_plan_create": (C.c_int, [_P, _i32, _i32, C.POINTER(_vp)]),
    "mfcc_plan_destroy": (None, [_vp]),
    "mfcc_plan_params": (C.c_int, [_vp, _P]),
    "mfcc_plan_kernel_name": (C.c_char_p, [_vp]),
    "mfcc_plan_window": (_i64, [_vp, _vp]),
    "mfcc_plan_mel_bins": (_i64, [_vp, _vp]),
    "mfcc_plan_mel_weights": (_i64, [_vp, _vp]),
    "mfcc_plan_dct": (_i64, [_vp, _vp]),
    "mfcc_batch_create": (C.c_int, [_vp, _vp, _i64, C.POINTER(_vp)]),
    "mfcc_batch_create_lead": (C.c_int, [_vp, _vp, _vp, _i64, C.POINTER(_vp)]),
    "mfcc_piece_span": (C.c_int, [_vp, _i64, _i64, _i64, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i32)]),
    "mfcc_batch_destroy": (None, [_vp]),
    "mfcc_batch_total_frames": (_i64, [_vp]),
    "mfcc_batch_total_samples": (_i64, [_vp]),
    "mfcc_batch_frame_offsets": (C.c_int, [_vp, _vp]),
    "mfcc_compute_batch": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "mfcc_compute_batch_f32": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "mfcc_compute_batch_g711": (C.c_int, [_vp, _vp, _vp, _i32, _vp, _vp]),
    "mfcc_compute_host": (C.c_int, [_vp, _vp, _vp, _i64, _vp, _vp]),
    "mfcc_compute_host_g711": (C.c_int, [_vp, _vp, _i32, _vp, _i64, _vp, _vp]),
    "mfcc_compute_host_post": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp]),
    "mfcc_compute": (C.c_int, [_vp, _vp, _i64, _vp, C.POINTER(_i64)]),
    "mfcc_cmvn_batch": (C.c_int, [_vp, _vp, _vp, _i32, _vp]),
    "mfcc_delta_batch": (C.c_int, [_vp, _vp, _vp, _i32, _vp, _vp]),
    "mfcc_post_batch": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp]),
    "mfcc_decode_g711": (C.c_int, [_vp, _i64, _i32, _vp, _vp]),
    "mfcc_wav_parse": (C.c_int, [_vp, _i64, _vp]),
    "mfcc_stream_create": (C.c_int, [_vp, C.POINTER(_vp)]),
    "mfcc_stream_destroy": (None, [_vp]),
    "mfcc_stream_pending": (_i64, [_vp, _i64, _i32]),
    "mfcc_stream_feed": (C.c_int, [_vp, _vp, _i64, _vp, _i64, C.POINTER(_i64)]),
    "mfcc_stream_flush": (C.c_int, [_vp, _vp, _i64, C.POINTER(_i64)]),
    "mfcc_stream_feed_many": (C.c_int, [_vp, _i64, _vp, _vp, _vp, _vp, _vp]),
    "mfcc_host_alloc": (C.c_int, [C.POINTER(_vp), _i64]),
    "mfcc_host_free": (C.c_int, [_vp]),
    "mfcc_launch_count": (C.c_uint64, []),
    "mfcc_strerror": (C.c_char_p, [C.c_int]),
    "mfcc_version": (C.c_char_p, []),
}


class WavInfo(C.Structure):
    """ctypes mirror of ``mfcc_wav_info``."""
    _fields_ = [("format", _i32), ("channels", _i32), ("sample_rate", _i32), ("bits_per_sample", _i32),
                ("data_offset", _i64), ("data_bytes", _i64), ("n_frames", _i64)]


WAV_PCM16, WAV_MULAW, WAV_ALAW, WAV_F32 = 1, 2, 3, 4


class MfccError(RuntimeError):
    def __init__(self, code: int, where: str):
        self.code = code
        msg = load().mfcc_strerror(code).decode() if _lib is not None else str(code)
        super().__init__(f"{where}: {msg} ({code})")


def load() -> C.CDLL:
    """Load libmfcc_b200.so (built by ``__graft_entry__.build()`` / ``make -C mfcc_b200/csrc``)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                "There is no CPU fallback for the MFCC path.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in ABI.items():
            fn = getattr(lib, name)  # AttributeError here = header/library drift
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def _check(rc: int, where: str) -> None:
    if rc != MFCC_OK:
        raise MfccError(rc, where)


def num_frames(p: MfccParams, n_samples: int) -> int:
    return int(load().mfcc_num_frames(C.byref(p), n_samples))


def launch_count() -> int:
    return int(load().mfcc_launch_count())


def _stream_handle(stream) -> Optional[int]:
    if stream is None:
        import torch
        return torch.cuda.current_stream().cuda_stream
    return getattr(stream, "cuda_stream", stream)


def piece_span(params: MfccParams, n_samples: int, f0: int, f1: int):
    """``mfcc_piece_span``: (begin, end, lead) of the piece holding frames [f0, f1) of a recording of n_samples samples."""
    b, e, l = _i64(0), _i64(0), _i32(0)
    _check(load().mfcc_piece_span(C.byref(params), n_samples, f0, f1, C.byref(b), C.byref(e), C.byref(l)), "mfcc_piece_span")
    return int(b.value), int(e.value), int(l.value)


class Batch:
    """The shape of one batch: utterance offsets -> frame rows -> tile table (device-resident)."""

    def __init__(self, plan: "Plan", offsets: Sequence[int], lead: Optional[Sequence[int]] = None):
        self.plan = plan
        self.offsets = np.ascontiguousarray(offsets, np.int64)
        if self.offsets.ndim != 1 or self.offsets.size < 1:
            raise ValueError("offsets must be a 1-D array of n_utts + 1 entries")
        self.n_utts = self.offsets.size - 1
        h = _vp()
        if lead is None:
            _check(load().mfcc_batch_create(plan._h, self.offsets.ctypes.data, self.n_utts, C.byref(h)),
                   "mfcc_batch_create")
        else:   # pieces of longer recordings: lead[u] = 1 marks a first sample that is history only
            self.lead = np.ascontiguousarray(lead, np.uint8)
            if self.lead.shape != (self.n_utts,):
                raise ValueError("lead must hold one flag per utterance")
            _check(load().mfcc_batch_create_lead(plan._h, self.offsets.ctypes.data, self.lead.ctypes.data, self.n_utts,
                                                 C.byref(h)), "mfcc_batch_create_lead")
        self._h = h
        self.total_frames = int(load().mfcc_batch_total_frames(h))
        self.total_samples = int(load().mfcc_batch_total_samples(h))
        self.frame_offsets = np.empty(self.n_utts + 1, np.int64)
        _check(load().mfcc_batch_frame_offsets(h, self.frame_offsets.ctypes.data), "mfcc_batch_frame_offsets")

    def close(self):
        if getattr(self, "_h", None):
            load().mfcc_batch_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:   # interpreter shutdown: module globals (load, _lib) may already be gone
            pass


class Plan:
    """Validated parameters + device tables on one GPU (``mfcc_plan`` in the C ABI)."""

    def __init__(self, params: MfccParams, device: int = -1, kernel: int = KERNEL_AUTO):
        self.params = params.copy()
        h = _vp()
        _check(load().mfcc_plan_create(C.byref(self.params), device, kernel, C.byref(h)), "mfcc_plan_create")
        self._h = h
        self.out_dim = self.params.out_dim
        self.kernel_name = load().mfcc_plan_kernel_name(h).decode()
        if device < 0:
            import torch
            device = torch.cuda.current_device()
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            load().mfcc_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:   # interpreter shutdown: module globals (load, _lib) may already be gone
            pass

    # ---- tables (host copies of what the kernels read) ----
    def window(self) -> np.ndarray:
        a = np.empty(load().mfcc_plan_window(self._h, None), np.float32)
        load().mfcc_plan_window(self._h, a.ctypes.data)
        return a

    def mel_bins(self) -> np.ndarray:
        a = np.empty(load().mfcc_plan_mel_bins(self._h, None), np.int32)
        load().mfcc_plan_mel_bins(self._h, a.ctypes.data)
        return a

    def mel_weights(self) -> np.ndarray:
        a = np.empty(load().mfcc_plan_mel_weights(self._h, None), np.float32)
        load().mfcc_plan_mel_weights(self._h, a.ctypes.data)
        return a.reshape(self.params.n_mel, self.params.nfft // 2 + 1)

    def dct(self) -> np.ndarray:
        a = np.empty(load().mfcc_plan_dct(self._h, None), np.float32)
        load().mfcc_plan_dct(self._h, a.ctypes.data)
        return a.reshape(self.params.n_cep, self.params.n_mel)

    # ---- the hot entry: device tensors in, device tensor out ----
    def batch(self, offsets: Sequence[int], lead: Optional[Sequence[int]] = None) -> Batch:
        return Batch(self, offsets, lead)

    def compute_batch(self, batch: Batch, pcm, out=None, stream=None, alaw: bool = False):
        """``pcm``: torch CUDA tensor, int16 (or float32 scaled to int16 range, or uint8 G.711 codes — mu-law unless
        ``alaw``), >= total_samples elements.  Returns ``out`` [total_frames, out_dim] float32 on the same device.
        Asynchronous."""
        import torch
        if not pcm.is_cuda or pcm.device.index != self.device or not pcm.is_contiguous():
            raise ValueError("pcm must be a contiguous CUDA tensor on the plan's device")
        if pcm.numel() < batch.total_samples:
            raise ValueError("pcm is shorter than the batch's offsets")
        if out is None:
            out = torch.empty((batch.total_frames, self.out_dim), dtype=torch.float32, device=pcm.device)
        elif (out.dtype != torch.float32 or not out.is_contiguous() or out.device != pcm.device
              or out.numel() < batch.total_frames * self.out_dim):
            raise ValueError("out must be a contiguous float32 CUDA tensor of total_frames * out_dim")
        with torch.cuda.device(self.device):
            if pcm.dtype == torch.int16:
                _check(load().mfcc_compute_batch(self._h, batch._h, pcm.data_ptr(), out.data_ptr(),
                                                 _stream_handle(stream)), "mfcc_compute_batch")
            elif pcm.dtype == torch.float32:
                _check(load().mfcc_compute_batch_f32(self._h, batch._h, pcm.data_ptr(), out.data_ptr(),
                                                     _stream_handle(stream)), "mfcc_compute_batch_f32")
            elif pcm.dtype == torch.uint8:
                _check(load().mfcc_compute_batch_g711(self._h, batch._h, pcm.data_ptr(), int(alaw), out.data_ptr(),
                                                      _stream_handle(stream)), "mfcc_compute_batch_g711")
            else:
                raise TypeError("pcm must be int16, float32 or uint8 (G.711 codes)")
        return out

    # ---- end to end with host buffers (H2D + kernels + D2H inside) ----
    def compute_host(self, pcm: np.ndarray, offsets: Sequence[int], out: Optional[np.ndarray] = None,
                     alaw: Optional[bool] = None, post: Optional[Sequence[int]] = None):
        """Host buffers in, host buffers out.  ``alaw`` given (False: mu-law, True: A-law): ``pcm`` holds G.711 codes
        (uint8) and goes through ``mfcc_compute_host_g711`` — 1 byte per sample over PCIe.  ``post`` = (cmvn, window,
        order): int16 PCM through ``mfcc_compute_host_post`` — the rows that come back are the stacked static | delta |
        delta-delta matrix (``out_dim * (1 + order)`` columns) after per-utterance CMVN."""
        g711 = alaw is not None
        if post is not None:
            if g711:
                raise ValueError("compute_host: post-processing is offered for int16 PCM")
            cmvn, window, order = (int(v) for v in post)
            if cmvn not in (0, 1, 2) or order not in (0, 1, 2) or (order > 0 and not 1 <= window <= 8):
                raise ValueError("compute_host: post = (cmvn in 0..2, window in 1..8, order in 0..2)")
        width = self.out_dim * (1 + (post[2] if post is not None else 0))
        pcm = np.ascontiguousarray(pcm, np.uint8 if g711 else np.int16)
        offsets = np.ascontiguousarray(offsets, np.int64)
        if offsets.ndim != 1 or offsets.size < 1:
            raise ValueError("offsets must be a 1-D array of n_utts + 1 entries")
        n_utts = offsets.size - 1
        if n_utts > 0 and (offsets[0] < 0 or (np.diff(offsets) < 0).any()):
            raise MfccError(-1, "mfcc_compute_host (offsets must be non-negative and non-decreasing)")
        if int(offsets[-1]) > pcm.size:
            raise ValueError(f"offsets run to sample {int(offsets[-1])} but pcm holds {pcm.size}")
        fo = np.empty(n_utts + 1, np.int64)
        from .sharding import frame_counts
        total = int(frame_counts(self.params, offsets).sum())
        if out is None:
            out = np.empty((total, width), np.float32)
        elif (not isinstance(out, np.ndarray) or out.dtype != np.float32 or not out.flags.c_contiguous
              or out.size < total * width):
            raise ValueError(f"out must be a C-contiguous float32 array of at least {total} x {width} elements")
        if post is not None:
            _check(load().mfcc_compute_host_post(self._h, pcm.ctypes.data, offsets.ctypes.data, n_utts, cmvn, window, order,
                                                 out.ctypes.data, fo.ctypes.data), "mfcc_compute_host_post")
        elif g711:
            _check(load().mfcc_compute_host_g711(self._h, pcm.ctypes.data, int(bool(alaw)), offsets.ctypes.data, n_utts,
                                                 out.ctypes.data, fo.ctypes.data), "mfcc_compute_host_g711")
        else:
            _check(load().mfcc_compute_host(self._h, pcm.ctypes.data, offsets.ctypes.data, n_utts,
                                            out.ctypes.data, fo.ctypes.data), "mfcc_compute_host")
        return out, fo

    def compute(self, pcm: np.ndarray) -> np.ndarray:
        """Single clip, host in / host out (``mfcc_compute``)."""
        pcm = np.ascontiguousarray(pcm, np.int16)
        nf = num_frames(self.params, pcm.size)
        out = np.empty((max(nf, 0), self.out_dim), np.float32)
        got = _i64(0)
        _check(load().mfcc_compute(self._h, pcm.ctypes.data, pcm.size, out.ctypes.data, C.byref(got)),
               "mfcc_compute")
        assert got.value == nf
        return out

    # ---- §8(f) widening ----
    def _check_feat(self, batch: Batch, feat, what: str):
        import torch
        if (not isinstance(feat, torch.Tensor) or not feat.is_cuda or feat.device.index != self.device
                or feat.dtype != torch.float32 or not feat.is_contiguous()
                or feat.numel() < batch.total_frames * self.out_dim):
            raise ValueError(f"{what}: feat must be a contiguous float32 CUDA tensor on device {self.device} with at "
                             f"least {batch.total_frames} x {self.out_dim} elements")

    def cmvn(self, batch: Batch, feat, norm_var: bool = False, stream=None):
        import torch
        self._check_feat(batch, feat, "cmvn")
        with torch.cuda.device(self.device):
            _check(load().mfcc_cmvn_batch(self._h, batch._h, feat.data_ptr(), int(norm_var),
                                          _stream_handle(stream)), "mfcc_cmvn_batch")
        return feat

    def delta(self, batch: Batch, feat, window: int = 2, stream=None):
        import torch
        self._check_feat(batch, feat, "delta")
        d = torch.empty_like(feat)
        with torch.cuda.device(self.device):
            _check(load().mfcc_delta_batch(self._h, batch._h, feat.data_ptr(), window, d.data_ptr(),
                                           _stream_handle(stream)), "mfcc_delta_batch")
        return d

    def post(self, batch: Batch, feat, cmvn: int = 1, window: int = 2, order: int = 2, out=None, stream=None):
        """Fused CMVN + delta + delta-delta (``mfcc_post_batch``): the stacked ``[frames][out_dim * (1 + order)]`` matrix
        static | delta | delta-delta.  ``cmvn``: 0 none, 1 mean, 2 mean and variance."""
        import torch
        self._check_feat(batch, feat, "post")
        if cmvn not in (0, 1, 2) or order not in (0, 1, 2) or (order > 0 and not 1 <= window <= 8):
            raise ValueError("post: cmvn in 0..2, order in 0..2, window in 1..8")
        od = self.out_dim * (1 + order)
        if out is None:
            out = torch.empty((batch.total_frames, od), dtype=torch.float32, device=feat.device)
        elif (not isinstance(out, torch.Tensor) or not out.is_cuda or out.device != feat.device
              or out.dtype != torch.float32 or not out.is_contiguous() or out.numel() < batch.total_frames * od):
            raise ValueError(f"post: out must be a contiguous float32 CUDA tensor with at least {batch.total_frames} x {od} elements")
        with torch.cuda.device(self.device):
            _check(load().mfcc_post_batch(self._h, batch._h, feat.data_ptr(), cmvn, window, order, out.data_ptr(),
                                          _stream_handle(stream)), "mfcc_post_batch")
        return out


class Stream:
    """Online front end for one audio stream (``mfcc_stream_*``): feed chunks, get the frames that are complete."""

    def __init__(self, plan: Plan):
        self.plan = plan
        h = _vp()
        _check(load().mfcc_stream_create(plan._h, C.byref(h)), "mfcc_stream_create")
        self._h = h

    def close(self):
        if getattr(self, "_h", None) and load is not None:
            load().mfcc_stream_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:   # interpreter shutdown: module globals (load, _lib) may already be gone
            pass

    def feed(self, pcm: np.ndarray) -> np.ndarray:
        pcm = np.ascontiguousarray(pcm, np.int16)
        due = int(load().mfcc_stream_pending(self._h, pcm.size, 0))
        if due < 0:
            raise MfccError(due, "mfcc_stream_pending")
        out = np.empty((due, self.plan.out_dim), np.float32)
        got = _i64(0)
        _check(load().mfcc_stream_feed(self._h, pcm.ctypes.data, pcm.size, out.ctypes.data, due, C.byref(got)),
               "mfcc_stream_feed")
        assert got.value == due
        return out

    def flush(self) -> np.ndarray:
        due = max(int(load().mfcc_stream_pending(self._h, 0, 1)), 0)
        out = np.empty((due, self.plan.out_dim), np.float32)
        got = _i64(0)
        _check(load().mfcc_stream_flush(self._h, out.ctypes.data, due, C.byref(got)), "mfcc_stream_flush")
        return out[: got.value]


class StreamGroup:
    """S live streams of one plan fed together (``mfcc_stream_feed_many``): one staging copy, one launch, one read-back
    per call whatever S is.  ``feed(chunks)`` takes either a [S, n] int16 array (every stream gets n samples) or a
    list of S 1-D arrays; it returns (rows [S, cap, out_dim] float32, counts [S]) — stream i's new frames are
    ``rows[i, :counts[i]]``."""

    def __init__(self, plan: Plan, n_streams: int, max_frames_per_feed: int = 64):
        self.plan, self.n = plan, int(n_streams)
        self.streams = [Stream(plan) for _ in range(self.n)]
        self._handles = (C.c_void_p * self.n)(*[st._h.value for st in self.streams])
        self.cap = int(max_frames_per_feed)
        self.rows = np.zeros((self.n, self.cap, plan.out_dim), np.float32)
        stride = self.rows.strides[0]
        self._out = (self.rows.ctypes.data + stride * np.arange(self.n, dtype=np.uint64)).astype(np.uint64)
        self._cap = np.full(self.n, self.cap, np.int64)
        self.counts = np.zeros(self.n, np.int64)

    def feed(self, chunks):
        if isinstance(chunks, np.ndarray) and chunks.ndim == 2:
            if chunks.dtype != np.int16 or chunks.shape[0] != self.n or chunks.strides[1] != 2:
                raise ValueError("chunks must be [n_streams, n] int16 with contiguous rows")
            ptrs = (chunks.ctypes.data + chunks.strides[0] * np.arange(self.n, dtype=np.uint64)).astype(np.uint64)
            lens = np.full(self.n, chunks.shape[1], np.int64)
            keep = chunks
        else:
            keep = [np.ascontiguousarray(c, np.int16) for c in chunks]
            if len(keep) != self.n:
                raise ValueError("one chunk per stream")
            ptrs = np.array([c.ctypes.data for c in keep], np.uint64)
            lens = np.array([c.size for c in keep], np.int64)
        _check(load().mfcc_stream_feed_many(C.addressof(self._handles), self.n, ptrs.ctypes.data, lens.ctypes.data,
                                            self._out.ctypes.data, self._cap.ctypes.data, self.counts.ctypes.data),
               "mfcc_stream_feed_many")
        del keep
        return self.rows, self.counts

    def close(self):
        for st in self.streams:
            st.close()


def decode_g711(codes, alaw: bool = False, stream=None):
    """uint8 CUDA tensor of G.711 codes -> int16 CUDA tensor of PCM."""
    import torch
    if codes.dtype != torch.uint8 or not codes.is_cuda or not codes.is_contiguous():
        raise ValueError("codes must be a contiguous uint8 CUDA tensor")
    out = torch.empty(codes.numel(), dtype=torch.int16, device=codes.device)
    with torch.cuda.device(codes.device):
        _check(load().mfcc_decode_g711(codes.data_ptr(), codes.numel(), int(alaw), out.data_ptr(),
                                       _stream_handle(stream)), "mfcc_decode_g711")
    return out


class PinnedBuffer:
    """Pinned host memory from the library (``mfcc_host_alloc``), viewed as a numpy array."""

    def __init__(self, shape, dtype):
        self.dtype = np.dtype(dtype)
        self.shape = tuple(np.atleast_1d(shape).tolist()) if not isinstance(shape, tuple) else shape
        n = int(np.prod(self.shape)) * self.dtype.itemsize
        p = _vp()
        _check(load().mfcc_host_alloc(C.byref(p), max(n, 1)), "mfcc_host_alloc")
        self._p = p
        buf = (C.c_char * max(n, 1)).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(self.shape))).reshape(self.shape)

    def close(self):
        if getattr(self, "_p", None):
            self.array = None
            load().mfcc_host_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.close()
        except Exception:   # interpreter shutdown: module globals (load, _lib) may already be gone
            pass


def wav_parse(data) -> WavInfo:
    """``mfcc_wav_parse`` over a bytes-like object holding a RIFF/WAVE file."""
    buf = np.frombuffer(data, np.uint8)
    info = WavInfo()
    _check(load().mfcc_wav_parse(buf.ctypes.data if buf.size else None, buf.size, C.byref(info)), "mfcc_wav_parse")
    return info


def wav_samples(data):
    """(info, samples): the data chunk viewed in place as [n_frames, channels] int16 (PCM16), uint8 (G.711 codes) or
    float32 (IEEE float, full scale 1.0) — no conversion; pick the device entry by ``info.format``."""
    info = wav_parse(data)
    dt = {WAV_PCM16: "<i2", WAV_MULAW: np.uint8, WAV_ALAW: np.uint8, WAV_F32: "<f4"}[info.format]
    a = np.frombuffer(data, dt, count=info.n_frames * info.channels, offset=info.data_offset)
    return info, a.reshape(info.n_frames, info.channels)
