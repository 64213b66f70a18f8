"""``torch.library`` custom op and DLPack entry over the C ABI.

    import mfcc_b200.torch_ops as ops
    h = ops.Handle(params, offsets, device=0)           # plan + batch (tables and tile table on the device)
    feat = torch.ops.mfcc_b200.compute_batch(pcm, h.id)  # [total_frames, out_dim] f32, on the CURRENT stream
    feat = ops.mfcc_from_dlpack(cupy_or_any_dlpack_array, h)   # zero-copy in, torch tensor (DLPack-exportable) out
    full = torch.ops.mfcc_b200.post(feat, h.id, 2, 2, 2)  # per-utterance CMVN + delta + delta-delta, [total_frames, 3 out_dim]

The op launches the library's own sm_100a kernels through ``mfcc_compute_batch`` / ``_f32`` / ``_g711`` on
``torch.cuda.current_stream()``; it has a fake (meta) implementation so it traces under ``torch.compile`` /
``torch.export`` as an opaque call — PyTorch is plumbing here (memory, streams), never the arithmetic.  There is no
CPU implementation: the op is registered for CUDA only.
"""
from __future__ import annotations

import itertools
from typing import Dict, Sequence

import torch

from . import api
from .params import MfccParams

_HANDLES: Dict[int, "Handle"] = {}
_ids = itertools.count(1)


class Handle:
    """A plan and the batch shape it will be called on, addressable from the op by an integer id."""

    def __init__(self, params: MfccParams, offsets: Sequence[int], device: int = -1, kernel: int = 0):
        self.plan = api.Plan(params, device=device, kernel=kernel)
        self.batch = self.plan.batch(offsets)
        self.id = next(_ids)
        _HANDLES[self.id] = self

    @property
    def shape(self):
        return (self.batch.total_frames, self.plan.out_dim)

    def close(self):
        _HANDLES.pop(self.id, None)
        self.batch.close()
        self.plan.close()


@torch.library.custom_op("mfcc_b200::compute_batch", mutates_args=(), device_types="cuda")
def compute_batch(pcm: torch.Tensor, handle: int, alaw: bool = False) -> torch.Tensor:
    h = _HANDLES[handle]
    return h.plan.compute_batch(h.batch, pcm.contiguous(), alaw=alaw)     # current stream


@compute_batch.register_fake
def _(pcm, handle, alaw=False):
    h = _HANDLES[handle]
    return pcm.new_empty(h.shape, dtype=torch.float32)


@torch.library.custom_op("mfcc_b200::post", mutates_args=(), device_types="cuda")
def post(feat: torch.Tensor, handle: int, cmvn: int = 2, window: int = 2, order: int = 2) -> torch.Tensor:
    """Fused per-utterance CMVN + delta + delta-delta (``mfcc_post_batch``) of the handle's batch: the stacked
    ``[frames, out_dim * (1 + order)]`` matrix, on the current stream."""
    h = _HANDLES[handle]
    return h.plan.post(h.batch, feat.contiguous(), cmvn, window, order)


@post.register_fake
def _(feat, handle, cmvn=2, window=2, order=2):
    h = _HANDLES[handle]
    return feat.new_empty((h.shape[0], h.shape[1] * (1 + order)), dtype=torch.float32)


def mfcc_from_dlpack(x, handle: Handle, alaw: bool = False) -> torch.Tensor:
    """Any object that speaks DLPack (``__dlpack__``: CuPy, Numba, JAX, another torch tensor ...) holding int16 / f32 /
    uint8 samples on the plan's GPU: imported without a copy, results returned as a torch tensor (itself exportable
    with ``torch.utils.dlpack.to_dlpack`` / ``__dlpack__``)."""
    t = x if isinstance(x, torch.Tensor) else torch.from_dlpack(x)
    return torch.ops.mfcc_b200.compute_batch(t.reshape(-1), handle.id, alaw)
