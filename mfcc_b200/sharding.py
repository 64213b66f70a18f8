"""Multi-GPU sharding of a batch of utterances.

The path shards by independent units: utterances.  Each rank takes a
CONTIGUOUS range of utterances chosen so that cumulative FRAME counts (not
utterance counts — it matters for ragged batches) are balanced; tables are
replicated per device; there is no data-path collective.  The only exchange is
the optional gather of the feature rows, done with torch.distributed
(NCCL on GPUs, gloo in the CPU tests).

The reference has no distributed support of any kind.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np

from .params import MfccParams


def frame_counts(p: MfccParams, offsets: Sequence[int]) -> np.ndarray:
    """Frames per utterance under the plan's framing rule (pure integer math,
    identical to mfcc_num_frames in the C ABI)."""
    off = np.asarray(offsets, np.int64)
    n = np.diff(off)
    if (n < 0).any():
        raise ValueError("offsets must be non-decreasing")
    L, H = int(p.frame_len), int(p.hop_len)
    if p.pad_mode == 0:
        return np.where(n < L, 0, 1 + (n - L) // H).astype(np.int64)
    return np.where(n == 0, 0, np.where(n <= L, 1, 1 + (n - L + H - 1) // H)).astype(np.int64)


def partition(p: MfccParams, offsets: Sequence[int], world_size: int) -> List[Tuple[int, int]]:
    """Contiguous utterance ranges [u0, u1) per rank, balanced by cumulative frames.

    Rank r ends at the utterance boundary whose cumulative frame count is
    closest to (r + 1) / world_size of the total.  Every utterance is assigned to
    exactly one rank; ranks may be empty when there are fewer utterances than ranks.
    """
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    nf = frame_counts(p, offsets)
    B = nf.size
    cum = np.concatenate([[0], np.cumsum(nf)])
    total = int(cum[-1])
    cuts = [0]
    for r in range(1, world_size):
        target = total * r / world_size
        u = int(np.searchsorted(cum, target, side="left"))
        u = min(u, B)
        if u > 0 and abs(cum[u - 1] - target) <= abs(cum[u] - target):
            u -= 1
        u = min(max(u, cuts[-1]), B)
        cuts.append(u)
    cuts.append(B)
    return [(cuts[r], cuts[r + 1]) for r in range(world_size)]


def split_stream(p: MfccParams, n_samples: int, pieces: int) -> List[Tuple[int, int, int, int, int]]:
    """ONE long recording over several GPUs ("long streams may additionally be split at frame boundaries with
    a (frame_len - hop) + 1-sample read overlap"): `pieces` contiguous frame ranges of near-equal size.  Returns per piece
    ``(f0, f1, begin, end, lead)``: frames [f0, f1) live in samples [begin, end) of the recording, and ``lead`` = 1 says the
    first of those samples is history only (the pre-emphasis predecessor of the piece's first frame) — pass it to
    ``Plan.batch(offsets, lead=...)`` / ``mfcc_batch_create_lead``.  The rows of the pieces, concatenated, are the rows of
    the whole recording bit for bit.  Pure integer math, the same as ``mfcc_piece_span`` in the C ABI."""
    if pieces < 1:
        raise ValueError("pieces must be >= 1")
    total = int(frame_counts(p, [0, n_samples])[0])
    L, H = int(p.frame_len), int(p.hop_len)
    out = []
    for r in range(pieces):
        f0, f1 = total * r // pieces, total * (r + 1) // pieces
        if f1 == f0:
            out.append((f0, f1, 0, 0, 0))
            continue
        lead = 1 if f0 > 0 else 0
        out.append((f0, f1, f0 * H - lead, n_samples if f1 == total else (f1 - 1) * H + L, lead))
    return out


def local_slice(offsets: Sequence[int], u0: int, u1: int) -> Tuple[int, int, np.ndarray]:
    """Sample range [s0, s1) of utterances [u0, u1) and their offsets rebased to 0."""
    off = np.asarray(offsets, np.int64)
    s0, s1 = int(off[u0]), int(off[u1])
    return s0, s1, (off[u0:u1 + 1] - s0).astype(np.int64)


def gather_features(local_feat, local_rows: int, out_dim: int, group=None, dst: Optional[int] = None):
    """Optional gather of the [rows_r, out_dim] feature blocks of all ranks, in rank
    order (= utterance order, because ranges are contiguous).  Rows differ per
    rank, so blocks are padded to the maximum and trimmed after the collective.
    ``dst=None`` -> all_gather (every rank gets the full matrix); else gather to ``dst``.
    Works on CUDA tensors over NCCL and on CPU tensors over gloo."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    rows = torch.tensor([local_rows], dtype=torch.int64, device=local_feat.device)
    all_rows = [torch.zeros_like(rows) for _ in range(world)]
    dist.all_gather(all_rows, rows, group=group)
    counts = [int(r.item()) for r in all_rows]
    mx = max(max(counts), 1)
    padded = torch.zeros((mx, out_dim), dtype=local_feat.dtype, device=local_feat.device)
    padded[:local_rows] = local_feat[:local_rows]
    if dst is None:
        blocks = [torch.empty_like(padded) for _ in range(world)]
        dist.all_gather(blocks, padded, group=group)
    else:
        blocks = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
        dist.gather(padded, blocks, dst=dst, group=group)
        if rank != dst:
            return None, counts
    return torch.cat([b[:c] for b, c in zip(blocks, counts)], dim=0), counts


def bind_near_gpu(device_index: int) -> Optional[List[int]]:
    """Pin the calling process to the CPU cores NVML reports as local to GPU ``device_index`` (its NUMA
    node), so that the pinned host buffers allocated afterwards are first-touched next to the GPU's PCIe
    root port.  With one process per GPU and host-resident audio (mfcc_compute_host) this keeps 8 concurrent
    H2D streams off the inter-socket link.  Returns the core list, or None when NVML / affinity is unavailable
    (nothing is changed then).  Honours CUDA_VISIBLE_DEVICES through the device's PCI bus id."""
    import os
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(device_index)
        bus_id = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        h = pynvml.nvmlDeviceGetHandleByPciBusId(bus_id.encode())
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cores = [64 * w + b for w, mask in enumerate(words) for b in range(64) if (int(mask) >> b) & 1]
        allowed = sorted(set(cores) & set(os.sched_getaffinity(0)))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return allowed
    except Exception:
        return None
