"""N > 1 path on CPU: world_size-2 `gloo` run of the sharding + optional feature gather
(SURVEY.md §8e).  Each rank takes its contiguous utterance range (balanced by frames), produces the
rows of its range — here with the CPU oracle standing in for the kernel, the tests being the one place
allowed to call it — and the gathered matrix must equal the single-process result row for row.
The same code runs over NCCL on GPUs (bench.py, mfcc_b200/sharding.py)."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

import oracle  # noqa: E402
from mfcc_b200 import config_a, config_b  # noqa: E402
from mfcc_b200.sharding import frame_counts, gather_features, local_slice, partition  # noqa: E402
from mfcc_b200.synth import ragged_batch  # noqa: E402


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, cfg, dst, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        p = {"A": config_a, "B": config_b}[cfg]()
        pcm, off = ragged_batch(37, p.frame_len - 5, 40 * p.hop_len, seed=11)   # same seed on every rank
        u0, u1 = partition(p, off, world)[rank]
        s0, s1, loc = local_slice(off, u0, u1)
        feat, fo = oracle.mfcc_batch(p, pcm[s0:s1], loc)
        rows = int(fo[-1])
        assert rows == int(frame_counts(p, off)[u0:u1].sum())
        full, counts = gather_features(torch.from_numpy(np.ascontiguousarray(feat)), rows, p.n_cep, dst=dst)
        assert sum(counts) == int(frame_counts(p, off).sum())
        if dst is None or rank == dst:
            ref, _ = oracle.mfcc_batch(p, pcm, off)
            assert full.shape == ref.shape
            assert np.array_equal(full.numpy(), ref)        # utterance order = rank order, bit for bit
            with open(f"{out_path}.{rank}", "w") as f:
                f.write("ok")
        else:
            assert full is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("cfg,dst", [("A", None), ("B", 0)])
def test_world_size_2_gloo_shard_and_gather(tmp_path, cfg, dst):
    out = str(tmp_path / "done")
    mp.spawn(_worker, args=(2, _free_port(), cfg, dst, out), nprocs=2, join=True)
    ranks = [0, 1] if dst is None else [dst]
    assert all(os.path.exists(f"{out}.{r}") for r in ranks)


def test_partition_with_more_ranks_than_utterances():
    p = config_a()
    off = np.array([0, 16000, 16000, 48000], np.int64)      # three utterances, one of them empty
    parts = partition(p, off, 8)
    assert parts[0][0] == 0 and parts[-1][1] == 3
    assert all(a <= b for a, b in parts) and all(parts[i][1] == parts[i + 1][0] for i in range(7))


def _gpu_worker(rank, world, port, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from mfcc_b200 import api
        p = config_b()
        pcm, off = ragged_batch(2000, 100, 24000, seed=13)          # same on every rank
        plan = api.Plan(p, device=rank)
        u0, u1 = partition(p, off, world)[rank]
        s0, s1, loc = local_slice(off, u0, u1)
        b = plan.batch(loc)
        feat = plan.compute_batch(b, torch.from_numpy(pcm[s0:s1].copy()).cuda())
        host, _ = plan.compute_host(pcm[s0:s1], loc)                 # the host-buffer entry on the slice
        torch.cuda.synchronize()
        assert np.array_equal(host, feat.cpu().numpy())
        full, counts = gather_features(feat, b.total_frames, plan.out_dim)     # NCCL all_gather
        assert sum(counts) == int(frame_counts(p, off).sum())
        whole = plan.compute_batch(plan.batch(off), torch.from_numpy(pcm).cuda())
        torch.cuda.synchronize()
        assert full.shape == whole.shape and torch.equal(full, whole)   # sharded + gathered == one GPU, bit for bit
        # ONE long 48 kHz recording spread over the ranks: pieces cut at frame boundaries with one history sample
        # (sharding.split_stream + mfcc_batch_create_lead), gathered == the whole recording on one GPU, bit for bit
        from mfcc_b200 import config_c
        from mfcc_b200.sharding import split_stream
        from mfcc_b200.synth import noise_utterance
        pc = config_c()
        n = 48000 * 30 + 77
        x = noise_utterance(n, seed=5)
        planc = api.Plan(pc, device=rank)
        f0, f1, sb, se, lead = split_stream(pc, n, world)[rank]
        bc = planc.batch(np.array([0, se - sb], np.int64), lead=[lead])
        piece = planc.compute_batch(bc, torch.from_numpy(x[sb:se].copy()).cuda())
        fullc, countsc = gather_features(piece, bc.total_frames, planc.out_dim)
        wholec = planc.compute_batch(planc.batch(np.array([0, n], np.int64)), torch.from_numpy(x).cuda())
        torch.cuda.synchronize()
        assert bc.total_frames == f1 - f0 and torch.equal(fullc, wholec)
        with open(f"{out_path}.{rank}", "w") as f:
            f.write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_sharded_batch_gathered_over_nccl_equals_single_gpu(tmp_path):
    """SURVEY.md 8e on hardware: a ragged configs[2]-shaped batch partitioned by frames over every visible GPU, the fused
    kernel per rank on its slice, NCCL all_gather of the feature rows; the gathered matrix equals the one-GPU result
    bit for bit on every rank."""
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    world = min(n, 8)
    out = str(tmp_path / "done")
    mp.spawn(_gpu_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert all(os.path.exists(f"{out}.{r}") for r in range(world))
