#!/usr/bin/env python
"""The box's aggregate pinned-host -> device copy rate with N ranks copying at once (one process per GPU): the ceiling of
every e2e (host-buffer) number at N GPUs.  Plain torch copies, nothing of this library involved.

    python tools/h2d_ceiling.py                                  # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/h2d_ceiling.py
"""
import json
import os
import time

import torch
import torch.distributed as dist

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 327_680_000 // 2 * 2            # configs[1]: 327.7 MB of int16 per batch
h = torch.empty(n // 2, dtype=torch.int16).pin_memory()
h.random_(-3000, 3000)
d = torch.empty_like(h, device="cuda")
ho = torch.empty(53_141_504 // 4, dtype=torch.float32).pin_memory()     # 53.1 MB of features back
do = torch.empty_like(ho, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def sync():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()


def run(both, reps=10):
    sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        with torch.cuda.stream(s1):
            d.copy_(h, non_blocking=True)
        if both:
            with torch.cuda.stream(s2):
                ho.copy_(do, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    return dt / reps


for _ in range(2):
    run(True, 2)
t_h2d, t_both = run(False), run(True)
if rank == 0:
    print(json.dumps({"bench": "h2d_ceiling", "n_gpus": world, "bytes_h2d_per_rank": h.numel() * 2, "bytes_d2h_per_rank": ho.numel() * 4,
                      "h2d_only_gbs_per_rank": h.numel() * 2 / t_h2d / 1e9, "h2d_only_gbs_aggregate": world * h.numel() * 2 / t_h2d / 1e9,
                      "both_ms_per_batch": t_both * 1e3,
                      "both_h2d_gbs_aggregate": world * h.numel() * 2 / t_both / 1e9,
                      "implied_e2e_ceiling_frames_per_s_configs1": world * 1021952 / t_both,
                      "host_cpus": os.cpu_count()}))
if world > 1:
    dist.destroy_process_group()
