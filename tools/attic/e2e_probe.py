"""On-box probe of the end-to-end (host buffer) path: raw PCIe copy rates and per-call times of mfcc_compute_host."""
import json, sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from mfcc_b200 import api, config_a
from mfcc_b200.synth import fast_fixed_batch

p = config_a()
plan = api.Plan(p)
pcm, off = fast_fixed_batch(1024, 160000, seed=1000)
frames = plan.batch(off).total_frames
h_in = api.PinnedBuffer((pcm.size,), np.int16); h_in.array[:] = pcm
h_out = api.PinnedBuffer((frames, 13), np.float32)
d = torch.empty(pcm.size, dtype=torch.int16, device='cuda')
t_in = torch.from_numpy(h_in.array)
res = {}
for name, fn, nbytes in (("h2d_pinned", lambda: d.copy_(t_in, non_blocking=True), pcm.nbytes),):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(10):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    res[name] = {"GBps_best": nbytes / min(ts) / 1e9, "GBps_median": nbytes / sorted(ts)[5] / 1e9}
dd = torch.empty((frames, 13), dtype=torch.float32, device='cuda'); t_out = torch.from_numpy(h_out.array)
for _ in range(3): t_out.copy_(dd, non_blocking=True)
torch.cuda.synchronize(); ts = []
for _ in range(10):
    t0 = time.perf_counter(); t_out.copy_(dd, non_blocking=True); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
res["d2h_pinned"] = {"GBps_best": dd.numel() * 4 / min(ts) / 1e9}
ts = []
for i in range(12):
    t0 = time.perf_counter(); plan.compute_host(h_in.array, off, h_out.array); ts.append(time.perf_counter() - t0)
res["compute_host_ms"] = [round(t * 1e3, 2) for t in ts]
res["compute_host_frames_per_s_median"] = frames / sorted(ts[2:])[5]
# pageable source for comparison
ts = []
for i in range(4):
    t0 = time.perf_counter(); plan.compute_host(pcm, off); ts.append(time.perf_counter() - t0)
res["compute_host_pageable_ms"] = [round(t * 1e3, 2) for t in ts]
import os
res["cpus"] = os.cpu_count()
try:
    res["numa"] = open('/sys/devices/system/node/online').read().strip()
except Exception: pass
print(json.dumps(res))
