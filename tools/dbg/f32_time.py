import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.getcwd())
import bench
from mfcc_b200 import api, CONFIGS
for w in ("A", "B3"):
    cfg, maker, desc, _ = bench.WORKLOADS[w]
    p = CONFIGS[cfg]()
    plan = api.Plan(p)
    pcm, off = maker(1000)
    b = plan.batch(off)
    for dt in (torch.int16, torch.float32):
        d = torch.from_numpy(pcm).cuda().to(dt)
        out = torch.empty((b.total_frames, plan.out_dim), dtype=torch.float32, device="cuda")
        s = torch.cuda.current_stream()
        for _ in range(5): plan.compute_batch(b, d, out, s)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(30): plan.compute_batch(b, d, out, s)
        e1.record(s); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 30
        print(json.dumps({"workload": w, "dtype": str(dt), "ms": ms, "frames_per_s": b.total_frames / ms * 1e3}))
        del d, out
