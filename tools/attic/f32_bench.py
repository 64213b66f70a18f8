"""Throughput of the f32 PCM entry (mfcc_compute_batch_f32) next to the int16 entry; `f32 scalar` misaligns the base
pointer by one sample so that the kernel falls back to its per-sample staging (what every f32 call took before the
vector path existed)."""
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from mfcc_b200 import api, config_a, config_b, config_c
from mfcc_b200.synth import fast_fixed_batch, ragged_batch
for name, p, (pcm, off) in (("A", config_a(), fast_fixed_batch(1024, 160000, seed=1)), ("B3", config_b(), ragged_batch(16384, 4000, 24000, seed=3)),
                             ("C8", config_c(), fast_fixed_batch(8, 7200000, seed=4))):
    plan = api.Plan(p); b = plan.batch(off)
    for dt, mis in ((torch.int16, 0), (torch.float32, 0), (torch.float32, 1)):
        d = torch.from_numpy(np.concatenate([np.zeros(mis, pcm.dtype), pcm])).cuda().to(dt)[mis:]
        out = plan.compute_batch(b, d)
        for _ in range(3): plan.compute_batch(b, d, out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): plan.compute_batch(b, d, out)
        e1.record(); torch.cuda.synchronize()
        print(name, dt, "scalar staging" if mis else "", round(b.total_frames * 20 / (e0.elapsed_time(e1) * 1e-3) / 1e6, 1), "M frames/s")
