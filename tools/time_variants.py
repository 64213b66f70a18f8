#!/usr/bin/env python
"""Kernel-only A/B timing of library builds ON ONE BOX (run under gpurun).  Each library (a path, or a variant name
under build_variants/) is loaded in its own process through MFCC_B200_LIB, times the named workloads with CUDA events
(inputs resident, 5 warm-up launches, `steps` timed launches, best of `inner` rounds) and prints one JSON line per
(library, workload); the whole list is walked `reps` times so that drift of the box shows up as disagreement
between the passes.  Results are also checked against the first library's output (max abs difference).

    python tools/time_variants.py [--reps 2] [--steps 30] [--workloads A,B3,C8] default name1 name2 ...
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import json, os, sys, hashlib
import numpy as np, torch
sys.path.insert(0, os.environ["MFCC_ROOT"])
import bench
from mfcc_b200 import api, CONFIGS
steps = int(sys.argv[1]); names = sys.argv[2].split(",")
for w in names:
    cfg, maker, desc, _ = bench.WORKLOADS[w]
    p = CONFIGS[cfg]()
    plan = api.Plan(p)
    pcm, off = maker(1000)
    b = plan.batch(off)
    d = torch.from_numpy(pcm).cuda()
    out = torch.empty((b.total_frames, plan.out_dim), dtype=torch.float32, device="cuda")
    s = torch.cuda.current_stream()
    for _ in range(5):
        plan.compute_batch(b, d, out, s)
    torch.cuda.synchronize()
    best = 1e30
    for r in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(steps):
            plan.compute_batch(b, d, out, s)
        e1.record(s)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / steps)
    o = out[:: max(1, b.total_frames // 4096)].cpu().numpy()
    print(json.dumps({"workload": w, "ms": best, "frames_per_s": b.total_frames / (best * 1e-3), "kernel": plan.kernel_name,
                      "finite": bool(np.isfinite(o).all()), "checksum": float(np.abs(o).sum())}), flush=True)
    del d, out, b, plan
    torch.cuda.empty_cache()
'''


def main():
    args = sys.argv[1:]
    reps, steps, workloads = 2, 30, "A"
    while args and args[0].startswith("--"):
        if args[0] == "--reps":
            reps = int(args[1])
        elif args[0] == "--steps":
            steps = int(args[1])
        elif args[0] == "--workloads":
            workloads = args[1]
        args = args[2:]
    libs = []
    for a in args:
        if a == "default":
            libs.append((a, os.path.join(ROOT, "mfcc_b200", "libmfcc_b200.so")))
        elif os.path.exists(a):
            libs.append((os.path.basename(a), os.path.abspath(a)))
        else:
            libs.append((a, os.path.join(ROOT, "build_variants", f"libmfcc_b200_{a}.so")))
    res = {}
    for rep in range(reps):
        for name, path in libs:
            env = dict(os.environ, MFCC_B200_LIB=path, MFCC_ROOT=ROOT)
            r = subprocess.run([sys.executable, "-c", CHILD, str(steps), workloads], env=env, capture_output=True, text=True)
            if r.returncode != 0:
                print(json.dumps({"variant": name, "error": r.stderr[-400:]}), flush=True)
                continue
            for line in r.stdout.splitlines():
                d = json.loads(line)
                d.update(variant=name, rep=rep)
                res.setdefault((name, d["workload"]), []).append(d)
                print(json.dumps(d), flush=True)
    base = {}
    print("---- summary (best of passes; checksum relative to the first library) ----")
    for (name, w), ds in res.items():
        best = max(d["frames_per_s"] for d in ds)
        base.setdefault(w, (best, ds[0]["checksum"]))
        print(f"{name:28s} {w:4s} {best / 1e9:8.4f} G frames/s  x{best / base[w][0]:.4f}  all={[round(d['frames_per_s'] / 1e9, 4) for d in ds]}"
              f"  checksum_rel={ds[0]['checksum'] / base[w][1]:.7f} finite={all(d['finite'] for d in ds)}", flush=True)


if __name__ == "__main__":
    main()
