#!/usr/bin/env python
"""A/B timing of build-time variants of the fused_sp kernel ON ONE BOX (run under gpurun: it rebuilds the library
in the box's disposable copy).  Box-to-box differences on this pool reach several per cent for the 8 kHz workload,
so a variant is only judged against the default built and timed in the same call, interleaved and repeated.

    python tools/variants_sp.py [--reps 2] [--workloads A,B3] name1:"-DX=1" name2:"-DX=0 -DY=2" ...
The variant named `default` (no defines) is always included.
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "mfcc_b200", "csrc")
args = sys.argv[1:]
reps, workloads = 2, ["A", "B3"]
while args and args[0].startswith("--"):
    if args[0] == "--reps":
        reps = int(args[1])
    elif args[0] == "--workloads":
        workloads = args[1].split(",")
    args = args[2:]
variants = [("default", "")] + [tuple(a.split(":", 1)) for a in args]


def build(extra):
    for f in ("mfcc_fused_sp.o",):
        try:
            os.remove(os.path.join(CSRC, f))
        except FileNotFoundError:
            pass
    subprocess.run(["make", "-C", CSRC, f"NVEXTRA={extra}"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def bench(w):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "40", "--warmup", "5", "--no-cpu",
                        "--e2e-steps", "1", "--workload", w], capture_output=True, text=True)
    return json.loads(r.stdout.splitlines()[-1])["value"] if r.stdout.strip() else float("nan")


res = {}
for rep in range(reps):
    for name, extra in variants:
        build(extra)
        for w in workloads:
            v = bench(w)
            res.setdefault((name, w), []).append(v)
            print(json.dumps({"rep": rep, "variant": name, "defines": extra, "workload": w, "frames_per_s": v}), flush=True)
build("")
for (name, w), vs in res.items():
    print(json.dumps({"variant": name, "workload": w, "best": max(vs), "all": vs}), flush=True)
