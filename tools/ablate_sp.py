#!/usr/bin/env python
"""Phase ablation of the fused_sp kernel (run on a disposable GPU box copy: it REWRITES mfcc_fused_sp.cu in place).

For each mask the marked phase bodies are compiled out (barriers stay), the library is rebuilt and bench.py times
configs[1].  Results are wrong by construction; the point is what each phase costs in the overlapped steady state
(the stall-sample shares of an ncu capture are not additive).
    bit 0 S0 staging   bit 1 S1 pass 1   bit 2 S2 pass 2   bit 3 S3 filterbank sums   bit 4 S3b bands + log   bit 5 S4 DCT + store
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "mfcc_b200", "csrc", "mfcc_fused_sp.cu")
orig = open(SRC).read()
markers = [("        // ---- S0: stage", "        if (fast) {", 1),
           ("        // ---- S1: pass 1.", "        {", 2),
           ("        // ---- S2: pass 2.", "        {", 4),
           ("        // ---- S3: filterbank sums.", "        {", 8),
           ("        // ---- S3b: band m", "        {", 16),
           ("        // ---- S4: log-mel rows", "        if (a.logmel) {", 32)]


def variant(mask):
    s = orig
    if mask & 1:   # S0: keep the mbarrier wait (its phase must advance), stage nothing
        old = "            const int nchunks = G::tceil_s(n_frames, sh) >> 3;\n            const float na = -a.preemph;\n            // chunk c"
        assert old in s
        s = s.replace(old, old.replace("G::tceil_s(n_frames, sh) >> 3", "0"))
    for mark, opener, bit in markers[1:]:
        i = s.index(mark)
        j = s.index("\n" + opener + "\n", i) + 1
        s = s[:j] + f"        if constexpr (({mask} & {bit}) == 0)\n" + s[j:]
    return s


rows = []
for mask in [int(x) for x in (sys.argv[1:] or ["0", "1", "2", "4", "8", "16", "32", "56", "6", "63"])]:
    open(SRC, "w").write(variant(mask))
    subprocess.run(["make", "-C", os.path.dirname(SRC), "-j4"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "20", "--warmup", "3", "--no-cpu",
                        "--e2e-steps", "1"], capture_output=True, text=True)
    if not r.stdout.strip():
        print(json.dumps({"mask": mask, "error": r.stderr[-600:]}), flush=True)
        continue
    d = json.loads(r.stdout.splitlines()[-1])
    row = {"mask": mask, "frames_per_s": d["value"], "ms_per_step": d["ms_per_step"]}
    rows.append(row)
    print(json.dumps(row), flush=True)
open(SRC, "w").write(orig)
