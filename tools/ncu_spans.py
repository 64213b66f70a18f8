#!/usr/bin/env python
"""Per-span (between BAR.SYNC) opcode histogram of executed warp-instructions from an .ncu-rep.

    python tools/ncu_spans.py gpurun_out/prof.ncu-rep [frames_per_launch]
"""
import csv, io, subprocess, sys
from collections import Counter, defaultdict

rep = sys.argv[1]
frames = float(sys.argv[2]) if len(sys.argv) > 2 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
h = rows[hi]
cs, cn, ce = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
span = 0
hist = defaultdict(Counter)
samp = Counter()
tot = Counter()
for r in rows[hi + 1:]:
    if len(r) <= ce:
        continue
    s = r[cs].strip()
    toks = s.split()
    if not toks:
        continue
    op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
    base = op.split(".")[0]
    if base in ("LDS", "STS", "LDG", "STG", "LDL", "STL"):
        w = [x for x in op.split(".") if x in ("64", "128", "U16", "S16", "U8")]
        base = base + ("." + w[0] if w else "")
    ex = int(float(r[ce] or 0))
    hist[span][base] += ex
    tot[span] += ex
    samp[span] += int(float(r[cn] or 0))
    if base == "BAR":
        span += 1
scale = 1.0 / (frames / 32.0) if frames else 1.0
unit = "warp-instr per 32-frame tile" if frames else "warp-instr"
print(f"# per-span opcode mix ({unit})")
for sp in sorted(hist):
    line = ", ".join(f"{k} {v * scale:.0f}" for k, v in hist[sp].most_common(14))
    print(f"span {sp}: total {tot[sp] * scale:.0f}, samples {samp[sp]} :: {line}")
print("all:", f"{sum(tot.values()) * scale:.0f}")
