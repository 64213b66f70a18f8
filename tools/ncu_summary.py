#!/usr/bin/env python
"""Summarise an .ncu-rep into the handful of numbers DESIGN.md / profiles/ quote.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--top 25] > profiles/<name>.md
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

WANT = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "smsp__inst_executed.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def run(args):
    return subprocess.run(["ncu", "-i"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 25
    rows = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    print(f"# ncu summary of `{rep}`\n")
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(f"## {d.get('Kernel Name', '?')}  (grid {d.get('Grid Size')}, block {d.get('Block Size')})\n")
        print("| metric | value | unit |\n|---|---|---|")
        for h, u, v in zip(hdr, units, r):
            if h in WANT or h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
                print(f"| {h} | {v} | {u} |")
        print()
    src = list(csv.reader(io.StringIO(run([rep, "--page", "source", "--csv"]))))
    # find the header row of the source table
    hi = next(i for i, r in enumerate(src) if "Source" in r and "# Samples" in r)
    h = src[hi]
    ci = {n: h.index(n) for n in ("Source", "# Samples", "Instructions Executed")}
    stall_cols = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
    total = 0
    recs = []
    for r in src[hi + 1:]:
        if len(r) <= max(ci.values()):
            continue
        try:
            n = int(r[ci["# Samples"]])
        except ValueError:
            continue
        total += n
        stalls = {h[i]: int(r[i]) for i in stall_cols if r[i].isdigit() and int(r[i])}
        recs.append((n, r[ci["Source"]].strip(), r[ci["Instructions Executed"]], stalls))
    print(f"## Top {top} SASS instructions by warp-stall samples (total samples {total})\n")
    print("| samples | % | executed | instruction | main stall reasons |\n|---|---|---|---|---|")
    for n, s, ex, st in sorted(recs, key=lambda x: -x[0])[:top]:
        main_st = ", ".join(f"{k[6:]}={v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
        print(f"| {n} | {100.0 * n / max(total, 1):.1f} | {ex} | `{s[:70]}` | {main_st} |")
    agg = defaultdict(int)
    for n, s, ex, st in recs:
        for k, v in st.items():
            agg[k] += v
    print("\n## Stall samples by reason\n")
    tot = sum(agg.values()) or 1
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
        print(f"- {k[6:]}: {v} ({100.0 * v / tot:.1f} %)")
    ops = defaultdict(int)
    for n, s, ex, st in recs:
        op = s.split()[0] if s and not s.startswith("@") else (s.split()[1] if len(s.split()) > 1 else "?")
        try:
            ops[op.split(".")[0]] += int(ex)
        except ValueError:
            pass
    print("\n## Executed warp-instructions by opcode\n")
    tot = sum(ops.values()) or 1
    for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:20]:
        print(f"- {k}: {v} ({100.0 * v / tot:.1f} %)")


if __name__ == "__main__":
    main()


def phases(rep):
    """Split the SASS listing at BAR.SYNC instructions and total samples / executed instructions per span."""
    src = list(csv.reader(io.StringIO(run([rep, "--page", "source", "--csv"]))))
    hi = next(i for i, r in enumerate(src) if "Source" in r and "# Samples" in r)
    h = src[hi]
    cs, cn, ce = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
    spans, cur = [], {"n": 0, "ex": 0, "ins": 0, "first": None}
    for r in src[hi + 1:]:
        if len(r) <= max(cs, cn, ce) or not r[cn].isdigit():
            continue
        cur["n"] += int(r[cn]); cur["ex"] += int(r[ce]) if r[ce].isdigit() else 0; cur["ins"] += 1
        if cur["first"] is None:
            cur["first"] = r[cs].strip()[:40]
        if "BAR.SYNC" in r[cs]:
            spans.append(cur); cur = {"n": 0, "ex": 0, "ins": 0, "first": None}
    spans.append(cur)
    tot = sum(s["n"] for s in spans) or 1
    print("\n## Spans between BAR.SYNC (in SASS order)\n")
    print("| span | SASS instrs | executed warp-instrs | stall samples | % samples |\n|---|---|---|---|---|")
    for i, s in enumerate(spans):
        print(f"| {i} | {s['ins']} | {s['ex']} | {s['n']} | {100.0 * s['n'] / tot:.1f} |")


if __name__ == "__main__" and "--phases" in sys.argv:
    phases(sys.argv[1])
