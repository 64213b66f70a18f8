#!/usr/bin/env python
"""Opcode histogram + resource usage of the shipped kernels, straight from the library's SASS (no GPU needed).

    python tools/sass_hist.py [--lib mfcc_b200/libmfcc_b200.so] [--out profiles] [--tag r2]

For each hot kernel (the BASELINE.json shapes: fused_sp <short,400,160,32,16,26,13>, <short,200,80,16,16,20,13>,
fused_wide <short,1200,480,80>) writes profiles/<tag>_sass_<name>.md: registers / spills / shared memory from
`cuobjdump -res-usage`, the opcode histogram of `cuobjdump -sass`, and the counts of the instruction classes the design
claims rest on (bulk copy + mbarrier, constant-bank loads by width, FP32 mix, tensor-core / TMEM opcodes = 0).
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOT = {
    "sp_A": ("fused_sp_kernel", ["Is", "Li400", "Li160", "Li32", "Li16", "Li26", "Li13"]),
    "sp_B": ("fused_sp_kernel", ["Is", "Li200", "Li80", "Li16", "Li16", "Li20", "Li13"]),
    "wide_C": ("fused_wide_kernel", ["Is", "Li1200", "Li480", "Li80"]),
    # post-processing (mfcc_post.cu): matched by an exact substring of the mangled name
    "post_apply": ("post_apply_kernel", "post_apply_kernelILi2EE"),
    "post_stats": ("post_stats_kernel", "post_stats_kernel"),
}


def main():
    args = sys.argv[1:]
    lib, out, tag = os.path.join(ROOT, "mfcc_b200", "libmfcc_b200.so"), os.path.join(ROOT, "profiles"), "r2"
    while args:
        if args[0] == "--lib":
            lib = args[1]
        elif args[0] == "--out":
            out = args[1]
        elif args[0] == "--tag":
            tag = args[1]
        args = args[2:]
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True, check=True).stdout
    # split the SASS dump by function
    funcs, cur = {}, None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
        elif cur is not None:
            funcs[cur].append(line)
    usage = {}
    lines = res.splitlines()
    for i, line in enumerate(lines):
        m = re.match(r"\s*Function (\S+):", line)
        if m and i + 1 < len(lines):
            usage[m.group(1)] = lines[i + 1].strip()
    for name, (stem, keys) in HOT.items():
        if isinstance(keys, str):
            cands = [f for f in funcs if keys in f]
        else:
            cands = [f for f in funcs if stem in f and all(k in f for k in keys)]
            # template args appear in order: require the exact sequence
            seq = "I" + keys[0][1:] + "".join(k + "E" for k in keys[1:]) + "E"     # e.g. IsLi400ELi160E...Li13EE
            cands = [f for f in cands if seq in f]
        if not cands:
            print("not found:", name)
            continue
        fn = sorted(cands, key=len)[0]
        ops = collections.Counter()
        n = 0
        for line in funcs[fn]:
            m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if m:
                ops[m.group(1)] += 1
                n += 1
        base = collections.Counter()
        for op, c in ops.items():
            base[op.split(".")[0]] += c
        def count(pred):
            return sum(c for op, c in ops.items() if pred(op))
        claims = [
            ("UBLKCP (cp.async.bulk, TMA engine, 1-D)", count(lambda o: o.startswith("UBLKCP"))),
            ("SYNCS (mbarrier arrive / expect_tx / try_wait)", count(lambda o: o.startswith("SYNCS"))),
            ("BAR (named barriers)", count(lambda o: o.startswith("BAR"))),
            ("LDC (per-thread constant bank loads), all widths", count(lambda o: o.startswith("LDC") and not o.startswith("LDCU"))),
            ("  LDC.64", count(lambda o: o.startswith("LDC.64"))),
            ("  LDC.128", count(lambda o: o.startswith("LDC.128"))),
            ("LDCU (uniform constant loads), all widths", count(lambda o: o.startswith("LDCU"))),
            ("  LDCU.64", count(lambda o: o.startswith("LDCU.64"))),
            ("  LDCU.128", count(lambda o: o.startswith("LDCU.128"))),
            ("LDS / STS (shared memory)", f"{count(lambda o: o.startswith('LDS'))} / {count(lambda o: o.startswith('STS'))}"),
            ("LDG / STG (global)", f"{count(lambda o: o.startswith('LDG'))} / {count(lambda o: o.startswith('STG'))}"),
            ("LDL / STL (local memory: spills)", f"{count(lambda o: o.startswith('LDL'))} / {count(lambda o: o.startswith('STL'))}"),
            ("FADD / FFMA / FMUL", f"{base['FADD']} / {base['FFMA']} / {base['FMUL']}"),
            ("FADD2 / FFMA2 / FMUL2 (packed)", f"{base['FADD2']} / {base['FFMA2']} / {base['FMUL2']}"),
            ("MUFU (lg2)", base["MUFU"]),
            ("I2F / F2I (conversion pipe)", f"{base['I2F'] + base['I2FP']} / {base['F2I']}"),
            ("UTCHMMA / UTCQMMA / UTCMMA (tcgen05.mma)", count(lambda o: o.startswith("UTC") and "MMA" in o)),
            ("LDTM / STTM (tcgen05.ld / st, tensor memory)", count(lambda o: o.startswith("LDTM") or o.startswith("STTM"))),
            ("UTMALDG / UTMASTG (tensor-map TMA)", count(lambda o: o.startswith("UTMA"))),
            ("HMMA / IMMA (mma.sync)", base["HMMA"] + base["IMMA"]),
        ]
        path = os.path.join(out, f"{tag}_sass_{name}.md")
        with open(path, "w") as f:
            f.write(f"# SASS of the shipped `{stem}` kernel — {name} ({tag})\n\n")
            f.write(f"Source: `cuobjdump -sass` / `-res-usage` of `{os.path.relpath(lib, ROOT)}` (sm_100a only), function\n`{fn}`.\n\n")
            f.write(f"Resource usage: `{usage.get(fn, '?')}`\n\n")
            f.write(f"Static instructions: {n} ({n * 16 / 1024:.1f} KB of code).  Counts below are STATIC (instructions in the binary, "
                    "not executed counts; the executed mix per tile is in the ncu summaries).\n\n")
            f.write("## What the design claims rest on\n\n| instruction class | static count |\n|---|---|\n")
            for k, v in claims:
                f.write(f"| {k} | {v} |\n")
            f.write("\n## Opcode histogram (base opcode, static)\n\n| opcode | count | % |\n|---|---|---|\n")
            for op, c in base.most_common(40):
                f.write(f"| {op} | {c} | {100.0 * c / n:.1f} |\n")
            f.write("\n## Full opcodes with modifiers (top 60)\n\n| opcode | count |\n|---|---|\n")
            for op, c in ops.most_common(60):
                f.write(f"| {op} | {c} |\n")
        print("wrote", path, n, "instructions", usage.get(fn, ""))


if __name__ == "__main__":
    main()
