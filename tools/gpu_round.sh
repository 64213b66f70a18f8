#!/bin/bash
# Round pass: parity tests, one bench line per BASELINE.json config (full sizes), the reference arm,
# the tensor-core GEMM bound probe, and ncu captures of the A and C kernels.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_A.json 2> gpurun_out/bench_A.err; echo "bench A rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_A_ref.json 2> gpurun_out/bench_A_ref.err; echo "ref A rc=$?"
for w in B B3 C8 C; do
  timeout 900 python bench.py --steps 10 --warmup 3 --workload $w --e2e-steps 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "bench $w rc=$?"
done
python - <<'PY'
import json
for n in ("A", "B", "B3", "C8", "C"):
    try:
        d = json.load(open(f"gpurun_out/bench_{n}.json"))
        c = d.get("cpu_baseline") or {}
        print(n, f"{d['value']:.4g} frames/s", d["kernel"], f"e2e {d['e2e']['value']:.4g}", f"frac {d['roofline']['frac']:.3f}",
              f"hbm {d['roofline_hbm']['frac']:.3f}", f"cpu {c.get('value', 0):.4g} x{c.get('cores')}", d["clocks"])
    except Exception as e:
        print(n, "parse failed", e)
PY
timeout 300 python tools/tc_dft_bound.py > gpurun_out/tc_dft_bound.jsonl 2> gpurun_out/tc_dft_bound.err; echo "tc bound rc=$?"; cat gpurun_out/tc_dft_bound.jsonl
if [ "$1" != "noncu" ]; then
for w in A B3 C8; do
  CMD="python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 --workload $w"
  $CMD > gpurun_out/plain_$w.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:fused -s 5 -c 1 -f -o gpurun_out/prof_$w $CMD > gpurun_out/ncu_$w.log 2>&1
  echo "ncu $w rc=$?"; tail -2 gpurun_out/ncu_$w.log
done
fi
# launch list of the default bench command (per-launch durations, cold-cache and serialised)
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
# phase ablation of the A kernel (build-time switch MFCC_SP_ABLATE, rebuilt and timed on this box)
if [ "$2" != "noablate" ]; then
python tools/variants_sp.py --reps 1 --workloads A s0:"-DMFCC_SP_ABLATE=1" s1:"-DMFCC_SP_ABLATE=2" s2:"-DMFCC_SP_ABLATE=4" \
  s3:"-DMFCC_SP_ABLATE=8" tail:"-DMFCC_SP_ABLATE=16" all:"-DMFCC_SP_ABLATE=31" > gpurun_out/ablate_sp.jsonl 2> gpurun_out/ablate_sp.err
echo "ablation rc=$?"; tail -7 gpurun_out/ablate_sp.jsonl
fi
