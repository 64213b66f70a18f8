#!/bin/bash
# Iteration pass: parity tests, bench lines (A with the specialised and the ct kernel, B), then one
# ncu capture of the top kernel.  Everything lands in gpurun_out/.
MODE="$1"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu.log
for spec in "A auto bench" "A fused_ct bench_ct" "B auto bench_B"; do
  set -- $spec
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --workload $1 --kernel $2 > gpurun_out/$3.json 2> gpurun_out/$3.err; echo "$3 rc=$?"
  python - "$3" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.load(open(f'gpurun_out/{n}.json'))
    print(n, {k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')}, d['config']['kernel'], 'e2e', d['e2e']['value'],
          'frac', d['roofline']['frac'], d['clocks'])
except Exception as e:
    print(n, 'parse failed', e); print(open(f'gpurun_out/{n}.err').read()[-2000:])
PY
done
if [ "$MODE" != "noncu" ]; then
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fused -s 5 -c 1 -f -o gpurun_out/prof_fused $CMD > gpurun_out/ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
fi
