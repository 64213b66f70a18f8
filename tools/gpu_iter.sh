#!/bin/bash
# Iteration pass: parity tests then a bench line (and optionally workload B).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench.json'))
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['config']['kernel'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], d['clocks'])
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/bench.err').read()[-2000:])
PY
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --workload B > gpurun_out/bench_B.json 2>> gpurun_out/bench.err; echo "benchB rc=$?"
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench_B.json'))
    print('B', {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['config']['kernel'], 'frac', d['roofline']['frac'])
except Exception as e:
    print('benchB parse failed', e)
PY
