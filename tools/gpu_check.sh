#!/bin/bash
# Quick whole-suite check on one box: GPU tests, smoke, default bench line (the driver's round-end sequence minus the reference arm).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_check.json 2> gpurun_out/bench_check.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_check.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_check.json'))
print("A", d['value'], "e2e", d['e2e']['value'], "e2e_post", d.get('e2e_post'), "frac", d['roofline']['frac'], d['clocks'])
for k, v in d['workloads'].items():
    print(k, v.get('value'), v.get('ms_per_step'), (v.get('roofline') or {}).get('frac'))
PY
