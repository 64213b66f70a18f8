mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err; echo "ref rc=$?"
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
wc -l gpurun_out/bench_final.json gpurun_out/bench_final_ref.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_final.json')); r=json.load(open('gpurun_out/bench_final_ref.json'))
print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline']['value'], d['gpu_launches'], d['clocks'])
print('ref', r['value'], r['e2e'])
PY
