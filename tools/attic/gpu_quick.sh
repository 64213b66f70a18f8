mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu.log
for w in A B3; do
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --e2e-steps 2 --workload $w > gpurun_out/q_$w.json 2> gpurun_out/q_$w.err; echo "bench $w rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/q_$w.json')); print('$w', d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks'])"
done
