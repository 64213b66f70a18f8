#!/bin/bash
# First GPU pass: pipe-rate microbenchmarks, parity tests, a bench line.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/smi.txt 2>&1
timeout 120 ./tools/microbench > gpurun_out/microbench.jsonl 2>&1; echo "microbench rc=$?"
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 300 python bench.py --steps 5 --warmup 3 --kernel generic --no-cpu > gpurun_out/bench_generic.json 2>> gpurun_out/bench.err; echo "bench generic rc=$?"
cat gpurun_out/bench_generic.json
cat gpurun_out/microbench.jsonl
