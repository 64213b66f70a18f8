#!/bin/bash
# The 8-GPU pass (run with gpurun --gpus 8): NCCL parity test over every GPU, the box's host -> device ceiling at
# N = 1, 2, 4, 8, the weak-scaling bench line and the strong-scaling --shard line at N = 8 (and N = 4, 2).
mkdir -p gpurun_out
N=${1:-8}
timeout 300 python -m pytest tests/test_distributed.py -m gpu -q > gpurun_out/r2_n8_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2_n8_pytest.log
timeout 120 python tools/h2d_ceiling.py > gpurun_out/r2_h2d_n1.json 2>/dev/null; cat gpurun_out/r2_h2d_n1.json
for n in 2 4 8; do
  [ $n -le $N ] || continue
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n tools/h2d_ceiling.py > gpurun_out/r2_h2d_n$n.json 2>/dev/null; cat gpurun_out/r2_h2d_n$n.json
done
for n in $N; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/r2_bench_n$n.json 2> gpurun_out/r2_bench_n$n.err; echo "bench n=$n rc=$?"
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2971$n bench.py --gpus $n --shard > gpurun_out/r2_shard_n$n.json 2> gpurun_out/r2_shard_n$n.err; echo "shard n=$n rc=$?"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_bench_n*.json")):
    try:
        d = json.load(open(f)); print(f, "value %.4g" % d["value"], "e2e %.4g" % d["e2e"]["value"], d["e2e"]["step_ms"])
    except Exception as e: print(f, "parse failed", e)
for f in sorted(glob.glob("gpurun_out/r2_shard_n*.json")):
    try:
        d = json.load(open(f)); print(f, "n", d["n_gpus"], "value %.4g" % d["value"], "ms %.3f" % d["ms_per_step"], "gather ms", d["collective"]["ms"], "with gather %.4g" % d["value_with_gather"])
    except Exception as e: print(f, "parse failed", e)
PY
