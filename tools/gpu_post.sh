#!/bin/bash
# Post-processing kernels: parity tests, then the bench leg alone (bench.measure_post), optionally an ncu capture.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_post.py -m gpu -q -x --timeout 300 > gpurun_out/pytest_post.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_post.log
cat > /tmp/post_leg.py <<'PY'
import json, sys, os
sys.path.insert(0, os.getcwd())
import torch, bench
from mfcc_b200 import KERNEL_AUTO
ctx = bench.Ctx(); ctx.world = 1; ctx.rank = 0; ctx.local = 0; ctx.kernel = KERNEL_AUTO
torch.cuda.set_device(0)
print(json.dumps(bench.measure_post(ctx)))
PY
timeout 300 python /tmp/post_leg.py > gpurun_out/post_leg.json 2> gpurun_out/post_leg.err; echo "post leg rc=$?"; cat gpurun_out/post_leg.json; tail -3 gpurun_out/post_leg.err
if [ "$1" = "ncu" ]; then
  ncu --set full --clock-control none --import-source on -k regex:post_ -s 9 -c 3 -f -o gpurun_out/prof_post python /tmp/post_leg.py > gpurun_out/ncu_post.log 2>&1
  echo "ncu rc=$?"; tail -2 gpurun_out/ncu_post.log
fi
