#!/bin/bash
# chunk-size sweep of the post-processing kernels (MFCC_POST_ROWS), bench leg only
mkdir -p gpurun_out
cat > /tmp/post_leg.py <<'PY'
import json, sys, os
sys.path.insert(0, os.getcwd())
import torch, bench
from mfcc_b200 import KERNEL_AUTO
ctx = bench.Ctx(); ctx.world = 1; ctx.rank = 0; ctx.local = 0; ctx.kernel = KERNEL_AUTO
torch.cuda.set_device(0)
r = bench.measure_post(ctx)
print(json.dumps({"rows": os.environ.get("MFCC_POST_ROWS", "default"), "ms": r["ms_per_step"], "frac": r["roofline"]["frac"], "delta_only_ms": r["delta2_only"]["ms_per_step"], "delta_only_frac": r["delta2_only"]["roofline_frac"]}))
PY
for rep in 1 2; do
for rows in default 128 192 256 384 512; do
  if [ $rows = default ]; then unset MFCC_POST_ROWS; else export MFCC_POST_ROWS=$rows; fi
  timeout 120 python /tmp/post_leg.py 2>/dev/null | tee -a gpurun_out/post_rows.jsonl
done; done
