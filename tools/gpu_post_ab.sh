#!/bin/bash
# A/B of post-kernel builds on one box: the default library and every build_variants/libmfcc_b200_post*.so, bench leg only
mkdir -p gpurun_out
cat > /tmp/post_leg.py <<'PY'
import json, sys, os
sys.path.insert(0, os.getcwd())
import torch, bench
from mfcc_b200 import KERNEL_AUTO
ctx = bench.Ctx(); ctx.world = 1; ctx.rank = 0; ctx.local = 0; ctx.kernel = KERNEL_AUTO
torch.cuda.set_device(0)
r = bench.measure_post(ctx)
print(json.dumps({"lib": os.path.basename(os.environ.get("MFCC_B200_LIB", "default")), "ms": r["ms_per_step"], "frac": r["roofline"]["frac"], "delta_only_ms": r["delta2_only"]["ms_per_step"], "delta_only_frac": r["delta2_only"]["roofline_frac"]}))
PY
for rep in 1 2; do
  unset MFCC_B200_LIB; timeout 120 python /tmp/post_leg.py 2>/dev/null | tee -a gpurun_out/post_ab.jsonl
  for lib in build_variants/libmfcc_b200_post*.so; do
    MFCC_B200_LIB=$PWD/$lib timeout 120 python /tmp/post_leg.py 2>/dev/null | tee -a gpurun_out/post_ab.jsonl
  done
done
