#!/bin/bash
# ncu --set full capture of the two post-processing kernels (after a plain run of the same command) and the launch list of
# the default bench command.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
cat > /tmp/post_leg.py <<'PY'
import json, sys, os
sys.path.insert(0, os.getcwd())
import torch, bench
from mfcc_b200 import KERNEL_AUTO
ctx = bench.Ctx(); ctx.world = 1; ctx.rank = 0; ctx.local = 0; ctx.kernel = KERNEL_AUTO
torch.cuda.set_device(0)
print(json.dumps(bench.measure_post(ctx, steps=4)))
PY
python /tmp/post_leg.py > gpurun_out/plain_post.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:post_ -s 9 -c 3 -f -o gpurun_out/r2_prof_post python /tmp/post_leg.py > gpurun_out/ncu_post.log 2>&1
echo "ncu post rc=$?"; tail -2 gpurun_out/ncu_post.log
if [ "$1" != "nolist" ]; then
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1"
$CMD > gpurun_out/plain_default.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
fi
