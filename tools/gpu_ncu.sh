#!/bin/bash
# ncu capture of the fused kernel (one launch) after a plain run of the same command.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fused -s 5 -c 1 -f -o gpurun_out/prof_fused $CMD > gpurun_out/ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu.log; cat gpurun_out/plain.log | cut -c1-300
