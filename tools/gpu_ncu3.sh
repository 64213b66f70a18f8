#!/bin/bash
# One ncu --set full capture per hot kernel (A, B3, C8 workloads), each after a plain run of the same command, and the
# launch list of the default bench command.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
for w in ${1:-A B3 C8}; do
  CMD="python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 --extra none --workload $w"
  $CMD > gpurun_out/plain_$w.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:fused -s 5 -c 1 -f -o gpurun_out/r2_prof_$w $CMD > gpurun_out/ncu_$w.log 2>&1
  echo "ncu $w rc=$?"; tail -2 gpurun_out/ncu_$w.log
done
if [ "$2" != "nolist" ]; then
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1"
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
fi
