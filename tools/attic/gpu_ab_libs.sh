#!/bin/bash
# A/B of two prebuilt libraries on one box: build_variants/lib_<name>.so are copied over the in-tree library in turn.
# usage: gpu_ab_libs.sh <variant under test> [baseline]   (parity tests run on the variant under test only)
V=${1:-fused}; B=${2:-default}
mkdir -p gpurun_out
cp mfcc_b200/libmfcc_b200.so /tmp/lib_saved.so
run() {
  cp build_variants/lib_$1.so mfcc_b200/libmfcc_b200.so
  for w in A B3 C8; do
    timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu --e2e-steps 1 --workload $w 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', '$w', d['value'], d['ms_per_step'])"
  done
}
cp build_variants/lib_$V.so mfcc_b200/libmfcc_b200.so
timeout 600 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/pytest_gpu_$V.log 2>&1; echo "pytest($V) rc=$?"; tail -2 gpurun_out/pytest_gpu_$V.log
run $V; run $B; run $V
cp /tmp/lib_saved.so mfcc_b200/libmfcc_b200.so
