#!/bin/bash
# One compute-sanitizer tool per call (B200_PROFILING.md).  Small inputs: parity tests of the fused kernels only.
# NOTE (round 1): compute-sanitizer is closed on this pool — the call returns a notice instead of running; the
# alignment / option-matrix parity tests are what guards the staging and tail code instead.
TOOL="${1:-racecheck}"
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool "$TOOL" --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -m gpu \
  -k "aligned_ragged or ragged_batch_matches_oracle or host_entry or every_start_alignment or wide_kernel_option" --timeout 1400 > gpurun_out/sanitize_$TOOL.log 2>&1
echo "sanitizer($TOOL) rc=$?"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Race reported|Invalid" gpurun_out/sanitize_$TOOL.log | sort | uniq -c | head -20
