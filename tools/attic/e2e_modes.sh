#!/bin/bash
# e2e step time of mfcc_compute_host per workload (median of 8 calls)
for w in A C8 B3; do
python bench.py --steps 3 --warmup 3 --no-cpu --workload $w --e2e-steps 8 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('$w', round(d['e2e']['value']/1e6,1), 'M frames/s, median step ms', sorted(d['e2e']['step_ms'])[4], d['e2e']['matches_device_path'])"
done
