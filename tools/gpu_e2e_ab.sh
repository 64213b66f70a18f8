#!/bin/bash
# e2e legs of the default bench for the shipped library and every build_variants/libmfcc_b200_*.so, on one box
mkdir -p gpurun_out
for rep in 1 2; do
for lib in default build_variants/libmfcc_b200_*.so; do
  if [ "$lib" = default ]; then unset MFCC_B200_LIB; else export MFCC_B200_LIB=$PWD/$lib; fi
  python bench.py --steps 3 --warmup 3 --no-cpu --extra ${1:-none} --e2e-steps 5 2>/dev/null | python -c "
import json,sys,os
d=json.loads(sys.stdin.read())
o={'lib': os.path.basename(os.environ.get('MFCC_B200_LIB','default')), 'e2e': round(d['e2e']['value']/1e6,1), 'e2e_post': round(d['e2e_post']['value']/1e6,1)}
for k,v in (d.get('workloads') or {}).items():
    if 'e2e' in v: o[k]=round(v['e2e']['value']/1e6,1)
    if v.get('e2e_g711'): o[k+'_g711']=round(v['e2e_g711']['value']/1e6,1)
print(json.dumps(o))" | tee -a gpurun_out/e2e_ab.jsonl
done; done
