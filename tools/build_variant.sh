#!/bin/bash
# Build a VARIANT of libmfcc_b200.so here (nvcc cross-compiles without a GPU): the int16 translation unit of the
# 512- / 256-point kernel and/or the 2048-point kernel recompiled with extra -D flags, every other object taken from the
# main build.  Output: build_variants/libmfcc_b200_<name>.so (travels to the GPU box; timed there by tools/time_variants.py).
#   tools/build_variant.sh <name> "<nvcc flags>" [sp|wide|both]
set -e
name=$1; flags=$2; which=${3:-sp}
root=$(cd "$(dirname "$0")/.." && pwd)
csrc=$root/mfcc_b200/csrc
out=$root/build_variants/$name
mkdir -p $out
NV="nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -I$csrc"
objs=""
for o in mfcc_api mfcc_generic mfcc_fused_sp mfcc_fused_sp_f32 mfcc_fused_sp_g711 mfcc_fused_wide mfcc_post mfcc_tables mfcc_wav; do
  src=$csrc/$o.o
  if [ "$o" = mfcc_fused_sp ] && [ "$which" != wide ]; then
    $NV -Xptxas -v $flags -DMFCC_SP_PCM_TYPES=1 -c -o $out/$o.o $csrc/mfcc_fused_sp.cu 2> $out/$o.ptxas; src=$out/$o.o
  fi
  if [ "$o" = mfcc_fused_wide ] && [ "$which" != sp ]; then
    $NV -Xptxas -v $flags -c -o $out/$o.o $csrc/mfcc_fused_wide.cu 2> $out/$o.ptxas; src=$out/$o.o
  fi
  objs="$objs $src"
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $root/build_variants/libmfcc_b200_$name.so $objs
echo built build_variants/libmfcc_b200_$name.so
