#!/bin/bash
# One-off experiment runner for the GPU box (A/B of the round-2 kernels); output under gpurun_out/.
OUT=gpurun_out
mkdir -p $OUT
make -C cgc_net_b200/csrc clean > /dev/null; make -C cgc_net_b200/csrc LAB=1 -j > $OUT/build_lab.log 2>&1 || { tail -20 $OUT/build_lab.log; exit 1; }
: > $OUT/exp_chunk.log
for c in 1024 2048 4096 8192 16384 32768; do
  echo "chunk $c" >> $OUT/exp_chunk.log
  CGC_GATHER_CHUNK=$c timeout 300 python tools/gather_lab.py --variants ell,ell-1persm >> $OUT/exp_chunk.log 2>&1
done
cat $OUT/exp_chunk.log
