#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
make -C cgc_net_b200/csrc -j > $OUT/build.log 2>&1 || { tail -20 $OUT/build.log; exit 1; }
export BATCH=64 M=1140 K=1140 TB=0
for n in 20 60; do
  N=$n python tools/profile_gemm.py; N=$n TA=1 python tools/profile_gemm.py
done
unset BATCH M K TB
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "gemm or tcgen05 or tf32" > $OUT/exp_ops.log 2>&1
echo "gemm tests exit=$? $(tail -1 $OUT/exp_ops.log)"
timeout 900 python -m pytest tests/test_gpu_full_size.py tests/test_gpu_model.py -m gpu -q > $OUT/exp_model.log 2>&1
echo "model exit=$? $(tail -1 $OUT/exp_model.log)"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-roofline > $OUT/exp_bench.log 2> $OUT/exp_bench.err; tail -1 $OUT/exp_bench.log | cut -c1-300
