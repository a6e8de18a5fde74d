#!/bin/bash
# One-off experiment runner for the GPU box (A/B of the round-2 element-wise kernels); output under gpurun_out/.
OUT=gpurun_out
mkdir -p $OUT
make -C cgc_net_b200/csrc -j > $OUT/build.log 2>&1 || { tail -20 $OUT/build.log; exit 1; }
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "epilogue or sage_dh or wide or batchnorm or sage_layer" > $OUT/exp_ops.log 2>&1
echo "ops exit=$? $(tail -1 $OUT/exp_ops.log)"
timeout 300 python tools/epi_lab.py > $OUT/exp_epi.log 2>&1; cat $OUT/exp_epi.log
timeout 900 python -m pytest tests/test_gpu_full_size.py tests/test_gpu_model.py -m gpu -q -s > $OUT/exp_model.log 2>&1
echo "model exit=$? $(tail -1 $OUT/exp_model.log)"; grep "worst gradient" $OUT/exp_model.log | cut -c1-250
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/exp_bench.log 2> $OUT/exp_bench.err; tail -1 $OUT/exp_bench.log | cut -c1-300
