#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
make -C cgc_net_b200/csrc -j > $OUT/build.log 2>&1 || { tail -20 $OUT/build.log; exit 1; }
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "sage_dh or epilogue or two_per_sm" > $OUT/exp_ops.log 2>&1
echo "ops exit=$? $(tail -1 $OUT/exp_ops.log)"
timeout 900 python -m pytest tests/test_gpu_full_size.py tests/test_gpu_model.py -m gpu -q -s > $OUT/exp_model.log 2>&1
echo "model exit=$? $(tail -1 $OUT/exp_model.log)"; grep "worst gradient" $OUT/exp_model.log | cut -c1-200
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/exp_bench.log 2> $OUT/exp_bench.err; tail -1 $OUT/exp_bench.log | cut -c1-300
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --log-file $OUT/sanitizer_$tool.log python tools/sanitizer_workload.py > $OUT/sanitizer_$tool.out 2>&1
  echo "sanitizer $tool exit=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $OUT/sanitizer_$tool.log | tail -1)"
done
