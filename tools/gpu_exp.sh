#!/bin/bash
# One-off experiment runner for the GPU box (A/B of the round-2 kernels); output under gpurun_out/.
OUT=gpurun_out
mkdir -p $OUT
make -C cgc_net_b200/csrc -j > $OUT/build.log 2>&1 || { tail -20 $OUT/build.log; exit 1; }
timeout 300 python tools/epi_lab.py > $OUT/exp_epi.log 2>&1; grep sage_dh $OUT/exp_epi.log
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "sage_dh or epilogue or two_per_sm" > $OUT/exp_ops2.log 2>&1
echo "ops2 exit=$? $(tail -1 $OUT/exp_ops2.log)"
timeout 900 compute-sanitizer --tool racecheck --log-file $OUT/sanitizer_racecheck.log python tools/sanitizer_workload.py > $OUT/sanitizer_racecheck.out 2>&1
echo "racecheck exit=$? $(grep -E 'RACECHECK SUMMARY' $OUT/sanitizer_racecheck.log | tail -1)"
STAGES="ncu_epi" bash tools/gpu_round2.sh
