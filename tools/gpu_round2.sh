#!/bin/bash
# Round-2 GPU session (via gpurun): stages selected by $STAGES (space separated).  Logs land in gpurun_out/.
#   build ops model full smoke budget gather bench launches ncu_gather ncu_layer sanitizer
mkdir -p gpurun_out
OUT=gpurun_out
STAGES=${STAGES:-"build ops model smoke gather bench"}
has() { [[ " $STAGES " == *" $1 "* ]]; }
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/host_cores.txt; lscpu | grep "Model name" >> $OUT/host_cores.txt; free -g | head -2 >> $OUT/host_cores.txt
if has build; then python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1 || { echo BUILD_FAILED; tail -30 $OUT/build.log; }; fi
if has ops; then
  timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=short -x -k "${OPS_K:-not nothing}" > $OUT/ops.log 2>&1
  echo "ops exit=$? $(tail -1 $OUT/ops.log)"; grep -E "^E  |FAILED|Error" $OUT/ops.log | head -40
fi
if has smoke; then timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke exit=$? $(tail -1 $OUT/smoke.log)"; fi
if has model; then
  timeout 1200 python -m pytest tests/test_gpu_model.py -m gpu -q --tb=short > $OUT/model.log 2>&1; echo "model exit=$? $(tail -1 $OUT/model.log)"
  grep -E "^E  |FAILED" $OUT/model.log | head -30
fi
if has full; then
  timeout 1500 python -m pytest tests/test_gpu_full_size.py -m gpu -q --tb=short -s > $OUT/full.log 2>&1; echo "full exit=$? $(tail -1 $OUT/full.log)"
  grep -E "rel err|cluster indices|worst gradient|^E  |FAILED" $OUT/full.log | head -40
fi
if has dropin; then
  timeout 900 python -m pytest tests/test_gpu_train_dropin.py -m gpu -q --tb=short -s > $OUT/dropin.log 2>&1; echo "dropin exit=$? $(tail -1 $OUT/dropin.log)"
  grep -E "losses|cpu \[|^E  |SKIP|skipped" $OUT/dropin.log | head -10
fi
if has refarm; then
  timeout 900 python bench.py --impl reference --steps ${REF_STEPS:-3} --warmup 1 > $OUT/bench_ref.log 2>$OUT/bench_ref.err; echo "refarm exit=$?"; tail -1 $OUT/bench_ref.log | cut -c1-600
fi
if has budget; then
  timeout 900 python tools/pass_budget.py > $OUT/pass_budget.log 2>&1; echo "budget exit=$?"; cat $OUT/pass_budget.log | tail -30
fi
if has gather; then
  timeout 600 python tools/gather_lab.py > $OUT/gather_lab.log 2>&1; echo "gather exit=$?"; cat $OUT/gather_lab.log | tail -12
fi
if has bench; then
  timeout 900 python bench.py --steps ${BENCH_STEPS:-10} --warmup 3 ${BENCH_ARGS} > $OUT/bench.log 2>$OUT/bench.err; echo "bench exit=$?"; tail -1 $OUT/bench.log; tail -5 $OUT/bench.err
fi
if has configs; then
  for c in ${CONFIGS:-4 5}; do
    timeout 900 python bench.py --config $c --steps ${BENCH_STEPS:-5} --warmup 3 --no-roofline --no-cpu-baseline \
      > $OUT/bench_c${c}_n1.json 2>$OUT/bench_c${c}_n1.err; echo "config $c exit=$?"; tail -1 $OUT/bench_c${c}_n1.json | cut -c1-900; tail -3 $OUT/bench_c${c}_n1.err
  done
fi
if has launches; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $OUT/launches_r2.csv python tools/profile_step.py > $OUT/prof_step.log 2>&1; echo "launches exit=$?"
  python tools/summarize_launches.py $OUT/launches_r2.csv > $OUT/launches_r2_summary.txt 2>&1; head -40 $OUT/launches_r2_summary.txt
fi
if has ncu_step; then
  timeout 1500 ncu --section SpeedOfLight --metrics dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
    --clock-control none --profile-from-start off -f -o /tmp/step_r2 python tools/profile_step.py > $OUT/prof_step_full.log 2>&1; echo "ncu_step exit=$?"
  # the report is > 64 MB (gpurun_out is capped): summarise it here, bring back the table only
  python tools/summarize_step_ncu.py /tmp/step_r2.ncu-rep > $OUT/step_r2_kernels.txt 2>&1; head -30 $OUT/step_r2_kernels.txt
fi
if has ncu_epi; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_epi_rowpair -c 2 --profile-from-start off -f \
    -o $OUT/prof_epi_r2 python tools/profile_step.py > $OUT/prof_epi.log 2>&1; echo "ncu_epi exit=$?"; ls -la $OUT/prof_epi_r2.ncu-rep
fi
if has ncu_gather; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sage_gather_ell -s 2 -c 1 -f \
    -o $OUT/prof_gather_r2 python tools/profile_gather.py > $OUT/prof_gather.log 2>&1; echo "ncu_gather exit=$?"
fi
if has ncu_layer; then
  VARIANT=layer timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sage_layer_ell -s 2 -c 1 -f \
    -o $OUT/prof_layer_r2 python tools/profile_gather.py > $OUT/prof_layer.log 2>&1; echo "ncu_layer exit=$?"
fi
if has gemmlab; then
  # A/B of the GEMM lab switches: needs a LAB build (overwrites the release objects ON THE BOX only)
  make -C cgc_net_b200/csrc clean > /dev/null; make -C cgc_net_b200/csrc LAB=1 -j > $OUT/build_lab.log 2>&1
  : > $OUT/gemm_lab.log
  for cfg in "" "CGC_TC_BWARPS=1" "CGC_TC_AFIRST=1" "CGC_TC_BWARPS=1 CGC_TC_AFIRST=1" "CGC_TC_CHUNK=4" "CGC_TC_PAIR=0"; do
    for prec in 1 3; do
      env $cfg PREC=$prec timeout 120 python tools/profile_gemm.py >> $OUT/gemm_lab.log 2>&1
    done
  done
  env TA=1 TB=0 M=1140 N=1180 K=115200 timeout 120 python tools/profile_gemm.py >> $OUT/gemm_lab.log 2>&1
  env TA=1 TB=0 M=1140 N=1180 K=115200 CGC_TC_BWARPS=1 timeout 120 python tools/profile_gemm.py >> $OUT/gemm_lab.log 2>&1
  cat $OUT/gemm_lab.log
fi
if has gatherstages; then
  # timeline of the gather kernel: LAB build, kernel truncated at entry (1) / after the tile copy (2)
  make -C cgc_net_b200/csrc clean > /dev/null; make -C cgc_net_b200/csrc LAB=1 -j > $OUT/build_lab.log 2>&1
  : > $OUT/gather_stages.log
  for st in 0 1 2; do
    CGC_GATHER_STAGE=$st timeout 300 python tools/gather_lab.py --variants ell,ell-nopdl >> $OUT/gather_stages.log 2>&1
  done
  cat $OUT/gather_stages.log
fi
if has sanitizer; then
  for tool in memcheck racecheck synccheck; do
    timeout 900 compute-sanitizer --tool $tool --log-file $OUT/sanitizer_$tool.log \
      python tools/sanitizer_workload.py > $OUT/sanitizer_$tool.out 2>&1
    echo "sanitizer $tool exit=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $OUT/sanitizer_$tool.log | tail -1)"
  done
fi
