#!/bin/bash
# Multi-GPU session (gpurun --gpus N): NCCL gradient-equality test, then bench.py configs 3 / 4 / 5 on N ranks.
mkdir -p gpurun_out
OUT=gpurun_out
N=${N:-2}
nvidia-smi --query-gpu=index,name --format=csv > $OUT/gpus_n$N.txt 2>&1
nvidia-smi topo -m >> $OUT/gpus_n$N.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1 || { echo BUILD_FAILED; tail -30 $OUT/build.log; }
timeout 900 python -m pytest tests/test_gpu_nccl.py -m gpu -q -s --tb=short > $OUT/nccl_test_n$N.log 2>&1
echo "nccl test exit=$? $(tail -1 $OUT/nccl_test_n$N.log)"; grep -E "^world|^E  " $OUT/nccl_test_n$N.log | head
for c in ${CONFIGS:-3 4 5}; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2950$c \
    bench.py --gpus $N --config $c --steps ${BENCH_STEPS:-10} --warmup 3 --no-roofline --no-cpu-baseline \
    > $OUT/bench_c${c}_n$N.json 2> $OUT/bench_c${c}_n$N.err
  echo "config $c N=$N exit=$?"; tail -1 $OUT/bench_c${c}_n$N.json | cut -c1-400; grep -E "Error|error" $OUT/bench_c${c}_n$N.err | tail -3
done
