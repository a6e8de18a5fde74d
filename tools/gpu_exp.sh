#!/bin/bash
# One-off runner for short GPU checks; output under gpurun_out/.
OUT=gpurun_out
mkdir -p $OUT
timeout 60 python -m pytest tests/test_gpu_nccl.py -m gpu -q -s -k "2-0 or 2-2" > $OUT/nccl_overlap_n2.log 2>&1
echo "nccl exit=$? $(tail -1 $OUT/nccl_overlap_n2.log)"; grep "^world" $OUT/nccl_overlap_n2.log
timeout 70 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --config 3 --steps 8 --warmup 3 --no-roofline --no-cpu-baseline --overlap-allreduce 2 \
  > $OUT/bench_c3_n2_overlap.json 2> $OUT/bench_c3_n2_overlap.err
echo "bench overlap exit=$?"; tail -1 $OUT/bench_c3_n2_overlap.json | cut -c1-330
