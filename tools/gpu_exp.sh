#!/bin/bash
# One-off runner for short GPU checks; output under gpurun_out/.
OUT=gpurun_out
mkdir -p $OUT
make -C cgc_net_b200/csrc -j > $OUT/build.log 2>&1 || { tail -20 $OUT/build.log; exit 1; }
timeout 400 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "sage_dh or epilogue or sage_layer_fwd" > $OUT/exp_ops.log 2>&1
echo "ops exit=$? $(tail -1 $OUT/exp_ops.log)"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
