#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/attn_x3_check.py 2>&1 | grep -E "16-bit" | cut -c1-200 | tee gpurun_out/r2aj_attn_x3_check.log
timeout 200 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "attention" 2>&1 | tail -3
for rep in 1 2; do
for v in attn_prev cur; do
  if [ $v = cur ]; then unset VAPB_LIB; else export VAPB_LIB=$PWD/voiceactivityprojection_b200/libvapb_$v.so; fi
  timeout 100 python tools/attn_time.py 2>&1 | tail -1
done
done
