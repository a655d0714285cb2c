#!/bin/bash
# round 2: same-box comparison of attention kernel variants
mkdir -p gpurun_out
for rep in 1 2; do
for v in attn_prev cur; do
  if [ $v = cur ]; then unset VAPB_LIB; else export VAPB_LIB=$PWD/voiceactivityprojection_b200/libvapb_$v.so; fi
  timeout 100 python tools/attn_time.py 2>&1 | tail -1
done
done | tee gpurun_out/r2x_attn_variants.log
