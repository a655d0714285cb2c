#!/bin/bash
# round 2: FFN: staggered start of odd clusters (are the final epilogues of all SMs colliding on HBM?)
mkdir -p gpurun_out
for v in "" st8 st16 "" st8 st16; do
  if [ -n "$v" ]; then export VAPB_LIB=$PWD/voiceactivityprojection_b200/libvapb_$v.so; else unset VAPB_LIB; fi
  echo "== variant '$v'"
  timeout 100 python tools/ffn_probe.py 2>&1 | tail -1
done | tee gpurun_out/r2ab_ffn_exp.log
