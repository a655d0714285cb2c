#!/bin/bash
# round 2: timing-only FFN experiments (variants produce wrong numbers on purpose): what bounds a chunk?
mkdir -p gpurun_out
for v in "" halfw halfg2 both; do
  if [ -n "$v" ]; then export VAPB_LIB=$PWD/voiceactivityprojection_b200/libvapb_$v.so; else unset VAPB_LIB; fi
  echo "== variant '$v'"
  timeout 100 python tools/ffn_probe.py 2>&1 | tail -4
done | tee gpurun_out/r2ab_ffn_exp.log
