#!/bin/bash
# round 2: same-box A/B of the whole step: new attention kernel (default library) vs the two-pass one
mkdir -p gpurun_out
for rep in 1 2; do
  for v in new old; do
    if [ $v = old ]; then export VAPB_LIB=$PWD/voiceactivityprojection_b200/libvapb_attn_old.so; else unset VAPB_LIB; fi
    timeout 300 python bench.py --steps 20 --warmup 3 --no-modes --no-e2e --no-cpu-baseline > gpurun_out/r2z_bench_${v}_$rep.json 2>/dev/null
    python -c "
import json; d=json.load(open('gpurun_out/r2z_bench_${v}_$rep.json')); print('$v', round(d['ms_per_step'],3), d['roofline']['families_ms_per_step']['attention'], d['clocks']['sm_mhz'])"
  done
done
