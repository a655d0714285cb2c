#!/bin/bash
# round 2, GPU call 2: is the fused kernel's producer bound by the parameter-bank (LDCU) weight stream?
mkdir -p gpurun_out
VAPB_LIB=$PWD/voiceactivityprojection_b200/libvapb_smalltab.so timeout 300 python bench.py --steps 4 --warmup 3 --precision fp16 --no-cpu-baseline --no-e2e > gpurun_out/r2b_bench_smalltab.json 2> gpurun_out/r2b_bench_smalltab.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2b_bench_smalltab.json'))
print(d['ms_per_step'], d['roofline']['families_ms_per_step'])
PY
