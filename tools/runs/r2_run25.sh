#!/bin/bash
# round 2: where does fp16 lose 1 ms against bf16? per-family times of both modes on one box
mkdir -p gpurun_out
for rep in 1 2; do
for pr in fp16 bf16; do
  timeout 300 python bench.py --precision $pr --steps 20 --warmup 3 --no-modes --no-e2e --no-cpu-baseline > gpurun_out/r2ac_bench_$pr.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/r2ac_bench_$pr.json')); print('$pr', round(d['ms_per_step'],3), d['roofline']['families_ms_per_step'], d['clocks']['sm_mhz'])"
done
done
