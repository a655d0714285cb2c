#!/bin/bash
# round 2: fp32-class tensor-core attention (mode fp32_tc): unit tests, fp32_tc parity tests, speed
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "attention_x3" 2>&1 | tail -8
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fp32_tc or tc" 2>&1 | tail -5
for x3 in 1 0; do
  VAPB_ATTN_X3=$x3 timeout 300 python bench.py --precision fp32_tc --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-modes > gpurun_out/r2ai_bench_fp32tc_x3$x3.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/r2ai_bench_fp32tc_x3$x3.json')); print('attn_x3=$x3', round(d['ms_per_step'],2), d['roofline']['families_ms_per_step'])"
done
