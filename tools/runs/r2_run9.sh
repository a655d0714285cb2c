#!/bin/bash
# round 2: bit-identical outputs under concurrency with the fused encoder kernel (and with programmatic dependent launch)
mkdir -p gpurun_out
timeout 900 python tools/stress_identical.py fp16 40 > gpurun_out/r2l_stress_fp16.log 2>&1; echo "stress fp16 exit $?"
VAPB_PDL=1 timeout 900 python tools/stress_identical.py fp16 40 > gpurun_out/r2l_stress_fp16_pdl.log 2>&1; echo "stress pdl exit $?"
timeout 900 python tools/stress_identical.py bf16 20 > gpurun_out/r2l_stress_bf16.log 2>&1; echo "stress bf16 exit $?"
for p in 0 1; do VAPB_PDL=$p timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-modes > gpurun_out/r2l_bench_pdl$p.json 2>/dev/null; done
tail -4 gpurun_out/r2l_stress_fp16.log gpurun_out/r2l_stress_fp16_pdl.log gpurun_out/r2l_stress_bf16.log
python - <<'PY'
import json
for p in (0, 1):
    d = json.load(open(f'gpurun_out/r2l_bench_pdl{p}.json'))
    print('pdl', p, round(d['ms_per_step'], 3), d['clocks']['sm_mhz'], d['roofline']['families_ms_per_step'])
PY
