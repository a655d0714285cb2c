#!/bin/bash
# round 2: VAD head inside the last FFN kernel: GPU suite, same-box A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2an_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2an_pytest_gpu.log
for rep in 1 2; do
for vf in 1 0; do
  VAPB_VAD_FUSED=$vf timeout 300 python bench.py --steps 20 --warmup 3 --no-modes --no-cpu-baseline > gpurun_out/r2an_bench_vf$vf.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/r2an_bench_vf$vf.json')); print('vad_fused=$vf', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), d['roofline']['families_ms_per_step']['heads'], d['gpu_launches'], d['clocks']['sm_mhz'])"
done
done
