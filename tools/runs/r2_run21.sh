#!/bin/bash
# round 2: new attention kernel: GPU suite, concurrency stress (short timeouts), bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2y_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2y_pytest_gpu.log
timeout 150 python tools/stress_identical.py fp16 25 > gpurun_out/r2y_stress_fp16.log 2>&1; echo "stress fp16 exit $?"; tail -3 gpurun_out/r2y_stress_fp16.log
timeout 150 python tools/stress_identical.py bf16 25 > gpurun_out/r2y_stress_bf16.log 2>&1; echo "stress bf16 exit $?"; tail -3 gpurun_out/r2y_stress_bf16.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err; echo "bench exit $?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2y_bench.json'))
print('value', round(d['value']), 'ms', round(d['ms_per_step'], 3), d['clocks'], 'frac', round(d['roofline']['frac'], 4), 'launches', d['gpu_launches'])
print('e2e', round(d['e2e']['value']), 'full', round(d['e2e']['full_outputs']['value']))
print(d['roofline']['families_ms_per_step'])
for k, v in d['modes'].items():
    if isinstance(v, dict) and 'ms_per_step' in v:
        print(k, round(v['ms_per_step'], 2), {a: float('%.2g' % b) for a, b in v['max_abs_err_vs_oracle'].items()}, round(v['argmax_agreement'], 4))
PY
