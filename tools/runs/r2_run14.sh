#!/bin/bash
# round 2: full GPU suite with the fp32_tc mode, then the default bench line
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2q_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/r2q_pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err; echo "bench exit $?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2q_bench.json'))
print('value', round(d['value']), 'ms', round(d['ms_per_step'], 3), d['clocks'], 'frac', round(d['roofline']['frac'], 4))
print('e2e', round(d['e2e']['value']), 'full', round(d['e2e']['full_outputs']['value']))
for k, v in d['modes'].items():
    if isinstance(v, dict) and 'ms_per_step' in v:
        print(k, round(v['ms_per_step'], 2), {a: float('%.2g' % b) for a, b in v['max_abs_err_vs_oracle'].items()}, v['argmax_agreement'], v['vad_threshold_agreement'])
print(d['modes'].get('eager_b200'))
PY
