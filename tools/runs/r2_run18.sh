#!/bin/bash
# round 2: validation of the current tree: full GPU suite, smoke, default bench, reference arm
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2v_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2v_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2v_smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/r2v_smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2v_bench_reference.json 2>/dev/null; echo "ref exit $?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2v_bench.json'))
print('value', round(d['value']), 'ms', round(d['ms_per_step'], 3), d['clocks'], 'frac', round(d['roofline']['frac'], 4), 'launches', d['gpu_launches'])
print('e2e', round(d['e2e']['value']), 'full', round(d['e2e']['full_outputs']['value']))
print(d['roofline']['families_ms_per_step'])
for k, v in d['modes'].items():
    if isinstance(v, dict) and 'ms_per_step' in v:
        print(k, round(v['ms_per_step'], 2), {a: float('%.2g' % b) for a, b in v['max_abs_err_vs_oracle'].items()}, round(v['argmax_agreement'], 4))
r = json.load(open('gpurun_out/r2v_bench_reference.json')); print('reference', round(r['value'], 1), r['cpu_baseline']['cores'])
PY
