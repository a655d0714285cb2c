#!/bin/bash
# round 2, GPU call: PCM16 direct read + counters + new BulkRunner + new bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2i_pytest_gpu.log 2>&1
echo "pytest gpu exit $?" > gpurun_out/r2i_status.txt
tail -15 gpurun_out/r2i_pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
echo "bench exit $?" >> gpurun_out/r2i_status.txt
tail -5 gpurun_out/r2i_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2i_bench.json'))
print('value', round(d['value']), 'ms', round(d['ms_per_step'], 3), d['clocks'])
print('e2e', {k: v for k, v in d['e2e'].items() if k not in ('bulk', 'api', 'full_outputs')}, 'full', d['e2e']['full_outputs']['value'])
print('roofline frac', d['roofline']['frac'], d['roofline']['families_ms_per_step'], d['roofline']['families_frac_of_peak'])
print('modes', json.dumps(d['modes'], indent=1)[:3000])
print('cpu', d['cpu_baseline'])
PY
cat gpurun_out/r2i_status.txt
