#!/bin/bash
# round 2: fused vap_head + probs kernel: tests and same-box A/B
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2t_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -12 gpurun_out/r2t_pytest_gpu.log
for f in 0 1; do VAPB_HEAD_FUSED=$f timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-modes > gpurun_out/r2t_bench_head$f.json 2>/dev/null; done
python - <<'PY'
import json
for f in (0, 1):
    try:
        d = json.load(open(f'gpurun_out/r2t_bench_head{f}.json'))
        print('head_fused', f, round(d['ms_per_step'], 3), round(d['value']), 'e2e', round(d['e2e']['value']), 'full', round(d['e2e']['full_outputs']['value']), d['clocks']['sm_mhz'], d['roofline']['families_ms_per_step'])
    except Exception as e:
        print(f, e)
PY
