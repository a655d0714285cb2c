#!/bin/bash
# round 2: attention softmax with packed FMAs + compile-time 16-bit format: tests and bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu -k "attention or tensor_modes or fused_head or golden" > gpurun_out/r2u_pytest.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/r2u_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-modes --no-e2e > gpurun_out/r2u_bench.json 2>/dev/null
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2u_bench.json'))
print(round(d['ms_per_step'], 3), round(d['value']), d['clocks']['sm_mhz'], d['roofline']['families_ms_per_step'])
PY
