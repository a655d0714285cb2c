#!/bin/bash
# round 2: split-fp16 tensor-core GEMM for the fp32 parity mode: unit tests, fp32 parity tests, speed A/B
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -k "gemm_x3" > gpurun_out/r2p_pytest_x3.log 2>&1; echo "pytest x3 exit $?"; tail -15 gpurun_out/r2p_pytest_x3.log
timeout 900 python -m pytest tests -x -q -m gpu -k "fp32 or golden or parity or oracle" > gpurun_out/r2p_pytest_fp32.log 2>&1; echo "pytest fp32 exit $?"; tail -15 gpurun_out/r2p_pytest_fp32.log
for tc in 1 0; do VAPB_FP32_TC=$tc timeout 300 python bench.py --steps 3 --warmup 3 --precision fp32 --no-cpu-baseline --no-e2e --no-modes > gpurun_out/r2p_bench_fp32_tc$tc.json 2>/dev/null; done
timeout 200 python tools/parity_report.py fp32 > gpurun_out/r2p_parity_fp32.log 2>&1
grep -E "==|logits|probs|vad|argmax" gpurun_out/r2p_parity_fp32.log | head -40
python - <<'PY'
import json
for tc in (1, 0):
    try:
        d = json.load(open(f'gpurun_out/r2p_bench_fp32_tc{tc}.json'))
        print('fp32_tc', tc, round(d['ms_per_step'], 2), d['roofline']['families_ms_per_step'])
    except Exception as e:
        print(tc, e)
PY
