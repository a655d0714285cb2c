#!/bin/bash
# round 2: tightened fp32 tolerances: parity tests + measured errors per golden case
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2af_pytest_parity.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/r2af_pytest_parity.log
VAPB_PRECISION=fp32 timeout 300 python tools/parity_report.py fp32 > gpurun_out/r2af_parity_fp32.log 2>&1
grep -E "==|logits|vad_lg|loss" gpurun_out/r2af_parity_fp32.log | cut -c1-110 | head -40
