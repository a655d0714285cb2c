#!/bin/bash
# round 2: full GPU suite with PDL default + stage-by-stage error of the fp16 mode
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2o_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2o_pytest_gpu.log
timeout 300 python tools/parity_report.py fp16 > gpurun_out/r2o_parity_fp16.log 2>&1; echo "parity exit $?"
VAPB_CONV01=0 timeout 300 python tools/parity_report.py fp16 > gpurun_out/r2o_parity_fp16_unfused.log 2>&1
grep -E "==|stage|vad|logits|probs" gpurun_out/r2o_parity_fp16.log | head -60
