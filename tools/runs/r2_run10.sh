#!/bin/bash
# round 2: concurrency stress after the window hand-over fix (short timeouts: a hang must not eat the budget)
mkdir -p gpurun_out
timeout 150 python tools/stress_identical.py fp16 25 > gpurun_out/r2m_stress_fp16.log 2>&1; echo "stress fp16 exit $?"
timeout 150 python tools/stress_identical.py bf16 25 > gpurun_out/r2m_stress_bf16.log 2>&1; echo "stress bf16 exit $?"
VAPB_PDL=1 timeout 150 python tools/stress_identical.py fp16 25 > gpurun_out/r2m_stress_fp16_pdl.log 2>&1; echo "stress pdl exit $?"
cat gpurun_out/r2m_stress_fp16.log gpurun_out/r2m_stress_bf16.log gpurun_out/r2m_stress_fp16_pdl.log
