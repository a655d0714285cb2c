#!/bin/bash
# round 2: longer concurrency stress of the final tree (short timeouts: a hang must not eat the budget) + poisoned workspace
mkdir -p gpurun_out
timeout 250 python tools/stress_identical.py fp16 60 > gpurun_out/r2am_stress_fp16.log 2>&1; echo "stress fp16 exit $?"; tail -4 gpurun_out/r2am_stress_fp16.log
timeout 250 python tools/stress_identical.py fp32_tc 6 > gpurun_out/r2am_stress_fp32tc.log 2>&1; echo "stress fp32_tc exit $?"; tail -4 gpurun_out/r2am_stress_fp32tc.log
timeout 200 python tools/poison_probe.py > gpurun_out/r2am_poison.log 2>&1; echo "poison exit $?"; tail -3 gpurun_out/r2am_poison.log
