#!/bin/bash
# round 2: attention chain timeline + attention unit tests
mkdir -p gpurun_out
timeout 200 python tools/attn_probe.py > gpurun_out/r2w_attn_probe.log 2>&1; echo "probe exit $?"; tail -90 gpurun_out/r2w_attn_probe.log | cut -c1-150
timeout 100 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "attn or attention" 2>&1 | tail -3
