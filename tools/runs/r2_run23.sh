#!/bin/bash
# round 2: fused FFN timeline + unit tests
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "ffn" 2>&1 | tail -3
timeout 150 python tools/ffn_probe.py > gpurun_out/r2aa_ffn_probe.log 2>&1; echo "probe exit $?"; tail -16 gpurun_out/r2aa_ffn_probe.log
