#!/bin/bash
# round 2: recurrence with the input half issued for two steps at a time
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "rnn" 2>&1 | tail -3
timeout 150 python tools/rnn_probe.py > gpurun_out/r2ad_rnn_probe.log 2>&1; echo "probe exit $?"; tail -12 gpurun_out/r2ad_rnn_probe.log
