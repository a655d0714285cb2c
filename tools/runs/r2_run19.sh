#!/bin/bash
# round 2: attention chain timeline
mkdir -p gpurun_out
timeout 200 python tools/attn_probe.py > gpurun_out/r2w_attn_probe.log 2>&1; echo "probe exit $?"; tail -4 gpurun_out/r2w_attn_probe.log | cut -c1-150
