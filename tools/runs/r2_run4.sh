#!/bin/bash
# round 2, GPU call 4: timeline of the fused kernel's first tiles
mkdir -p gpurun_out
timeout 300 python tools/conv01_probe.py > gpurun_out/r2d_probe_timeline.log 2>&1
echo "exit $?"
tail -20 gpurun_out/r2d_probe_timeline.log
