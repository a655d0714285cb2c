#!/bin/bash
mkdir -p gpurun_out
echo "== current library"
timeout 200 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "far_key" 2>&1 | tail -3
echo "== library with the per-tile bound (expected to fail)"
VAPB_LIB=$PWD/voiceactivityprojection_b200/libvapb_attn_prev.so timeout 200 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "far_key" 2>&1 | grep -E "passed|failed|assert|Error" | tail -5
