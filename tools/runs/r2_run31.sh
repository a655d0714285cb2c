#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/attn_x3_check.py 2>&1 | grep -E "^nseq|16-bit|bad" | cut -c1-200 | tee gpurun_out/r2aj_attn_x3_check.log
timeout 200 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "attention" 2>&1 | tail -3
timeout 100 python tools/attn_time.py 2>&1 | tail -1
