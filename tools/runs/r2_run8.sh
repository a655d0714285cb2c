#!/bin/bash
# round 2: new tests, then compute-sanitizer memcheck on the small-shape run of every mode
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "conv0_tc or extractor or full_size or pcm16 or compact or bulk" > gpurun_out/r2k_pytest_new.log 2>&1
echo "pytest new exit $?"; tail -5 gpurun_out/r2k_pytest_new.log
python tools/quick_check.py > gpurun_out/r2k_quick_plain.log 2>&1 &&
timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 python tools/quick_check.py > gpurun_out/r2k_memcheck.log 2>&1
echo "memcheck exit $?"; tail -12 gpurun_out/r2k_memcheck.log
