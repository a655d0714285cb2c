#!/bin/bash
# round 2, GPU call 3 (tensor-core conv0 inside the fused kernel): fused conv0->conv1 kernel: correctness probe (both store modes), unit tests, A/B bench
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r2c_gpu.txt 2>&1
for mode in 0 1; do
  VAPB_CONV01_STORE=$mode timeout 300 python tools/conv01_probe.py > gpurun_out/r2c_probe_store$mode.log 2>&1
  echo "probe store=$mode exit $?" >> gpurun_out/r2c_status.txt
done
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k conv01 > gpurun_out/r2c_pytest_conv01.log 2>&1
echo "pytest conv01 exit $?" >> gpurun_out/r2c_status.txt
if grep -q "passed" gpurun_out/r2c_pytest_conv01.log && ! grep -q "failed" gpurun_out/r2c_pytest_conv01.log; then
  for c in 0 1; do
    VAPB_CONV01=$c timeout 600 python bench.py --steps 6 --warmup 3 --precision fp16 --no-cpu-baseline --no-e2e > gpurun_out/r2c_bench_conv01_$c.json 2> gpurun_out/r2c_bench_conv01_$c.err
    echo "bench conv01=$c exit $?" >> gpurun_out/r2c_status.txt
  done
  timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2c_pytest_gpu.log 2>&1
  echo "pytest gpu exit $?" >> gpurun_out/r2c_status.txt
fi
cat gpurun_out/r2c_status.txt
tail -5 gpurun_out/r2c_probe_store0.log gpurun_out/r2c_probe_store1.log
