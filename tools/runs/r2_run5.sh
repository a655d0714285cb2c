#!/bin/bash
# round 2, GPU call 5: fused kernel with early X + cheap arrives: timeline, unit tests, A/B bench, full gpu tests
mkdir -p gpurun_out
VAPB_LIB=$PWD/voiceactivityprojection_b200/libvapb_timers.so timeout 300 python tools/conv01_probe.py > gpurun_out/r2h_probe_timeline.log 2>&1
echo "probe exit $?" > gpurun_out/r2h_status.txt
head -14 gpurun_out/r2h_probe_timeline.log
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k conv01 > gpurun_out/r2h_pytest_conv01.log 2>&1
echo "pytest conv01 exit $?" >> gpurun_out/r2h_status.txt
if grep -q "passed" gpurun_out/r2h_pytest_conv01.log && ! grep -q "failed" gpurun_out/r2h_pytest_conv01.log; then
  for c in 0 1; do
    VAPB_CONV01=$c timeout 600 python bench.py --steps 6 --warmup 3 --precision fp16 --no-cpu-baseline --no-e2e > gpurun_out/r2h_bench_conv01_$c.json 2> gpurun_out/r2h_bench_conv01_$c.err
    echo "bench conv01=$c exit $?" >> gpurun_out/r2h_status.txt
  done
  timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2h_pytest_gpu.log 2>&1
  echo "pytest gpu exit $?" >> gpurun_out/r2h_status.txt
fi
cat gpurun_out/r2h_status.txt
python - <<'PY'
import json
for c in (0, 1):
    try:
        d = json.load(open(f'gpurun_out/r2h_bench_conv01_{c}.json'))
        print(c, round(d['ms_per_step'], 3), d['clocks'], d['roofline']['families_ms_per_step'])
    except Exception as e:
        print(c, 'no bench', e)
PY
