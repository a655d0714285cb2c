#!/bin/bash
# round 2, final evidence: GPU suite, smoke, bench + reference arm, then the ncu launch list with DRAM bytes and
# --set full of one launch of each main kernel (CSV exported on the box)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2ae_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/r2ae_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2ae_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/r2ae_smoke.log
timeout 150 python tools/stress_identical.py fp16 25 > gpurun_out/r2ae_stress_fp16.log 2>&1; echo "stress fp16 exit $?"; tail -3 gpurun_out/r2ae_stress_fp16.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2ae_bench.json 2> gpurun_out/r2ae_bench.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2ae_bench_reference.json 2>/dev/null; echo "ref exit $?"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-modes"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv \
    --log-file gpurun_out/r2ae_launches.csv $CMD > gpurun_out/r2ae_ncu1.log 2>&1
echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none -k regex:'attention_tc|rnn_tc|gemm_lin|ffn_fused|gemm_2sm|conv01' -s 60 -c 16 -o /tmp/r2ae_full $CMD > gpurun_out/r2ae_ncu2.log 2>&1
echo "ncu full exit $?"
ncu -i /tmp/r2ae_full.ncu-rep --page raw --csv > gpurun_out/r2ae_full_raw.csv 2>/dev/null
du -sh gpurun_out
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2ae_bench.json'))
print('value', round(d['value']), 'ms', round(d['ms_per_step'], 3), d['clocks'], 'frac', round(d['roofline']['frac'], 4), 'launches', d['gpu_launches'])
print('e2e', round(d['e2e']['value']), 'full', round(d['e2e']['full_outputs']['value']))
print(d['roofline']['families_ms_per_step'])
r = json.load(open('gpurun_out/r2ae_bench_reference.json')); print('reference', round(r['value'], 1), r['cpu_baseline']['cores'])
PY
