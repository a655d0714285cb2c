#!/bin/bash
# round 2: 8-GPU bench (weak scaling; per-step NCCL collectives inside the timed e2e region)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 10 --warmup 3 --no-modes > gpurun_out/r2s_bench_n8.json 2> gpurun_out/r2s_bench_n8.err
echo "bench n8 exit $?"; tail -3 gpurun_out/r2s_bench_n8.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2s_bench_n8.json'))
print('value', round(d['value']), 'ms', round(d['ms_per_step'], 3), d['clocks'])
e = d['e2e']; print('e2e', round(e['value']), e['h2d_bytes_per_step'], e['d2h_bytes_per_step'], e.get('collective_bytes_per_step'))
print('full', round(e['full_outputs']['value']))
PY
