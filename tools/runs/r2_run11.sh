#!/bin/bash
# round 2: 2-GPU bench (per-step NCCL all-gather + all-reduce inside the timed e2e region)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 --no-modes > gpurun_out/r2n_bench_n2.json 2> gpurun_out/r2n_bench_n2.err
echo "bench n2 exit $?"; tail -3 gpurun_out/r2n_bench_n2.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2n_bench_n2.json'))
print('value', round(d['value']), 'ms', round(d['ms_per_step'], 3), d['clocks'])
e = d['e2e']; print('e2e', round(e['value']), e['h2d_bytes_per_step'], e['d2h_bytes_per_step'], e.get('collective_bytes_per_step'), e['bulk'])
print('full', round(e['full_outputs']['value']))
PY
