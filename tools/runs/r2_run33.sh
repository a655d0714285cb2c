#!/bin/bash
# round 2: 8-GPU bench of the final tree (torchrun, NCCL collectives inside the timed e2e region)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 10 --warmup 3 --no-modes > gpurun_out/r2al_bench_n8.json 2> gpurun_out/r2al_bench_n8.err; echo "bench n8 exit $?"
tail -2 gpurun_out/r2al_bench_n8.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2al_bench_n8.json'))
print('n_gpus', d['n_gpus'], 'value', round(d['value']), 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value']), 'full', round(d['e2e']['full_outputs']['value']), d['clocks'])
PY
