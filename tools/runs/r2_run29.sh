#!/bin/bash
# round 2: 2-GPU bench of the final tree (torchrun, NCCL collectives inside the timed e2e region)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 --no-modes > gpurun_out/r2ag_bench_n2.json 2> gpurun_out/r2ag_bench_n2.err; echo "bench n2 exit $?"
tail -3 gpurun_out/r2ag_bench_n2.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2ag_bench_n2.json'))
print('n_gpus', d['n_gpus'], 'value', round(d['value']), 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value']), d['clocks'])
PY
