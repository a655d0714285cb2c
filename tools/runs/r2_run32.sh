#!/bin/bash
# round 2: smoke() and the default bench invocation on the final tree
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2ak_smoke.log 2>&1; echo "smoke exit $?"; tail -6 gpurun_out/r2ak_smoke.log
timeout 600 python bench.py > gpurun_out/r2ak_bench_default.json 2> gpurun_out/r2ak_bench_default.err; echo "bench (no flags) exit $?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2ak_bench_default.json'))
print('steps', d['steps'], 'warmup', d['warmup'], 'value', round(d['value']), 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value']), 'frac', round(d['roofline']['frac'], 4), d['clocks'])
PY
