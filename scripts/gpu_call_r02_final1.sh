#!/bin/bash
# 1 GPU: the whole GPU suite, smoke, the default bench line
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 200 2>&1 | tail -4
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py --steps 200 --warmup 3 > gpurun_out/final_bench_1.json 2> gpurun_out/final_bench_1.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/final_bench_1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'], d['roofline']['frac'], d['roofline']['kernel_ms_avg'], d['parity_check']['ok'], d.get('cpu_baseline',{}).get('value'))
for k in d:
    if k.startswith('also') or k in ('clustered','ingest'): print(k, json.dumps(d[k])[:400])
PY
