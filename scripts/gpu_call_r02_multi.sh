#!/bin/bash
# Multi-GPU check: parity of the sharded search on every exchange path, then the bench line at N ranks.  usage: ... <N>
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29611 scripts/sharded_check.py > gpurun_out/r02_sharded_check_$N.log 2>&1; echo "sharded_check x$N rc=$?"; grep -E "sharded x|SHARDED|Error|error" gpurun_out/r02_sharded_check_$N.log | tail -40
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 200 --warmup 3 > gpurun_out/r02_bench_$N.json 2> gpurun_out/r02_bench_$N.err; echo "bench x$N rc=$?"; tail -c 600 gpurun_out/r02_bench_$N.err; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02_bench_$N.json").read().strip().splitlines()[-1])
    print("value", round(d["value"], 1), "q/s  ms/step", round(d["ms_per_step"], 4), " e2e", round(d["e2e"]["value"], 1), " path", d.get("run", d["config"]).get("search_path"), "|", d.get("run", d["config"]).get("exchange"))
    print("parity", d["parity_check"]["ok"], d["parity_check"]["failures"], " kernel ms", round(d["roofline"]["kernel_ms_avg"], 4), "frac", round(d["roofline"]["frac"], 3))
    r = d.get("regimes", {}).get("batch_4096")
    if r: print("batch 4096:", round(r["value"]), "q/s parity", r["parity_check"]["ok"], r["exchange"])
except Exception as e:
    print("no bench line:", e)
PY
