#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_fused_gpu.py -m gpu -x -q --timeout 120 -k "pinned or matches_oracle" 2>&1 | tail -3
timeout 400 python scripts/midbatch_profile.py time > gpurun_out/mid_time.log 2>&1; echo "time rc=$?"; cat gpurun_out/mid_time.log
timeout 200 python scripts/midbatch_profile.py list 128 > gpurun_out/mid_list_plain.log 2>&1 && timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/mid_launches_128.csv python scripts/midbatch_profile.py list 128 > gpurun_out/mid_ncu.log 2>&1
echo "list rc=$?"
timeout 600 python bench.py --steps 200 --warmup 3 --no-extra-regimes --also-batch 0 --no-cpu-baseline > gpurun_out/mid_bench_1.json 2> gpurun_out/mid_bench_1.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/mid_bench_1.json
