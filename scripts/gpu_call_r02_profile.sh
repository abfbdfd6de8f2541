#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra-regimes"
$B > gpurun_out/r02_plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_default_bench.csv $B > gpurun_out/r02_ncu_bench.log 2>&1
echo "launch list rc=$?"; grep -c "sweep_fused_kernel\|gemm_pair" gpurun_out/r02_launches_default_bench.csv
python scripts/profile_targets.py > gpurun_out/r02_plain_targets.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"sweep_fused_kernel|gemm_pair_kernel|ingest_vec_kernel" -c 7 -o gpurun_out/prof_r02 python scripts/profile_targets.py > gpurun_out/r02_ncu_targets.log 2>&1
echo "full capture rc=$?"; cat gpurun_out/r02_plain_targets.log; ls -la gpurun_out/prof_r02.ncu-rep
timeout 600 python -m pytest tests -m gpu -x -q --timeout 150 2>&1 | tail -4
timeout 900 python bench.py --steps 200 --warmup 3 > gpurun_out/r02_bench_1.json 2> gpurun_out/r02_bench_1.err; echo "bench rc=$?"
