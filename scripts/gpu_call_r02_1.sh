#!/bin/bash
# Round 2, first GPU call: new full-size / third-party / device-feed tests, the default bench with parity_check and the
# full-corpus CPU baseline, then the experiments prepared at the end of round 1.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader; nproc; free -g | head -2
timeout 1500 python -m pytest tests/test_fullsize_gpu.py tests/test_config0_gpu.py -m gpu -x -q > gpurun_out/r02_1_fullsize.log 2>&1; echo "fullsize tests rc=$?"; tail -n 5 gpurun_out/r02_1_fullsize.log
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "thirdparty or reserve or device_feed or shim or chunk" > gpurun_out/r02_1_newparity.log 2>&1; echo "new parity tests rc=$?"; tail -n 5 gpurun_out/r02_1_newparity.log
timeout 900 python bench.py --steps 100 --warmup 3 > gpurun_out/r02_1_bench.json 2> gpurun_out/r02_1_bench.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/r02_1_bench.err
bash scripts/experiments_first_run.sh
