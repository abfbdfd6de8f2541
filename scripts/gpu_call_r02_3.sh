#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/fused_phases.py > gpurun_out/r02_3_phases.log 2>&1; cat gpurun_out/r02_3_phases.log
timeout 420 python -m pytest tests/test_fused_gpu.py -m gpu -x -q > gpurun_out/r02_3_fused_tests.log 2>&1; rc=$?; echo "fused tests rc=$rc"; tail -n 12 gpurun_out/r02_3_fused_tests.log
if [ $rc -eq 0 ]; then
  timeout 600 python scripts/fused_check.py > gpurun_out/r02_3_fused_check.log 2>&1; echo "fused_check rc=$?"; cat gpurun_out/r02_3_fused_check.log
  timeout 600 python -m pytest tests/test_fullsize_gpu.py -m gpu -x -q > gpurun_out/r02_3_fullsize.log 2>&1; echo "fullsize rc=$?"; tail -n 5 gpurun_out/r02_3_fullsize.log
fi
