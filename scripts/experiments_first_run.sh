#!/bin/bash
# First GPU run of the opt-in experiments prepared without a GPU at the end of round 1.
# Every check runs its stages in child processes under their own timeouts; a failure of one does not stop the others.
#   make -C ragfin_b200/csrc pdl                       # here, before the call: the .so travels with the snapshot
#   gpurun --timeout 2400 -- 'bash scripts/experiments_first_run.sh'
# Logs land in gpurun_out/{pair,seeded,pdl,pipeline}_check.log and gpurun_out/experimental_tests.log.
mkdir -p gpurun_out
export RAGFIN_EXPERIMENTAL=1   # the library refuses variants 4 / 5 and views without it
timeout 900 python scripts/pair_check.py   > gpurun_out/pair_check.log   2>&1; echo "pair_check rc=$?"
timeout 800 python scripts/seeded_check.py > gpurun_out/seeded_check.log 2>&1; echo "seeded_check rc=$?"
if [ -f ragfin_b200/csrc/libragfin_pdl.so ]; then
  timeout 900 python scripts/pdl_check.py  > gpurun_out/pdl_check.log    2>&1; echo "pdl_check rc=$?"
else
  echo "pdl_check skipped: libragfin_pdl.so not built"
fi
timeout 600 python scripts/pipeline_check.py > gpurun_out/pipeline_check.log 2>&1; echo "pipeline_check rc=$?"
RAGFIN_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_experimental_gpu.py -m gpu -q > gpurun_out/experimental_tests.log 2>&1
echo "experimental tests rc=$?"
tail -n 3 gpurun_out/pair_check.log gpurun_out/seeded_check.log gpurun_out/pdl_check.log gpurun_out/pipeline_check.log gpurun_out/experimental_tests.log 2>/dev/null
exit 0
