#!/bin/bash
# Runs every GPU test file in its own process (a trapped kernel poisons the CUDA context of its process only),
# each under a timeout, logs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
rc=0
for f in tests/test_gpu_*.py; do
  n=$(basename $f .py)
  timeout 600 python -m pytest $f -q -m gpu --timeout 300 -p no:cacheprovider > gpurun_out/$n.log 2>&1
  r=$?
  echo "== $f exit $r"; tail -n ${TAIL:-25} gpurun_out/$n.log
  [ $r -ne 0 ] && rc=1
done
exit $rc
