#!/bin/bash
# Runs every GPU test file in its own process (a CUDA fault in one cannot poison the others)
# and leaves one log per file under gpurun_out/.
mkdir -p gpurun_out
rc=0
for f in tests/test_gpu_*.py; do
  n=$(basename "$f" .py)
  timeout 900 python -m pytest "$f" -m gpu -q -p no:cacheprovider --timeout=600 > "gpurun_out/$n.log" 2>&1
  s=$?
  echo "== $n exit $s"; tail -n 15 "gpurun_out/$n.log"
  [ $s -ne 0 ] && rc=1
done
exit $rc
