#!/bin/bash
# One gpurun call: GPU parity tests, default bench (N=1), optional ncu launch list.  Usage: tools/gpu_check.sh TAG [ncu]
TAG=${1:-x}
mkdir -p gpurun_out
bash tests/run_gpu_tests.sh > gpurun_out/tests_summary.log 2>&1; echo "tests rc $?"
grep -E "^== |passed|failed|error" gpurun_out/tests_summary.log | tail -30
python bench.py --steps 10 --warmup 3 --trace-out gpurun_out/trace_$TAG.json > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench rc $?"; tail -c 1500 gpurun_out/bench_$TAG.json | cut -c1-1500
if [ "$2" = "ncu" ]; then
  python tools/profile_step.py --steps 2 > gpurun_out/plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv \
      python tools/profile_step.py --steps 2 > gpurun_out/ncu.log 2>&1
  echo "ncu rc $?"
fi
