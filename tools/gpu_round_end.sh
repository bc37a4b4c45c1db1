#!/bin/bash
# What the driver runs at round end (GPU suite, smoke, default bench, reference arm) plus the other configs at one GPU.
# usage: tools/gpu_round_end.sh TAG
TAG=${1:-z}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_gpu_tests.log 2>&1; echo "gpu tests rc $?"; tail -3 gpurun_out/${TAG}_gpu_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc $?"; tail -2 gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc $?"; tail -2 gpurun_out/${TAG}_bench.err; head -c 300 gpurun_out/${TAG}_bench.json; echo
timeout 400 python bench.py --impl reference > gpurun_out/${TAG}_bench_reference_arm.json 2> gpurun_out/${TAG}_bench_reference_arm.err; echo "reference arm rc $?"; head -c 250 gpurun_out/${TAG}_bench_reference_arm.json; echo
for c in 4 2 3; do
  timeout 300 python bench.py --config $c > gpurun_out/${TAG}_bench_cfg${c}_n1.json 2> gpurun_out/${TAG}_bench_cfg${c}.err; echo "config $c rc $?"; head -c 200 gpurun_out/${TAG}_bench_cfg${c}_n1.json; echo
done
