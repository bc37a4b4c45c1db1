#!/bin/bash
# Whole GPU suite + smoke + default bench (what the driver runs at round end), then the ncu capture of the path's kernels.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/f_gpu_tests.log 2>&1; echo "gpu tests rc $?"; tail -4 gpurun_out/f_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1; echo "smoke rc $?"; tail -3 gpurun_out/f_smoke.log
timeout 900 python bench.py > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "bench rc $?"; tail -2 gpurun_out/f_bench.err; head -c 400 gpurun_out/f_bench.json; echo
bash tools/gpu_ncu_step.sh r3h_step
