#!/bin/bash
# Whole GPU suite + smoke + default bench + reference arm (what the driver runs at round end), then the ncu launch list of
# two steps and the `--set full` capture of the path's kernels.  usage: tools/gpu_final_check.sh TAG
TAG=${1:-f}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_gpu_tests.log 2>&1; echo "gpu tests rc $?"; tail -4 gpurun_out/${TAG}_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc $?"; tail -3 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc $?"; tail -2 gpurun_out/${TAG}_bench.err; head -c 400 gpurun_out/${TAG}_bench.json; echo
timeout 600 python bench.py --impl reference > gpurun_out/${TAG}_bench_reference_arm.json 2> gpurun_out/${TAG}_bench_reference_arm.err; echo "reference arm rc $?"; head -c 300 gpurun_out/${TAG}_bench_reference_arm.json; echo
python tools/profile_step.py --steps 2 > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_ncu_launches.csv \
    python tools/profile_step.py --steps 2 > gpurun_out/ncu_launches.log 2>&1
echo "ncu launch list rc $?"
bash tools/gpu_ncu_step.sh ${TAG}_step
