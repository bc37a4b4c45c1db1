#!/bin/bash
# One `ncu --set full` capture of the path's own kernels at the benchmark shape (S model, 64 x 30 s) with ONE encoder layer
# (each kernel of the layer appears once; the kernel-name filter keeps torch's calibration kernels out and the report
# small enough to travel back).  Summaries: python tools/ncu_extract.py gpurun_out/<tag>.ncu-rep --tag <tag>
TAG=${1:-r3h_step}
RX='regex:fbank_tc2_kernel|fbank_tc_kernel|topdb_norm_kernel|conv0_tc_kernel|gemm_bf16_kernel|gemm_wres_kernel|mha2_bf16_kernel|ffn_fused_kernel|layernorm_kernel|ctc_reduce_kernel'
mkdir -p gpurun_out
python tools/profile_step.py --steps 1 --layers 1 > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none -k "$RX" -s 1 -c 15 -f -o gpurun_out/$TAG python tools/profile_step.py --steps 1 --layers 1 > gpurun_out/ncu_$TAG.log 2>&1
echo "ncu full rc $?"; tail -2 gpurun_out/ncu_$TAG.log; ls -la gpurun_out/$TAG.ncu-rep
