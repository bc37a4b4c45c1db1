#!/bin/bash
mkdir -p gpurun_out
python tools/bench_mha.py stac_speech_translation_b200/libstac_b200.so > gpurun_out/c5_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mha2_bf16 -s 6 -c 1 -o gpurun_out/c5_mha2 python tools/bench_mha.py stac_speech_translation_b200/libstac_b200.so > gpurun_out/c5_ncu.log 2>&1
echo "ncu rc $?"; tail -5 gpurun_out/c5_ncu.log; ls -la gpurun_out/c5_mha2.ncu-rep
