#!/bin/bash
mkdir -p gpurun_out
libs="stac_speech_translation_b200/libstac_b200.so"
for n in el mc4 off8 off16 all; do libs="$libs stac_speech_translation_b200/libstac_b200_$n.so"; done
timeout 200 python tools/bench_mha.py $libs > gpurun_out/c4_mha_bench.log 2>&1; echo "bench_mha rc $?"; grep v2 gpurun_out/c4_mha_bench.log
STAC_B200_LIB=$PWD/stac_speech_translation_b200/libstac_b200_all.so timeout 300 python -m pytest tests/test_gpu_tc_attention.py -q -x -m gpu -k v2 > gpurun_out/c4_v2_all_tests.log 2>&1; echo "variant all tests rc $?"; tail -2 gpurun_out/c4_v2_all_tests.log
timeout 900 python bench.py --config 3 --steps 5 --warmup 3 > gpurun_out/c4_bench_cfg3.json 2> gpurun_out/c4_bench_cfg3.err; echo "bench cfg3 rc $?"; tail -3 gpurun_out/c4_bench_cfg3.err; head -c 1500 gpurun_out/c4_bench_cfg3.json; echo
timeout 900 python bench.py --config 4 --steps 3 --warmup 3 > gpurun_out/c4_bench_cfg4.json 2> gpurun_out/c4_bench_cfg4.err; echo "bench cfg4 rc $?"; tail -3 gpurun_out/c4_bench_cfg4.err; head -c 1500 gpurun_out/c4_bench_cfg4.json; echo
timeout 1200 python bench.py --config 2 --steps 3 --warmup 3 > gpurun_out/c4_bench_cfg2.json 2> gpurun_out/c4_bench_cfg2.err; echo "bench cfg2 rc $?"; tail -3 gpurun_out/c4_bench_cfg2.err; head -c 1500 gpurun_out/c4_bench_cfg2.json; echo
