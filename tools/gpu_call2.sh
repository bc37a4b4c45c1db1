#!/bin/bash
# round 2, call 2: whole GPU suite (un-gated), attention v2 (double-buffered S) parity + timing + trace, bench with in-run parity
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/c2_gpu_tests.log 2>&1; echo "gpu tests rc $?"; tail -5 gpurun_out/c2_gpu_tests.log
timeout 120 python tools/bench_mha.py stac_speech_translation_b200/libstac_b200.so > gpurun_out/c2_mha_bench.log 2>&1; echo "bench_mha rc $?"; cat gpurun_out/c2_mha_bench.log
python -m stac_speech_translation_b200.build --variant mha2trace -- -DMHA2_TRACE > gpurun_out/c2_trace_build.log 2>&1
timeout 120 python tools/trace_mha2.py stac_speech_translation_b200/libstac_b200_mha2trace.so > gpurun_out/c2_mha2_trace.log 2>&1; echo "trace rc $?"; tail -45 gpurun_out/c2_mha2_trace.log
for v in "kv5:-DMHA2_KV_STAGES=5" "kv3:-DMHA2_KV_STAGES=3"; do
  name=${v%%:*}; flag=${v#*:}
  python -m stac_speech_translation_b200.build --variant $name -- $flag > gpurun_out/c2_build_$name.log 2>&1
  timeout 120 python tools/bench_mha.py stac_speech_translation_b200/libstac_b200_$name.so > gpurun_out/c2_mha_bench_$name.log 2>&1
  echo "variant $name rc $?"; cat gpurun_out/c2_mha_bench_$name.log
done
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/c2_bench.json 2> gpurun_out/c2_bench.err; echo "bench rc $?"; tail -c 2500 gpurun_out/c2_bench.json
STAC_MHA_V2=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c2_bench_v2.json 2> gpurun_out/c2_bench_v2.err; echo "bench v2 rc $?"; head -c 600 gpurun_out/c2_bench_v2.json
