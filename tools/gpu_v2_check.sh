#!/bin/bash
# First thing to run on a B200 next round (one gpurun call, one GPU):
#   gpurun --timeout 1500 -- bash tools/gpu_v2_check.sh          (about 15-20 minutes of box time)
# Runs the parity tests of the kernels that were written after the round-1 GPU budget was spent (STAC_EXPERIMENTAL=1),
# each under its own time limit (every mbarrier wait in them traps after ~2 s, so a protocol bug is a launch failure, not
# a hang), then times attention v2 against the default kernel and the whole path with STAC_MHA_V2=1.
mkdir -p gpurun_out
export STAC_EXPERIMENTAL=1
# tensor-memory conventions attention v2 relies on (exact integer test of the TS-form MMA under three layouts of P)
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 --expt-relaxed-constexpr -o /tmp/probe_ts_mma tools/probe_ts_mma.cu > gpurun_out/v2_probe_build.log 2>&1 && timeout 60 /tmp/probe_ts_mma > gpurun_out/v2_probe_ts_mma.log 2>&1; echo "probe_ts_mma rc $?"; cat gpurun_out/v2_probe_ts_mma.log
timeout 300 python -m pytest tests/test_gpu_tc_attention.py -q -x -m gpu -k 'v2 and not many_short' > gpurun_out/v2_mha_tests.log 2>&1
echo "attention v2 tests rc $?"; tail -5 gpurun_out/v2_mha_tests.log
timeout 300 python -m pytest tests/test_gpu_turns.py tests/test_gpu_wav_ingest.py tests/test_gpu_xcustom_ops.py tests/test_gpu_ytrain_norm.py -q -m gpu > gpurun_out/v2_turn_tests.log 2>&1
echo "turn-detection tests rc $?"; tail -5 gpurun_out/v2_turn_tests.log
timeout 600 python -m pytest tests/test_gpu_decoder.py -q -x -m gpu > gpurun_out/v2_decoder_tests.log 2>&1
echo "decoder tests rc $?"; tail -8 gpurun_out/v2_decoder_tests.log
# the per-buffer o_staged fix of the default attention kernel: same tests on the variant build, then time it
python -m stac_speech_translation_b200.build --variant ostaged -- -DMHA_OSTAGED_PER_BUFFER > gpurun_out/v2_variant_build.log 2>&1
STAC_B200_LIB=$PWD/stac_speech_translation_b200/libstac_b200_ostaged.so timeout 600 python -m pytest tests/test_gpu_tc_attention.py tests/test_gpu_bf16_path.py -q -x -m gpu -k 'not many_short and not v2' > gpurun_out/v2_ostaged_tests.log 2>&1
echo "o_staged-per-buffer variant tests rc $?"; tail -3 gpurun_out/v2_ostaged_tests.log
# the stress shape of that finding, each build in its own process (a trap poisons the CUDA context): variant first
STAC_B200_LIB=$PWD/stac_speech_translation_b200/libstac_b200_ostaged.so timeout 300 python -m pytest tests/test_gpu_tc_attention.py -q -m gpu -k many_short > gpurun_out/v2_stress_ostaged.log 2>&1
echo "stress (per-buffer variant) rc $?"; tail -3 gpurun_out/v2_stress_ostaged.log
timeout 300 python -m pytest tests/test_gpu_tc_attention.py -q -m gpu -k many_short > gpurun_out/v2_stress_default.log 2>&1
echo "stress (default build) rc $?"; tail -3 gpurun_out/v2_stress_default.log
timeout 120 python tools/bench_mha.py stac_speech_translation_b200/libstac_b200.so stac_speech_translation_b200/libstac_b200_ostaged.so > gpurun_out/v2_mha_bench.log 2>&1
echo "bench_mha rc $?"; cat gpurun_out/v2_mha_bench.log
STAC_MHA_V2=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/v2_bench.json 2> gpurun_out/v2_bench.err
echo "bench (STAC_MHA_V2=1) rc $?"; tail -c 1200 gpurun_out/v2_bench.json
timeout 300 python tools/bench_decoder.py > gpurun_out/v2_decoder_bench.log 2>&1; echo "bench_decoder rc $?"; cat gpurun_out/v2_decoder_bench.log
python -m stac_speech_translation_b200.build --variant mha2trace -- -DMHA2_TRACE > gpurun_out/v2_trace_build.log 2>&1
timeout 120 python tools/trace_mha2.py stac_speech_translation_b200/libstac_b200_mha2trace.so > gpurun_out/v2_mha2_trace.log 2>&1
echo "trace_mha2 rc $?"; tail -12 gpurun_out/v2_mha2_trace.log
# attention v2 knobs: share of exponentials on the FMA pipe, depth of the K/V ring, softmax groups taking turns (3 is the most that fits; each variant: its own build, timed alone)
for v in "poly2:-DMHA2_POLY=2" "poly4:-DMHA2_POLY=4" "kv2:-DMHA2_KV_STAGES=2" "seq:-DMHA2_SEQUENCE"; do
  name=${v%%:*}; flag=${v#*:}
  python -m stac_speech_translation_b200.build --variant $name -- $flag > gpurun_out/v2_build_$name.log 2>&1
  timeout 120 python tools/bench_mha.py stac_speech_translation_b200/libstac_b200_$name.so > gpurun_out/v2_mha_bench_$name.log 2>&1
  echo "variant $name rc $?"; cat gpurun_out/v2_mha_bench_$name.log
done
