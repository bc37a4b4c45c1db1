#!/bin/bash
# One gpurun call for the second Fbank kernel: bench, parity tests, timeline of CTA 0, timing-experiment builds.
# usage: tools/gpu_fbank2.sh TAG [variants...]
TAG=${1:-x}; shift
P=stac_speech_translation_b200
mkdir -p gpurun_out
timeout 40 python tools/bench_fbank.py > gpurun_out/${TAG}_fbank_bench.log 2>&1; rc=$?
echo "bench rc $rc"; tail -6 gpurun_out/${TAG}_fbank_bench.log | cut -c1-200
[ $rc = 0 ] || exit 1
timeout 90 python -m pytest tests/test_gpu_fp32_kernels.py -x -q -k "fbank" > gpurun_out/${TAG}_fbank_tests.log 2>&1
echo "tests rc $?"; tail -3 gpurun_out/${TAG}_fbank_tests.log | cut -c1-200
for pair in 1 0; do timeout 60 python tools/trace_fbank2.py $P/libstac_b200_fbtrace.so $pair; done > gpurun_out/${TAG}_trace.log 2>&1
echo "trace rc $?"; cat gpurun_out/${TAG}_trace.log
for v in "$@"; do
  echo "== $v"; timeout 60 python tools/bench_fbank.py $P/libstac_b200_fb2_$v.so v2_single v2_pair 2>&1 | grep -v "rel-L2"
done > gpurun_out/${TAG}_variants.log 2>&1
cat gpurun_out/${TAG}_variants.log
