#!/bin/bash
# configs[2] / configs[4] at N=1 with the reference's 200 s batches on 4 and 8 streams, and with 1600 s batches
mkdir -p gpurun_out
run() { tag=$1; shift
  timeout 900 python bench.py --no-cpu-baseline --steps 3 --warmup 3 "$@" > gpurun_out/p_$tag.json 2> gpurun_out/p_$tag.err; echo "$tag rc $?"; tail -1 gpurun_out/p_$tag.err | cut -c1-200
  python - $tag <<'PY'
import json,sys
try:
    d=json.loads(open(f'gpurun_out/p_{sys.argv[1]}.json').read().strip().splitlines()[-1])
    print('  ', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['config']['batches'], d['config']['streams'])
except Exception as e: print('   no line', e)
PY
}
run cfg2_s4 --config 2
run cfg2_s8 --config 2 --streams 8
run cfg2_b1600 --config 2 --max-batch-len 1600 --streams 2
run cfg4_s8 --config 4 --streams 8
run cfg4_b1600 --config 4 --max-batch-len 1600 --streams 2
