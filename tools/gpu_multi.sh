#!/bin/bash
# usage: tools/gpu_multi.sh N [extra bench args]   (run under gpurun --gpus N): the default bench at N GPUs with the push and
# the pull transport and with ids-only gather
N=$1; shift
mkdir -p gpurun_out
run() {  # tag, port, extra args
  tag=$1; port=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/m_${tag}_n$N.json 2> gpurun_out/m_${tag}_n$N.err
  echo "bench $tag N=$N rc $?"; grep -v "OMP_NUM_THREADS\|^\*\*\*\|NCCL version" gpurun_out/m_${tag}_n$N.err | tail -4 | cut -c1-300
  python - $tag $N <<'PY'
import json,sys
tag,n=sys.argv[1:3]
try:
    d=json.loads(open(f'gpurun_out/m_{tag}_n{n}.json').read().strip().splitlines()[-1])
    print('  ', d['value'], d['ms_per_step'], 'eager', d['roofline'].get('eager_ms_per_step'), 'e2e', d['e2e']['value'], d.get('gather_check'), (d.get('rank0_ingress') or {}).get('achieved_gbs'))
except Exception as e: print('   no line', e)
PY
}
run push 29511 "$@"
run pull 29512 --gather-transport peer "$@"
run push_ids 29513 --gather ids "$@"
