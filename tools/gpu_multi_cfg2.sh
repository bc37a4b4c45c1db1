#!/bin/bash
N=$1
mkdir -p gpurun_out
run() { tag=$1; port=$2; shift 2
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --no-cpu-baseline --config 2 --steps 3 --warmup 3 "$@" > gpurun_out/m_${tag}_n$N.json 2> gpurun_out/m_${tag}_n$N.err
  echo "bench $tag N=$N rc $?"; grep -v "OMP_NUM_THREADS\|^\*\*\*\|NCCL version" gpurun_out/m_${tag}_n$N.err | tail -4 | cut -c1-300
  python - $tag $N <<'PY'
import json,sys
tag,n=sys.argv[1:3]
try:
    d=json.loads(open(f'gpurun_out/m_{tag}_n{n}.json').read().strip().splitlines()[-1])
    print('  ', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d.get('gather_check'), d['config']['multi_gpu'][:90])
except Exception as e: print('   no line', e)
PY
}
run cfg2push 29521
if [ "$N" -le 2 ]; then run cfg2nccl 29522 --gather-transport nccl; fi
