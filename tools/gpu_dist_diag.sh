#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
for mode in ${2:-push}; do
  SPMV_B200_DIAG=1 SPMV_B200_EXCHANGE=$mode timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
     bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/diag_N${N}_$mode.json 2> gpurun_out/diag_N${N}_$mode.err; echo "rc $?"
  tail -3 gpurun_out/diag_N${N}_$mode.err | cut -c1-300
  python -c "
import json,sys
d=json.loads(open('gpurun_out/diag_N${N}_$mode.json').read().strip().splitlines()[-1])
print('$mode', d['ms_per_step'], d['value'], json.dumps(d['config']['diag'], indent=1))"
done
