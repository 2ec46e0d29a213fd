#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
run() { # name, env..., args
  name=$1; shift
  echo "== $name =="
  env "$@" timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 \
     bench.py --gpus $N --steps 50 --warmup 6 $EXTRA > gpurun_out/bench_N${N}_$name.json 2> gpurun_out/bench_N${N}_$name.err; echo "rc $?"
  tail -2 gpurun_out/bench_N${N}_$name.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_N${N}_$name.json').read().strip().splitlines()[-1])
    print('$name', 'ms/step', round(d['ms_per_step'],5), 'GFLOP/s', round(d['value'],1), 'graph', d['config']['cuda_graph'], d['config'].get('graph_error'), 'kernel frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value'],1), 'launches', d['gpu_launches'])
except Exception as e:
    print('$name: no json', e)
PY
}
for spec in ${2:-weak_push weak_nccl c5_push}; do
  case $spec in
    weak_push) EXTRA="" run weak_push SPMV_B200_EXCHANGE=push;;
    weak_nccl) EXTRA="" run weak_nccl SPMV_B200_EXCHANGE=nccl;;
    c5_push) EXTRA="--workload c5" run c5_push SPMV_B200_EXCHANGE=push;;
    c5_nccl) EXTRA="--workload c5" run c5_nccl SPMV_B200_EXCHANGE=nccl;;
  esac
done
