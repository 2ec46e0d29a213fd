#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
echo "== pytest dist =="
timeout 240 python -m pytest tests/test_gpu_dist.py -x -q -m gpu > gpurun_out/pytest_dist_$N.txt 2>&1; echo "rc $?"; tail -15 gpurun_out/pytest_dist_$N.txt
for mode in nccl push; do
  echo "== bench N=$N mode=$mode =="
  SPMV_B200_EXCHANGE=$mode timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
     bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/bench_N${N}_$mode.json 2> gpurun_out/bench_N${N}_$mode.err; echo "rc $?"
  tail -5 gpurun_out/bench_N${N}_$mode.err; cut -c1-1500 gpurun_out/bench_N${N}_$mode.json
done
echo "== bench N=1 =="
timeout 120 python bench.py --steps 50 --warmup 5 --no-cpu > gpurun_out/bench_N1.json 2>gpurun_out/bench_N1.err; cut -c1-300 gpurun_out/bench_N1.json
