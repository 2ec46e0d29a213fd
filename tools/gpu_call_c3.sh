#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu =="; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc $?"; tail -4 gpurun_out/pytest_gpu.txt
echo "== c3 persist on =="; timeout 600 bin/kbench c3 --profile --reps 10 > gpurun_out/kbench_c3_persist1.txt 2>&1; echo "rc $?"; grep "CSR\|HLL" gpurun_out/kbench_c3_persist1.txt | cut -c1-185
echo "== c3 persist off =="; timeout 600 bin/kbench c3 --profile --reps 10 --knob x_persist=0 > gpurun_out/kbench_c3_persist0.txt 2>&1; echo "rc $?"; grep "CSR\|HLL" gpurun_out/kbench_c3_persist0.txt | cut -c1-185
echo "== c4 persist on/off =="; timeout 600 bin/kbench c4 --profile --only csr --reps 10 > gpurun_out/kbench_c4_persist1.txt 2>&1; grep "CSR" gpurun_out/kbench_c4_persist1.txt | cut -c1-185
timeout 600 bin/kbench c4 --profile --only csr --reps 10 --knob x_persist=0 > gpurun_out/kbench_c4_persist0.txt 2>&1; grep "CSR" gpurun_out/kbench_c4_persist0.txt | cut -c1-185
echo "== c2 headline =="; timeout 300 bin/kbench c2 --profile --reps 20 > gpurun_out/kbench_c2_headline.txt 2>&1; grep "CSR\|HLL" gpurun_out/kbench_c2_headline.txt | cut -c1-185
echo "== c5 N=1 (512^3, 46 GB) =="
SPMV_B200_GRAPH=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 1 --workload c5 --steps 10 --warmup 4 > gpurun_out/bench_c5_N1.json 2> gpurun_out/bench_c5_N1.err; echo "rc $?"; tail -3 gpurun_out/bench_c5_N1.err | cut -c1-300; cut -c1-900 gpurun_out/bench_c5_N1.json
