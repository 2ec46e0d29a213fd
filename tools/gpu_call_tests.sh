#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu =="
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc $?"
tail -8 gpurun_out/pytest_gpu.txt
shift 0
i=0
for a in "$@"; do
  i=$((i+1))
  echo "== kbench $a =="
  timeout 900 bin/kbench $a > gpurun_out/kbench_$i.txt 2>&1; echo "rc $?"
  grep -v "thread_row  \|block_row\|warp_row" gpurun_out/kbench_$i.txt
done
