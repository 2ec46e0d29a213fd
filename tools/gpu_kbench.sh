#!/bin/bash
# usage: tools/gpu_kbench.sh "<kbench args>" [more arg strings...]; output in gpurun_out/kbench_<n>.txt
mkdir -p gpurun_out
i=0
for a in "$@"; do
  i=$((i+1))
  echo "== kbench $a =="
  timeout 900 bin/kbench $a > gpurun_out/kbench_$i.txt 2>&1; echo "rc $?"
  cat gpurun_out/kbench_$i.txt
done
