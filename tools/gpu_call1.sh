#!/bin/bash
# First GPU call: sanitizer on tiny inputs, then the GPU test-suite, smoke, kernel sweep, bench.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,clocks.max.mem --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt; lscpu | head -20 >> gpurun_out/nproc.txt
echo "== tiny kbench first (small inputs before any large launch) =="
timeout 300 bin/kbench ragged:3000:700 --reps 2 --warmup 1 > gpurun_out/kbench_ragged.txt 2>&1; echo "rc $?"
grep -c "PARITY-FAIL\|FAILED" gpurun_out/kbench_ragged.txt; tail -4 gpurun_out/kbench_ragged.txt
timeout 300 bin/kbench stencil:20:20:20 --reps 2 --warmup 1 > gpurun_out/kbench_tiny.txt 2>&1; echo "rc $?"
grep -c "PARITY-FAIL\|FAILED" gpurun_out/kbench_tiny.txt; tail -4 gpurun_out/kbench_tiny.txt
echo "== pytest gpu =="
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc $?"
tail -25 gpurun_out/pytest_gpu.txt
echo "== smoke =="
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; echo "smoke rc $?"; tail -8 gpurun_out/smoke.txt
echo "== kbench c2 =="
timeout 600 bin/kbench c2 --reps 20 > gpurun_out/kbench_c2.txt 2>&1; echo "kbench rc $?"
grep -c PARITY-FAIL gpurun_out/kbench_c2.txt
sort -t'%' -k1 gpurun_out/kbench_c2.txt | head -0
cat gpurun_out/kbench_c2.txt | awk '{print}' | tail -80
echo "== bench =="
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench rc $?"
tail -3 gpurun_out/bench_c2.err; cat gpurun_out/bench_c2.json
