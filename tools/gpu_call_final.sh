#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu =="; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc $?"; tail -4 gpurun_out/pytest_gpu.txt
echo "== c4 headline =="; timeout 400 bin/kbench c4 --profile --only csr --reps 10 > gpurun_out/kbench_c4_headline.txt 2>&1; grep "CSR" gpurun_out/kbench_c4_headline.txt | cut -c1-185
echo "== bench c4 =="; timeout 600 python bench.py --workload c4 --kernel 2 --steps 20 --no-cpu > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "rc $?"; tail -2 gpurun_out/bench_c4.err | cut -c1-300; cut -c1-400 gpurun_out/bench_c4.json
echo "== bench c3 hll =="; timeout 900 python bench.py --workload c3 --format hll --wpb 16 --steps 20 --no-cpu > gpurun_out/bench_c3_hll.json 2> gpurun_out/bench_c3_hll.err; echo "rc $?"; tail -2 gpurun_out/bench_c3_hll.err | cut -c1-300; cut -c1-400 gpurun_out/bench_c3_hll.json
echo "== bench c2 =="; timeout 300 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "rc $?"; cut -c1-300 gpurun_out/bench_default.json
