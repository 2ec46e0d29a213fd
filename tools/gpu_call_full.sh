#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu =="
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc $?"
tail -6 gpurun_out/pytest_gpu.txt
echo "== smoke =="
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; echo "smoke rc $?"; tail -6 gpurun_out/smoke.txt
echo "== bench c2 =="
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "rc $?"; tail -2 gpurun_out/bench_default.err; cut -c1-700 gpurun_out/bench_default.json
timeout 600 python bench.py --format hll --no-cpu > gpurun_out/bench_hll.json 2> gpurun_out/bench_hll.err; echo "rc $?"; cut -c1-300 gpurun_out/bench_hll.json
echo "== bench reference arm =="
timeout 600 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "rc $?"; cut -c1-900 gpurun_out/bench_reference.json
echo "== kbench c2 full =="
timeout 600 bin/kbench c2 > gpurun_out/kbench_c2_sweep.txt 2>&1; echo "rc $?"; grep -c PARITY-FAIL gpurun_out/kbench_c2_sweep.txt
echo "== kbench c3 =="
timeout 900 bin/kbench c3 --quick --reps 10 > gpurun_out/kbench_c3.txt 2>&1; echo "rc $?"; grep -v "thread_row  \|block_row\|cfg=1[0-9]\|cfg=2" gpurun_out/kbench_c3.txt | cut -c1-185
echo "== kbench c1 flush =="
timeout 300 bin/kbench c1 --quick --flush > gpurun_out/kbench_c1_flush.txt 2>&1; echo "rc $?"; grep "auto\|vec=" gpurun_out/kbench_c1_flush.txt | cut -c1-185
