#!/bin/bash
# ncu evidence for the headline kernels (run under gpurun, 1 GPU).
mkdir -p gpurun_out
W=${1:-c2}
echo "== plain bench =="
python bench.py --workload $W --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_plain_$W.json 2> gpurun_out/bench_plain_$W.err \
 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_$W.csv \
      python bench.py --workload $W --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_under_ncu_$W.log 2>&1
echo "rc $?"; tail -2 gpurun_out/bench_plain_$W.err; cut -c1-400 gpurun_out/bench_plain_$W.json
echo "== plain kbench profile =="
bin/kbench $W --profile --reps 3 --warmup 1 > gpurun_out/kbench_profile_plain_$W.txt 2>&1 \
 && ncu --set full --clock-control none --import-source on -k regex:"csr_stream_kernel|hll_warp_kernel|csr_vec_kernel" -c 16 \
      -o gpurun_out/prof_$W bin/kbench $W --profile --reps 3 --warmup 1 > gpurun_out/kbench_profile_ncu_$W.log 2>&1
echo "rc $?"; cat gpurun_out/kbench_profile_plain_$W.txt; tail -5 gpurun_out/kbench_profile_ncu_$W.log
ls -la gpurun_out/
