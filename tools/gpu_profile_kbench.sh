#!/bin/bash
# ncu --set full of the headline kernels on one matrix via kbench --profile (1 GPU)
mkdir -p gpurun_out
W=$1; shift
EXTRA="$@"
bin/kbench $W --profile --reps 2 --warmup 1 $EXTRA > gpurun_out/kbench_profile_plain_$W.txt 2>&1 \
 && ncu --set full --clock-control none --import-source on -k regex:"csr_stream_kernel|hll_warp_kernel|csr_vec_kernel|csr_block_row|csr_split|hll_stream" -c 40 \
      -o gpurun_out/prof_$W bin/kbench $W --profile --reps 2 --warmup 1 $EXTRA > gpurun_out/kbench_profile_ncu_$W.log 2>&1
echo "rc $?"; cat gpurun_out/kbench_profile_plain_$W.txt; tail -3 gpurun_out/kbench_profile_ncu_$W.log
