#!/bin/bash
mkdir -p gpurun_out
for g in 32 128; do
echo "== c3 l2_fetch_granularity=$g =="; timeout 400 bin/kbench c3 --profile --reps 10 --knob l2_fetch_granularity=$g > gpurun_out/kbench_c3_gran$g.txt 2>&1; echo "rc $?"; grep "CSR\|HLL\|ERROR" gpurun_out/kbench_c3_gran$g.txt | cut -c1-180
done
echo "== c2 gran 32 =="; timeout 200 bin/kbench c2 --profile --reps 20 --knob l2_fetch_granularity=32 > gpurun_out/kbench_c2_gran32.txt 2>&1; grep "CSR\|HLL" gpurun_out/kbench_c2_gran32.txt | cut -c1-180
