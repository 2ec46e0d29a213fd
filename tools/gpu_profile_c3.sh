#!/bin/bash
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,lts__t_bytes.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size"
timeout 300 bin/kbench c3 --profile --reps 2 --warmup 1 > gpurun_out/kbench_profile_plain_c3.txt 2>&1 \
 && timeout 600 ncu --metrics $M --clock-control none -k regex:"csr_stream_kernel|hll_warp_kernel|csr_vec_kernel|hll_stream" -c 15 --csv --log-file gpurun_out/prof_c3_metrics.csv \
      bin/kbench c3 --profile --reps 2 --warmup 1 > gpurun_out/kbench_profile_ncu_c3.log 2>&1
echo "rc $?"; grep "CSR\|HLL" gpurun_out/kbench_profile_plain_c3.txt | cut -c1-180; tail -3 gpurun_out/kbench_profile_ncu_c3.log | cut -c1-200
