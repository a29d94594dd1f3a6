#!/bin/bash
set -u
mkdir -p gpurun_out
for k in k_rank_count k_row_topk k_dist_tc k_prep_rows; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/prof_$k \
     python scripts/profile_kernels.py > gpurun_out/ncu_$k.log 2>&1
  tail -2 gpurun_out/ncu_$k.log
done
ls -la gpurun_out/*.ncu-rep
