#!/bin/bash
set -u
mkdir -p gpurun_out
for k in k_rank_count k_row_topk k_dist_tc k_jaccard k_build_v0; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/prof_$k \
     python scripts/profile_kernels.py > gpurun_out/ncu_$k.log 2>&1
  tail -1 gpurun_out/ncu_$k.log
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_rrfull.csv \
   python scripts/bench_rerank_multi.py --workload msmt17 --steps 1 > gpurun_out/rrfull_under_ncu.json 2> gpurun_out/ncu_rrfull.err
python scripts/summarize_launches.py gpurun_out/launches_rrfull.csv | head -20
