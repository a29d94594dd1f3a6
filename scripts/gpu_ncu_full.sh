#!/bin/bash
set -u
mkdir -p gpurun_out
for k in k_rank_count k_row_topk k_dist_tc k_jaccard k_prep_rows; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/prof_$k \
     python scripts/profile_kernels.py > gpurun_out/ncu_$k.log 2>&1
  tail -1 gpurun_out/ncu_$k.log
done
# the symmetric all-pairs launch is the 3rd k_dist_tc launch of the driver (2 MSMT17 passes, then the Market all-pairs)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dist_tc -s 2 -c 1 -f -o gpurun_out/prof_k_dist_tc_sym \
     python scripts/profile_kernels.py > gpurun_out/ncu_k_dist_tc_sym.log 2>&1; tail -1 gpurun_out/ncu_k_dist_tc_sym.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 3 --rerank market --cpu-queries 0 > gpurun_out/bench_under_ncu.json 2> gpurun_out/ncu.err
python scripts/summarize_launches.py gpurun_out/launches.csv > gpurun_out/launches_summary.txt; head -14 gpurun_out/launches_summary.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_rrfull.csv python scripts/bench_rerank_multi.py --workload msmt17 --steps 1 > gpurun_out/rrfull_under_ncu.json 2> gpurun_out/ncu_rrfull.err
python scripts/summarize_launches.py gpurun_out/launches_rrfull.csv > gpurun_out/launches_rrfull_summary.txt; head -10 gpurun_out/launches_rrfull_summary.txt
