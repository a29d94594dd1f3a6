#!/bin/bash
# Round-2 evidence under gpurun_out/ (copied to profiles/r02/ by scripts/collect_profiles_r02.py): launch lists and
# `ncu --set full` captures of the dominant kernels.  One GPU.
set -u
mkdir -p gpurun_out
# launch lists (per-kernel shares; cold-cache, serialised -- shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench.csv \
   python bench.py --steps 2 --warmup 3 --rerank market --cpu-queries 0 > gpurun_out/r02_bench_under_ncu.json 2> gpurun_out/r02_ncu.err
python scripts/summarize_launches.py gpurun_out/r02_launches_bench.csv > gpurun_out/r02_launches_bench.summary.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_rerank_msmt17.csv \
   python scripts/rerank_stages.py msmt17 1 > /dev/null 2>> gpurun_out/r02_ncu.err
python scripts/summarize_launches.py gpurun_out/r02_launches_rerank_msmt17.csv > gpurun_out/r02_launches_rerank_msmt17.summary.txt
# --set full captures.  profile_kernels.py: 2 x (prep, rect GEMM, rank) at MSMT17 shape, then 2 x Market re-ranking
for k in k_rank_count k_prep_rows_warp k_blend_default k_jaccard_bucket k_cand_topk k_build_v0 k_query_expand_warp k_row_kth_bound; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/r02_prof_$k \
     python scripts/profile_kernels.py > gpurun_out/r02_ncu_$k.log 2>&1
done
# rect GEMM (CTA pairs) = 2nd k_dist_tc launch of the driver; fused symmetric GEMM at Market shape = the last one
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dist_tc -s 1 -c 1 -f -o gpurun_out/r02_prof_k_dist_tc_rect \
   python scripts/profile_kernels.py > gpurun_out/r02_ncu_k_dist_tc_rect.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dist_tc -s 5 -c 1 -f -o gpurun_out/r02_prof_k_dist_tc_fused_market \
   python scripts/profile_kernels.py > gpurun_out/r02_ncu_k_dist_tc_fused.log 2>&1
# the same kernels at MSMT17 shape inside a re-ranking pass (fused GEMM = 2nd k_dist_tc of the pass: sample GEMM first)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_dist_tc -s 3 -c 1 -f -o gpurun_out/r02_prof_k_dist_tc_fused_msmt17 \
   python scripts/rerank_stages.py msmt17 1 > gpurun_out/r02_ncu_fused_msmt17.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_jaccard_bucket -s 1 -c 1 -f -o gpurun_out/r02_prof_k_jaccard_bucket_msmt17 \
   python scripts/rerank_stages.py msmt17 1 >> gpurun_out/r02_ncu_fused_msmt17.log 2>&1
# summarise on the box (the reports with imported source are ~20 MB each; only gpurun_out/ <= 64 MiB travels back)
python scripts/collect_profiles_r02.py gpurun_out/profiles_r02 > gpurun_out/r02_traffic.log 2>&1
find gpurun_out -name "r02_prof_*.ncu-rep" -size +6M -delete
ls -la gpurun_out/r02_prof_*.ncu-rep | wc -l
