#!/bin/bash
# DRAM traffic and duration of the MSMT17-shaped distance GEMM for band sizes / TMA L2 hints
set -u
mkdir -p gpurun_out
for cfg in "16 0" "16 1" "16 2" "16 3" "8 0" "32 0" "46 0" "32 3"; do
  set -- $cfg
  echo "band=$1 hint=$2"
  MPREID_GEMM_BAND=$1 MPREID_GEMM_HINT=$2 timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:k_dist_tc -s 1 -c 1 python scripts/profile_kernels.py 2>&1 | grep -E "dram__|gpu__time" | awk '{print "   ", $1, $2, $3}'
  MPREID_GEMM_BAND=$1 MPREID_GEMM_HINT=$2 timeout 200 python scripts/pair_check.py gpurun_out/sweep.json big 2>&1 | tail -1
done
