#!/bin/bash
# compute-sanitizer over the small-shape GPU tests (memcheck) and the hand-synchronised kernels (racecheck)
set -u
mkdir -p gpurun_out
SEL="not _shape and not retrieval and not many_positives and not medium"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -q -x --tb=line -p no:cacheprovider -k "$SEL" > gpurun_out/memcheck.log 2>&1
echo "memcheck exit $?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|=========     at" gpurun_out/memcheck.log | head -20
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests -m gpu -q -x --tb=line -p no:cacheprovider -k "(row_topk or rank_eval_bit or rerank_sparse or rank_eval_many_random) and not _shape" > gpurun_out/racecheck.log 2>&1
echo "racecheck exit $?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard|Race reported" gpurun_out/racecheck.log | sort | uniq -c | head -20
