#!/bin/bash
# bench (ours + reference arm) and the ncu launch list of the same command
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -k "rerank or rank_eval_bit" > gpurun_out/tests_fix.log 2>&1
tail -15 gpurun_out/tests_fix.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
tail -3 gpurun_out/bench_ours.err; cat gpurun_out/bench_ours.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
cat gpurun_out/bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 3 --rerank market --cpu-queries 0 > gpurun_out/bench_under_ncu.json 2> gpurun_out/ncu.err
tail -3 gpurun_out/ncu.err
