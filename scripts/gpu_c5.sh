#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -k "retrieval or evaluator" > gpurun_out/tests_ret.log 2>&1; tail -8 gpurun_out/tests_ret.log
timeout 900 python bench.py --steps 10 --warmup 3 --cpu-queries 0 --rerank none > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; tail -2 gpurun_out/bench_ours.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_ours.json")); print("C4 value %.4g ms %.3f e2e ms %.2f" % (d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"]))
PY
timeout 900 python scripts/bench_retrieval.py --scale 0.25 --steps 2 > gpurun_out/retrieval_q25.json 2> gpurun_out/retrieval_q25.err; tail -2 gpurun_out/retrieval_q25.err; cat gpurun_out/retrieval_q25.json
timeout 1200 python scripts/bench_retrieval.py --scale 1.0 --steps 2 > gpurun_out/retrieval_full.json 2> gpurun_out/retrieval_full.err; tail -2 gpurun_out/retrieval_full.err; cat gpurun_out/retrieval_full.json
