#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/tests_all.log 2>&1; tail -12 gpurun_out/tests_all.log
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -2 gpurun_out/bench_default.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_default.json"))
print("value %.4g ms %.3f e2e %.4g (%.2f ms) rerank %.2f ms mAP %.9f rr_mAP %.9f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["rerank"]["ms"], d["mAP"], d["rerank"]["mAP"]))
print(d["stages"], d["roofline"]["frac"], d["roofline"]["traffic"], d["clocks"], d["cpu_baseline"]["value"])
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()"
