#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/tests_all.log 2>&1
tail -25 gpurun_out/tests_all.log
for prec in 3xtf32 3xfp16; do
  timeout 900 python bench.py --steps 10 --warmup 3 --precision $prec --cpu-queries 0 > gpurun_out/bench_$prec.json 2> gpurun_out/bench_$prec.err
  tail -3 gpurun_out/bench_$prec.err; python - <<PY
import json
d=json.load(open("gpurun_out/bench_$prec.json"))
print("$prec", "value %.3g pairs/s  ms/step %.3f  e2e ms %.2f  mAP %.12f" % (d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["mAP"]))
print("   stages", {k:(round(v,3) if isinstance(v,float) else v) for k,v in d["stages"].items() if k.endswith("_ms")}, "roofline frac %.3f achieved %.1f" % (d["roofline"]["frac"], d["roofline"]["achieved"]), "rank hbm frac %.3f" % d["stages"]["rank_eval_roofline"]["frac"], "rerank", d["rerank"]["ms"], d["rerank"]["mAP"])
PY
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 3 --rerank market --cpu-queries 0 --precision 3xfp16 > gpurun_out/bench_under_ncu.json 2> gpurun_out/ncu.err
python scripts/summarize_launches.py gpurun_out/launches.csv > gpurun_out/launches_summary.txt; head -14 gpurun_out/launches_summary.txt
