#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/tests_all.log 2>&1
tail -12 gpurun_out/tests_all.log
show() { python - "$1" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print(d["config"]["precision"], "value %.4g pairs/s  ms/step %.3f  e2e ms %.2f (%.3g pairs/s) mAP %.12f launches %d" % (d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"]["value"], d["mAP"], d["gpu_launches"]))
print("   stages", {k:(round(v,3) if isinstance(v,float) else v) for k,v in d["stages"].items() if k.endswith("_ms")}, "roofline frac %.3f achieved %.1f peak %.1f" % (d["roofline"]["frac"], d["roofline"]["achieved"], d["roofline"]["peak"]), "rank hbm frac %.3f" % d["stages"]["rank_eval_roofline"]["frac"])
print("   rerank", d["rerank"], "clocks", d["clocks"], "cpu", d["cpu_baseline"])
PY
}
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
tail -3 gpurun_out/bench_ours.err; show gpurun_out/bench_ours.json
timeout 900 python bench.py --steps 3 --warmup 3 --rerank full --cpu-queries 0 > gpurun_out/bench_rrfull.json 2> gpurun_out/bench_rrfull.err
tail -3 gpurun_out/bench_rrfull.err; show gpurun_out/bench_rrfull.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 3 --rerank market --cpu-queries 0 > gpurun_out/bench_under_ncu.json 2> gpurun_out/ncu.err
python scripts/summarize_launches.py gpurun_out/launches.csv > gpurun_out/launches_summary.txt; head -16 gpurun_out/launches_summary.txt
