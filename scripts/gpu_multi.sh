#!/bin/bash
# N-GPU bench exactly as the driver launches it
set -u
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -k "evaluator or market_shape_matches" > gpurun_out/tests_eval.log 2>&1; tail -5 gpurun_out/tests_eval.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -2 gpurun_out/bench_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -5 gpurun_out/bench_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 1 --warmup 0 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err
python - <<PY
import json
for f in ["gpurun_out/bench_n1.json", "gpurun_out/bench_n$N.json", "gpurun_out/bench_ref_n$N.json"]:
    try:
        d=json.loads(open(f).read().strip().split("\n")[-1])
        print(f, "n_gpus", d["n_gpus"], "value %.4g" % d["value"], "ms/step %.3f" % d["ms_per_step"], "e2e", d["e2e"].get("ms_per_step"), "%.4g" % d["e2e"]["value"], "mAP", d.get("mAP"))
    except Exception as e:
        print(f, "ERR", e)
PY
