#!/bin/bash
# 8-GPU box: driver-style bench at N=8 (and N=4), single-process multi-GPU probe
set -u
mkdir -p gpurun_out
run() { n=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700+n)) "$@"; }
for n in 8 4; do
  run $n bench.py --gpus $n > gpurun_out/bench_n${n}_default.json 2> gpurun_out/bench_n${n}_default.err
  tail -1 gpurun_out/bench_n${n}_default.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench n=%d value %.4g ms %.3f e2e %.4g (%.2f ms) rerank %.2f ms mAP %.9f rr_mAP %.9f' % (d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['rerank']['ms'], d['mAP'], d['rerank']['mAP']))" || tail -3 gpurun_out/bench_n${n}_default.err
done
timeout 300 python scripts/multidev_probe.py 2>&1 | tail -4
MPREID_DEVICES=0,1,2,3 timeout 300 python - <<'PY' 2>&1 | tail -2
import os, sys
sys.argv = ["x"]
exec(open("scripts/multidev_probe.py").read().replace('for spec in ["", "all"]:', 'for spec in ["0,1,2,3"]:'))
PY
