#!/bin/bash
set -u
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -k "rerank or re_ranking" > gpurun_out/tests_rr.log 2>&1; tail -8 gpurun_out/tests_rr.log
for w in market msmt17; do
timeout 900 python scripts/bench_rerank_multi.py --workload $w > gpurun_out/rr_${w}_n1.json 2> gpurun_out/rr_${w}_n1.err; tail -2 gpurun_out/rr_${w}_n1.err; cat gpurun_out/rr_${w}_n1.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 scripts/bench_rerank_multi.py --workload $w > gpurun_out/rr_${w}_n$N.json 2> gpurun_out/rr_${w}_n$N.err; tail -3 gpurun_out/rr_${w}_n$N.err | cut -c1-300; cat gpurun_out/rr_${w}_n$N.json
done
