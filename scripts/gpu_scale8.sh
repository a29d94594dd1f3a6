#!/bin/bash
set -u
mkdir -p gpurun_out
run() { n=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) "$@"; }
for n in 1 2 4 8; do
  if [ $n -eq 1 ]; then timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 --rerank none > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  else run $n bench.py --gpus $n --steps 10 --warmup 3 --rerank none > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err; fi
  tail -1 gpurun_out/scale_n$n.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench n=%d value %.4g ms %.3f e2e %.4g (%.2f ms) h2d %.1f GB/s mAP %.9f' % (d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['h2d_gbs_measured'], d['mAP']))" || tail -3 gpurun_out/scale_n$n.err
done
for n in 4 8; do
  run $n scripts/bench_rerank_multi.py --workload msmt17 --steps 3 > gpurun_out/rr_msmt17_n$n.json 2> gpurun_out/rr_msmt17_n$n.err; grep '^{' gpurun_out/rr_msmt17_n$n.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('rerank n=%d %.2f ms mAP %.12f' % (d['n_gpus'], d['value'], d['mAP']))" || tail -3 gpurun_out/rr_msmt17_n$n.err
done
run 8 scripts/bench_retrieval.py --scale 1.0 --steps 2 > gpurun_out/retrieval_full_n8.json 2> gpurun_out/retrieval_full_n8.err; grep '^{' gpurun_out/retrieval_full_n8.json | cut -c1-400 || tail -3 gpurun_out/retrieval_full_n8.err
