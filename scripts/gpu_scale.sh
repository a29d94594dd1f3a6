#!/bin/bash
# usage: gpu_scale.sh N [sweep]  -- the round-2 multi-GPU measurements on N GPUs of one box (outputs under gpurun_out/)
N=${1:-8}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
MPREID_CHECK_BACKEND=nccl timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29501 scripts/sharded_eval_check.py > gpurun_out/r2_check_n$N.log 2>&1; tail -1 gpurun_out/r2_check_n$N.log
run 29502 bench.py --gpus $N --steps 10 --cpu-queries 0 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
run 29503 bench.py --gpus $N --steps 10 --scaling strong --rerank none --cpu-queries 0 > gpurun_out/r2_bench_n${N}_strong.json 2>> gpurun_out/r2_bench_n$N.err
run 29504 bench.py --gpus $N --workload c5 --steps 3 > gpurun_out/r2_bench_n${N}_c5.json 2>> gpurun_out/r2_bench_n$N.err
if [ "${2:-}" = "sweep" ]; then
  for s in 4096 8192; do
    MPREID_SHARD_SUB_ROWS=$s run 29505 bench.py --gpus $N --steps 3 --rerank none --cpu-queries 0 > gpurun_out/r2_bench_n${N}_sub$s.json 2>> gpurun_out/r2_bench_n$N.err
  done
  MPREID_SHARD_EXCHANGE=nccl run 29506 bench.py --gpus $N --steps 3 --rerank none --cpu-queries 0 > gpurun_out/r2_bench_n${N}_ncclx.json 2>> gpurun_out/r2_bench_n$N.err
fi
grep -i "error\|Traceback" gpurun_out/r2_bench_n$N.err | head -5
