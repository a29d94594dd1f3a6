#!/bin/bash
# usage: gpu_scale.sh N   -- the round-2 multi-GPU measurements on N GPUs of one box (outputs under gpurun_out/)
N=${1:-8}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
MPREID_CHECK_BACKEND=nccl run 29501 scripts/sharded_eval_check.py > gpurun_out/r2_check_n$N.log 2>&1; tail -2 gpurun_out/r2_check_n$N.log
run 29502 bench.py --gpus $N --steps 10 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
run 29503 bench.py --gpus $N --steps 10 --scaling strong --rerank none > gpurun_out/r2_bench_n${N}_strong.json 2>> gpurun_out/r2_bench_n$N.err
run 29504 bench.py --gpus $N --workload c5 --steps 3 > gpurun_out/r2_bench_n${N}_c5.json 2>> gpurun_out/r2_bench_n$N.err
MPREID_SHARD_EXCHANGE=nccl run 29505 bench.py --gpus $N --steps 5 --rerank none > gpurun_out/r2_bench_n${N}_ncclx.json 2>> gpurun_out/r2_bench_n$N.err
tail -c 1500 gpurun_out/r2_bench_n$N.err
