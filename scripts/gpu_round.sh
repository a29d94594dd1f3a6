#!/bin/bash
# Runs on the B200 box under gpurun: staged GPU tests (logic first with the SIMT distance path, then
# the tcgen05 kernels), each stage under its own timeout so a hung kernel cannot eat the lease.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "=== stage A: logic with SIMT distances" | tee gpurun_out/stageA.log
MPREID_PRECISION=simt timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider \
  -k "not 3xtf32 and not bf16 and not tcgen05 and not market_shape" >> gpurun_out/stageA.log 2>&1
echo "exit $?" >> gpurun_out/stageA.log
tail -25 gpurun_out/stageA.log
echo "=== stage B: tcgen05 kernels" | tee gpurun_out/stageB.log
timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider \
  -k "(3xtf32 or bf16 or tcgen05) and not market_shape" >> gpurun_out/stageB.log 2>&1
echo "exit $?" >> gpurun_out/stageB.log
tail -25 gpurun_out/stageB.log
echo "=== stage C: full shapes, default precision" | tee gpurun_out/stageC.log
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -k "market_shape or evaluator" >> gpurun_out/stageC.log 2>&1
echo "exit $?" >> gpurun_out/stageC.log
tail -25 gpurun_out/stageC.log
