"""torchrun check of distributed.sharded_evaluator: ragged query / gallery shards fed per rank from host memory must
give the cmc / mAP of a one-GPU evaluation of all queries, bit for bit.  Prints 'SHARDED_EVAL_OK' on rank 0."""
import contextlib, io, os, sys
import numpy as np, torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mp_reid_b200 import metrics, distributed as MD

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
os.environ["MPREID_DEVICE"] = f"cuda:{local}"
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rng = np.random.RandomState(21)
Q, G, D = 1501, 9003, 320
x = torch.from_numpy(rng.randn(Q + G, D).astype(np.float32))
pid = rng.randint(0, 120, Q + G); cam = rng.randint(0, 6, Q + G)
q_lo, q_hi = MD.shard_bounds(Q, world, rank)
g_lo, g_hi = MD.aligned_shard_bounds(G, world, rank)
for junk in ("none", "pid_cam"):
    ev = MD.sharded_evaluator(q_hi - q_lo, junk=junk); ev.reset()
    for s in range(q_lo, q_hi, 400):
        e = min(q_hi, s + 400); ev.update((x[s:e].pin_memory(), pid[s:e], cam[s:e]))
    for s in range(Q + g_lo, Q + g_hi, 1000):
        e = min(Q + g_hi, s + 1000); ev.update((x[s:e], pid[s:e], cam[s:e]))
    with contextlib.redirect_stdout(io.StringIO()):
        cmc, mAP, dmat, *_, qf, gf = ev.compute()
    ref = metrics.R1_mAP_eval(Q, junk=junk); ref.reset(); ref.update((x, pid, cam))
    with contextlib.redirect_stdout(io.StringIO()):
        cmc0, mAP0, d0, *_, qf0, gf0 = ref.compute()
    assert np.array_equal(cmc, cmc0) and mAP == mAP0, (rank, junk, mAP, mAP0)
    assert np.array_equal(np.asarray(dmat), np.asarray(d0)[q_lo:q_hi]) and torch.equal(gf, gf0) and torch.equal(qf, qf0[q_lo:q_hi])
dist.barrier()
if rank == 0:
    print("SHARDED_EVAL_OK", world)
dist.destroy_process_group()
