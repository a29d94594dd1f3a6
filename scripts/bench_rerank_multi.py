#!/usr/bin/env python
"""MSMT17-shaped k-reciprocal re-ranking + CMC/mAP, row-sharded over the ranks of a torchrun job
(BASELINE config 4).  Prints one JSON line on rank 0; with WORLD_SIZE=1 it runs the single-call path."""
import argparse, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mp_reid_b200 import engine as E, synth, distributed as D
from mp_reid_b200.reranking import _rerank_device

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="msmt17")
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--k1", type=int, default=20)
ap.add_argument("--k2", type=int, default=6)
ap.add_argument("--precision", default="3xfp16")
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=dev)
qf, gf, q_pid, g_pid, q_cam, g_cam = synth.make_shape(a.workload)
nq, G = qf.shape[0], gf.shape[0]
feats = torch.cat([qf, gf]).to(dev)
q_lo, q_hi = D.shard_bounds(nq, world, rank)
counts = [D.shard_bounds(nq, world, r)[1] - D.shard_bounds(nq, world, r)[0] for r in range(world)]

def step():
    prep = E.prep_rows(feats, normalize=True, precision=a.precision, keep_xn=False)
    if world == 1:
        fin = _rerank_device(prep, nq, a.k1, a.k2, 0.3, a.precision)
    else:
        fin, _ = D.rerank_sharded(prep, nq, a.k1, a.k2, 0.3, a.precision)
    fh, apv, nr = E.rank_eval(fin, q_pid[q_lo:q_hi], g_pid, q_cam[q_lo:q_hi], g_cam, "none")
    if world == 1:
        return E.reduce_cmc_map(fh.cpu().numpy(), apv.cpu().numpy(), nr.cpu().numpy(), 50, G)
    return D.sharded_reduce(fh, apv, nr, counts, 50, G)

res = step(); torch.cuda.synchronize()
if world > 1: dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    res = step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
if world > 1:
    t = torch.tensor([ms], dtype=torch.float64, device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
if rank == 0:
    cmc, mAP = res
    print(json.dumps({"metric": "re-rank ms (prep + (Q+G)^2 distance + k-reciprocal re-ranking + rank/CMC/mAP)", "value": ms, "unit": "ms",
                      "higher_is_better": False, "n_gpus": world, "scaling": "strong", "data": "synthetic",
                      "config": {"workload": f"{a.workload} shape {nq} x {G} x {qf.shape[1]}, k1={a.k1} k2={a.k2} lambda=0.3",
                                 "precision": a.precision, "sharding": "rows of the all-pairs matrix; all-gather of neighbour lists and V0 rows"},
                      "mAP": float(mAP), "rank1": float(cmc[0])}))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
