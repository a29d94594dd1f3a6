#!/usr/bin/env python
"""BASELINE config 5: 100k queries x 1M gallery x 768-d, top-100 + CMC/mAP, chunked so that the
400 GB distance matrix never exists.  One JSON line; run under torchrun for query sharding."""
import argparse, json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mp_reid_b200 import engine as E, retrieval, synth

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--k", type=int, default=100)
ap.add_argument("--precision", default="3xfp16")
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=dev)
s = synth.SHAPES["retrieval"]
Q, G = int(s.Q * a.scale), int(s.G * a.scale)
# synthetic features generated on the device (3 GB of host randn would take longer than the benchmark)
gen = torch.Generator(device=dev).manual_seed(s.seed)
n_id = max(2, int(s.n_id * a.scale))
centers = torch.randn(n_id, s.D, device=dev, generator=gen)
g_pid = torch.randint(0, n_id, (G,), device=dev, generator=gen)
gf = centers[g_pid] + s.sigma * torch.randn(G, s.D, device=dev, generator=gen)
from mp_reid_b200.distributed import shard_bounds
q_pid_all = torch.randint(0, n_id, (Q,), device=dev, generator=gen)
lo, hi = shard_bounds(Q, world, rank)
q_pid = q_pid_all[lo:hi]
qf = centers[q_pid] + s.sigma * torch.randn(hi - lo, s.D, device=dev, generator=gen)
del centers
torch.cuda.synchronize()
def step():
    return retrieval.retrieve(qf, gf, q_pid, g_pid, k=a.k, precision=a.precision, return_device=True)
r = step(); torch.cuda.synchronize()
if world > 1: dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    r = step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
if world > 1:
    t = torch.tensor([ms], dtype=torch.float64, device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
    from mp_reid_b200.distributed import gather_per_query
    counts = [shard_bounds(Q, world, r_)[1] - shard_bounds(Q, world, r_)[0] for r_ in range(world)]
    fh, apv, nr = gather_per_query(r["first_hit"], r["ap"], r["num_rel"], counts)
else:
    fh, apv, nr = r["first_hit"].cpu().numpy(), r["ap"].cpu().numpy(), r["num_rel"].cpu().numpy()
if rank == 0:
    cmc, mAP = E.reduce_cmc_map(fh, apv, nr, 50, G)
    print(json.dumps({"metric": "query x gallery pairs/sec (dist + top-%d + rank + mAP)" % a.k, "value": Q * G / (ms * 1e-3), "unit": "pairs/s",
                      "n_gpus": world, "ms_per_step": ms, "scaling": "strong", "data": "synthetic (generated on device)",
                      "config": {"workload": f"retrieval {Q} x {G} x {s.D}, top-{a.k} + CMC/mAP, query chunks of {r['chunk_rows']} rows",
                                 "precision": a.precision}, "mAP": float(mAP), "rank1": float(cmc[0]),
                      "gflops_algorithmic": 2.0 * Q * G * s.D / 1e9}))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
