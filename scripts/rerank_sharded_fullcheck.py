"""torchrun check at a BASELINE shape: the row-sharded fused re-ranking against the one-GPU fused pipeline run on the same
rank, element by element.  usage: torchrun ... scripts/rerank_sharded_fullcheck.py [msmt17|market]"""
import os, sys
import numpy as np, torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mp_reid_b200 import engine as E, synth, distributed as MD
from mp_reid_b200.reranking import _rerank_device

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
shape = sys.argv[1] if len(sys.argv) > 1 else "msmt17"
qf, gf, q_pid, g_pid, q_cam, g_cam = synth.make_shape(shape)
nq = qf.shape[0]
prep = E.prep_rows(torch.cat([qf, gf]).to(dev), normalize=True, keep_xn=True)
os.environ["MPREID_RERANK_FUSED"] = "1"
want = _rerank_device(prep, nq, 20, 6, 0.3)
got, ids = MD.rerank_sharded(prep, nq, 20, 6, 0.3)
diff = (got - want[ids]).abs()
nbad = int((diff != 0).sum())
rows_bad = int((diff != 0).any(1).sum())
print(f"[rank {rank}/{world}] rows {got.shape[0]} (first ids {ids[:3].tolist()}), differing elements {nbad} in {rows_bad} rows, max |diff| {float(diff.max()) if diff.numel() else 0.0:.3e}",
      flush=True)
if nbad:
    r = int((diff != 0).any(1).nonzero()[0])
    c = (diff[r] != 0).nonzero().flatten()[:5]
    print(f"[rank {rank}] e.g. query {int(ids[r])}: cols {c.tolist()} got {got[r, c].tolist()} want {want[ids[r], c].tolist()}", flush=True)
dist.barrier()
dist.destroy_process_group()
