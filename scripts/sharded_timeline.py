"""torchrun: GPU timeline (rank 0) of one cooperative evaluation at MSMT17 shape."""
import contextlib, io, json, os, sys, time
import numpy as np, torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mp_reid_b200 import metrics, synth, distributed as MD
from torch.profiler import profile, ProfilerActivity
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); os.environ["MPREID_DEVICE"] = f"cuda:{local}"
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
qf, gf, q_pid, g_pid, q_cam, g_cam = synth.make_shape("msmt17")
Q, G = qf.shape[0], gf.shape[0]
g_lo, g_hi = MD.aligned_shard_bounds(G, world, rank)
B = 8192
batches = [(qf[s:s + B].clone().pin_memory(), q_pid[s:s + B], q_cam[s:s + B]) for s in range(0, Q, B)]
batches += [(gf[s:min(g_hi, s + B)].clone().pin_memory(), g_pid[s:min(g_hi, s + B)], g_cam[s:min(g_hi, s + B)]) for s in range(g_lo, g_hi, B)]
def one():
    ev = MD.sharded_evaluator(Q); ev.reset()
    for f, p, c in batches: ev.update((f, p, c))
    with contextlib.redirect_stdout(io.StringIO()):
        return ev.compute()[1]
for _ in range(3): one()
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
for _ in range(5): m = one()
torch.cuda.synchronize(); dist.barrier()
if rank == 0: print(f"world {world}: {1e3 * (time.perf_counter() - t0) / 5:.2f} ms per evaluation, mAP {m:.9f}")
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    one(); torch.cuda.synchronize()
dist.barrier()
if rank == 0:
    prof.export_chrome_trace("gpurun_out/sh_trace.json")
    tr = json.load(open("gpurun_out/sh_trace.json"))["traceEvents"]
    gpu = sorted([e for e in tr if e.get("cat") in ("kernel", "gpu_memcpy") and "ts" in e], key=lambda e: e["ts"])
    t0 = gpu[0]["ts"]
    for e in gpu:
        if e["dur"] > 40:
            print(f"{(e['ts'] - t0) / 1e3:8.3f} +{e['dur'] / 1e3:6.3f} {e['cat'][:6]:6s} {e['name'][:48]}")
dist.destroy_process_group()
