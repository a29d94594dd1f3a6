import os, sys, time, io, contextlib
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mp_reid_b200 import metrics, synth
qf, gf, q_pid, g_pid, q_cam, g_cam = synth.make_shape("msmt17")
allf = torch.cat([qf, gf]); pids = np.concatenate([q_pid, g_pid]); cams = np.concatenate([q_cam, g_cam])
batches = [(allf[s:s+8192].clone().pin_memory(), pids[s:s+8192], cams[s:s+8192]) for s in range(0, allf.shape[0], 8192)]
Q = qf.shape[0]
def run(tag, sync_after_update):
    for it in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        ev = metrics.R1_mAP_eval(Q, feat_norm=True); ev.reset()
        for f, p, c in batches: ev.update((f, p, c))
        t1 = time.perf_counter()
        if sync_after_update: torch.cuda.synchronize()
        t2 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            cmc, mAP, *_ = ev.compute()
        torch.cuda.synchronize(); t3 = time.perf_counter()
        print(f"{tag} it{it}: update {1e3*(t1-t0):.2f} ms, wait-copies {1e3*(t2-t1):.2f} ms, compute {1e3*(t3-t2):.2f} ms, total {1e3*(t3-t0):.2f} ms  mAP {mAP:.6f}")
run("overlap", False)
run("serial ", True)
# pure python overhead of update: feed CUDA tensors
dev_batches = [(f.cuda(), p, c) for f, p, c in batches]
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ev = metrics.R1_mAP_eval(Q, feat_norm=True); ev.reset()
    for f, p, c in dev_batches: ev.update((f, p, c))
    t1 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        cmc, mAP, *_ = ev.compute()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"device-resident batches it{it}: update {1e3*(t1-t0):.2f} ms compute {1e3*(t2-t1):.2f} ms")
import cProfile, pstats
ev = metrics.R1_mAP_eval(Q, feat_norm=True); ev.reset()
for f, p, c in dev_batches: ev.update((f, p, c))
pr = cProfile.Profile(); pr.enable()
with contextlib.redirect_stdout(io.StringIO()):
    ev.compute()
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
