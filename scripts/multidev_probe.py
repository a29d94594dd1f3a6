"""Single-process multi-GPU evaluator (MPREID_DEVICES) vs one GPU at MSMT17 shape, from pinned host batches."""
import contextlib, io, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mp_reid_b200 import metrics, synth
qf, gf, q_pid, g_pid, q_cam, g_cam = synth.make_shape("msmt17")
allf = torch.cat([qf, gf]); pids = np.concatenate([q_pid, g_pid]); cams = np.concatenate([q_cam, g_cam])
B = 8192
batches = [(allf[s:s + B].clone().pin_memory(), pids[s:s + B], cams[s:s + B]) for s in range(0, allf.shape[0], B)]
dev_batches = [(f.cuda(), p, c) for f, p, c in batches]
Q = qf.shape[0]
def one(bs):
    ev = metrics.R1_mAP_eval(Q, feat_norm=True); ev.reset()
    for f, p, c in bs: ev.update((f, p, c))
    with contextlib.redirect_stdout(io.StringIO()):
        return ev.compute()[1]
for spec in ["", "all"]:
    os.environ["MPREID_DEVICES"] = spec
    for name, bs in (("host batches", batches), ("device-resident batches", dev_batches)):
        for _ in range(3): m = one(bs)
        for d in range(torch.cuda.device_count()): torch.cuda.synchronize(d)
        t0 = time.perf_counter()
        for _ in range(5): m = one(bs)
        for d in range(torch.cuda.device_count()): torch.cuda.synchronize(d)
        print(f"MPREID_DEVICES='{spec}' ({torch.cuda.device_count()} visible) {name}: {1e3 * (time.perf_counter() - t0) / 5:.2f} ms per evaluation, mAP {m:.9f}", flush=True)
