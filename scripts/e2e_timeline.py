"""Timeline of one R1_mAP_eval pass from pinned host batches: what runs after the last host->device copy ends."""
import contextlib, io, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mp_reid_b200 import metrics, synth
from torch.profiler import profile, ProfilerActivity

qf, gf, q_pid, g_pid, q_cam, g_cam = synth.make_shape("msmt17")
allf = torch.cat([qf, gf]); pids = np.concatenate([q_pid, g_pid]); cams = np.concatenate([q_cam, g_cam])
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
batches = [(allf[s:s + B].clone().pin_memory(), pids[s:s + B], cams[s:s + B]) for s in range(0, allf.shape[0], B)]
Q = qf.shape[0]

def one():
    ev = metrics.R1_mAP_eval(Q, feat_norm=True); ev.reset()
    for f, p, c in batches: ev.update((f, p, c))
    with contextlib.redirect_stdout(io.StringIO()):
        return ev.compute()[1]

for _ in range(3): one()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    one(); torch.cuda.synchronize()
prof.export_chrome_trace("gpurun_out/e2e_trace.json")
tr = json.load(open("gpurun_out/e2e_trace.json"))["traceEvents"]
gpu = [e for e in tr if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "ts" in e]
gpu.sort(key=lambda e: e["ts"])
t0 = gpu[0]["ts"]
h2d = [e for e in gpu if e["cat"] == "gpu_memcpy" and "HtoD" in e["name"] and e["dur"] > 100]
last = max(e["ts"] + e["dur"] for e in h2d)
print(f"first GPU activity at 0, last big H2D ends at {(last - t0) / 1e3:.3f} ms, last GPU activity ends at {(max(e['ts'] + e['dur'] for e in gpu) - t0) / 1e3:.3f} ms")
print("GPU activity starting within 0.3 ms before the end of the last copy, and after:")
for e in gpu:
    if e["ts"] + e["dur"] >= last - 300:
        print(f"  +{(e['ts'] - last) / 1e3:8.3f} ms  dur {e['dur'] / 1e3:7.3f} ms  {e['cat']:10s} {e['name'][:70]}")
cpu = [e for e in tr if e.get("cat") in ("cpu_op", "cuda_runtime", "user_annotation") and "ts" in e and e["ts"] >= last - 200 and e.get("dur", 0) > 50]
print("host calls > 50 us after the last copy:")
for e in sorted(cpu, key=lambda e: e["ts"])[:40]:
    print(f"  +{(e['ts'] - last) / 1e3:8.3f} ms  dur {e['dur'] / 1e3:7.3f} ms  {e['cat']:14s} {e['name'][:60]}")
