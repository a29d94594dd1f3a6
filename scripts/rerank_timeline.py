"""GPU timeline of one MSMT17-shaped re-ranking pass (single GPU): kernels, gaps, host syncs."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mp_reid_b200 import engine as E, synth
from mp_reid_b200.reranking import _rerank_device
from torch.profiler import profile, ProfilerActivity
shape = sys.argv[1] if len(sys.argv) > 1 else "msmt17"
qf, gf, q_pid, g_pid, q_cam, g_cam = synth.make_shape(shape)
dev = torch.device("cuda:0")
sub = torch.cat([qf, gf]).to(dev); Q = qf.shape[0]; G = gf.shape[0]
lab = [torch.from_numpy(x).to(dev) for x in (q_pid, g_pid, q_cam, g_cam)]
def rr():
    p = E.prep_rows(sub, normalize=True, precision="3xfp16", keep_xn=False)
    dfin = _rerank_device(p, Q, 20, 6, 0.3, "3xfp16")
    fh, ap, nr = E.rank_eval(dfin, *lab, "none")
    return E.reduce_cmc_map(fh.cpu().numpy(), ap.cpu().numpy(), nr.cpu().numpy(), 50, G)
for _ in range(2): rr()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    rr(); torch.cuda.synchronize()
prof.export_chrome_trace("gpurun_out/rr_trace.json")
tr = json.load(open("gpurun_out/rr_trace.json"))["traceEvents"]
gpu = sorted([e for e in tr if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "ts" in e], key=lambda e: e["ts"])
t0 = gpu[0]["ts"]; prev_end = t0
busy = 0
for e in gpu:
    gap = e["ts"] - prev_end
    if e["dur"] > 30 or gap > 30:
        print(f"{(e['ts'] - t0) / 1e3:8.3f} ms  gap {gap / 1e3:6.3f}  dur {e['dur'] / 1e3:7.3f}  {e['cat'][:6]:6s} {e['name'][:60]}")
    busy += e["dur"]; prev_end = max(prev_end, e["ts"] + e["dur"])
print(f"total {(prev_end - t0) / 1e3:.3f} ms, busy {busy / 1e3:.3f} ms")
