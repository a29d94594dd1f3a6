import os, sys, time, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mp_reid_b200 import engine as E, synth
qf, gf, q_pid, g_pid, q_cam, g_cam = synth.make_shape("msmt17")
dev = torch.device("cuda:0")
feats = torch.cat([qf, gf]).to(dev); Q, G = qf.shape[0], gf.shape[0]
p = E.prep_rows(feats, True, keep_xn=False)
d = E.dist_matrix(p.rows(0, Q), p.rows(Q, Q + G))
lab = [torch.from_numpy(x).to(dev) for x in (q_pid, g_pid, q_cam, g_cam)]
for _ in range(3): E.rank_eval_async(d, *lab, "none")
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50): r = E.rank_eval_async(d, *lab, "none")
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"enqueue {1e3*(t1-t0)/50:.3f} ms/call, total {1e3*(t2-t0)/50:.3f} ms/call")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(50): r = E.rank_eval_async(d, *lab, "none")
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(8)
# per-kernel device time inside the live (back-to-back) sequence
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(10): r = E.rank_eval_async(d, *lab, "none")
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))
