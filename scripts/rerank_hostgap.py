import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mp_reid_b200 import engine as E, synth
from mp_reid_b200.reranking import _rerank_device
qf, gf, q_pid, g_pid, q_cam, g_cam = synth.make_shape("msmt17")
dev = torch.device("cuda:0")
sub = torch.cat([qf, gf]).to(dev); Q = qf.shape[0]; G = gf.shape[0]
lab = [torch.from_numpy(x).to(dev) for x in (q_pid, g_pid, q_cam, g_cam)]
for it in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    e[0].record()
    p = E.prep_rows(sub, normalize=True, precision="3xfp16", keep_xn=False)
    t1 = time.perf_counter(); e[1].record()
    dfin = _rerank_device(p, Q, 20, 6, 0.3, "3xfp16")
    t2 = time.perf_counter(); e[2].record()
    fh, ap, nr = E.rank_eval(dfin, *lab, "none")
    e[3].record()
    r = E.reduce_cmc_map(fh.cpu().numpy(), ap.cpu().numpy(), nr.cpu().numpy(), 50, G)
    torch.cuda.synchronize(); t3 = time.perf_counter()
    print(f"it{it}: host prep call {1e3*(t1-t0):.3f} ms, rerank call {1e3*(t2-t1):.3f} ms, total wall {1e3*(t3-t0):.3f} ms | gpu prep {e[0].elapsed_time(e[1]):.3f} rerank {e[1].elapsed_time(e[2]):.3f} rank {e[2].elapsed_time(e[3]):.3f}")
