"""Stage timeline of one re-ranking pass (CUDA events between the stages, engine.timeline_*).
usage: python scripts/rerank_stages.py [msmt17|market] [reps]    (MPREID_RERANK_FUSED=0/1 selects the pipeline)"""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mp_reid_b200 import engine as E, synth
from mp_reid_b200.reranking import _rerank_device

shape = sys.argv[1] if len(sys.argv) > 1 else "msmt17"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
k1, k2 = int(os.environ.get("K1", 20)), int(os.environ.get("K2", 6))
qf, gf, q_pid, g_pid, q_cam, g_cam = synth.make_shape(shape)
dev = torch.device("cuda:0")
feats = torch.cat([qf, gf]).to(dev)
nq = qf.shape[0]
lab = [torch.from_numpy(a).to(dev) for a in (q_pid, g_pid, q_cam, g_cam)]


def run():
    p = E.prep_rows(feats, normalize=True, keep_xn=True)
    E.mark("prep")
    d = _rerank_device(p, nq, k1, k2, 0.3)
    E.mark("rerank.rest")
    fh, ap, nr = E.rank_eval(d, lab[0], lab[1], lab[2], lab[3])
    E.mark("rank_eval")
    return E.reduce_cmc_map(fh.cpu().numpy(), ap.cpu().numpy(), nr.cpu().numpy(), 50, gf.shape[0])


run(); torch.cuda.synchronize()
acc = {}
for _ in range(reps):
    E.timeline_start()
    cmc, mAP = run()
    for name, ms in E.timeline_stop():
        acc.setdefault(name, []).append(ms)
out = {k: float(np.mean(v)) for k, v in acc.items()}
out["total_ms"] = float(sum(out.values()))
out["mAP"] = float(mAP)
out["shape"] = shape
out["fused"] = os.environ.get("MPREID_RERANK_FUSED", "auto")
print(json.dumps(out))
