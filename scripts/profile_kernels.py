#!/usr/bin/env python
"""Short driver for `ncu --set full`: one MSMT17-shaped pass (prep, distance, rank/AP) and one
Market-shaped re-ranking pass, each kernel launched twice."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mp_reid_b200 import engine as E, synth
from mp_reid_b200.reranking import _rerank_device

prec = sys.argv[1] if len(sys.argv) > 1 else "3xfp16"
qf, gf, q_pid, g_pid, q_cam, g_cam = synth.make_shape("msmt17")
dev = torch.device("cuda:0")
feats = torch.cat([qf, gf]).to(dev)
Q, G = qf.shape[0], gf.shape[0]
for _ in range(2):
    p = E.prep_rows(feats, True, prec, keep_xn=False)
    d = E.dist_matrix(p.rows(0, Q), p.rows(Q, Q + G), "sqeuclid", prec)
    fh, ap, nr = E.rank_eval(d, q_pid, g_pid, q_cam, g_cam, "none")
torch.cuda.synchronize()
sub = torch.cat([feats[:3368], feats[Q:Q + 15913]])
for _ in range(2):
    ps = E.prep_rows(sub, True, prec, keep_xn=True)
    out = _rerank_device(ps, 3368, 20, 6, 0.3, prec)
torch.cuda.synchronize()
print("done", float(ap.sum()), float(out.sum()))
