import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mp_reid_b200 import engine as E
prec = sys.argv[1] if len(sys.argv) > 1 else "3xtf32"
torch.manual_seed(0)
q = torch.randn(130, 100, device="cuda"); g = torch.randn(257, 100, device="cuda")
pq, pg = E.prep_rows(q, False, prec), E.prep_rows(g, False, prec)
d = E.dist_matrix(pq, pg, "sqeuclid", prec)
torch.cuda.synchronize()
ref = (q.double()**2).sum(1)[:, None] + (g.double()**2).sum(1)[None] - 2 * q.double() @ g.double().T
print(prec, "max err", float((d.double() - ref).abs().max()))
