import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mp_reid_b200 import engine as E, synth
prec = sys.argv[1] if len(sys.argv) > 1 else "3xfp16"
qf, gf, *_ = synth.make_shape("msmt17")
dev = torch.device("cuda:0")
feats = torch.cat([qf, gf]).to(dev)
Q, G = qf.shape[0], gf.shape[0]
p = E.prep_rows(feats, True, prec, keep_xn=False)
q, g = p.rows(0, Q), p.rows(Q, Q + G)
out = E.alloc_dist(Q, G, dev)
for _ in range(3): E.dist_matrix(q, g, "sqeuclid", prec, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): E.dist_matrix(q, g, "sqeuclid", prec, out=out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
ref = E.dist_matrix(q, g, "sqeuclid", "3xfp16" if prec != "3xfp16" else prec)
print(f"ROWB={os.environ.get('MPREID_GEMM_ROWB','128')} {prec}: {ms:.3f} ms  {2*Q*G*1280/ms/1e9:.1f} TFLOP/s algorithmic  checksum {float(out.double().sum()):.6f}")
