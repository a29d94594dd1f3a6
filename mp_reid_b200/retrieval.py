"""Large-scale retrieval (BASELINE config 5: 100k queries x 1M gallery x 768-d, top-100 + CMC/mAP).

The Q x G matrix of that config is 400 GB and is never materialised whole: queries are processed in
chunks whose distance block fits a fixed HBM budget; each block goes through the same kernels as the
evaluator (tcgen05 distance, streaming top-k, counting rank/AP) and is then overwritten.  The gallery
planes are prepared once.  Results are identical to a single pass because queries are independent
(utils/metrics.py:47-80 keeps no cross-query state).
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import engine as E


def _chunk_rows(G: int, budget_bytes: int) -> int:
    rows = max(128, budget_bytes // (4 * ((G + 31) // 32 * 32)))
    return int(rows // 128 * 128)


def retrieve(qf: torch.Tensor, gf: torch.Tensor, q_pids, g_pids, q_camids=None, g_camids=None, k: int = 100,
             feat_norm=True, metric: str = "sqeuclid", precision: str | None = None, junk: str | None = None,
             max_rank: int = 50, block_bytes: int | None = None, return_device: bool = False):
    """-> dict(topk int32 [Q, k] (the first k entries of the stable argsort of every distance row),
               cmc float32[max_rank], mAP float64, first_hit, ap, num_rel (per query, numpy)).

    qf / gf: CUDA fp32 [Q, D] / [G, D] (already on the device that runs the query shard).
    """
    E.require_cuda()
    assert qf.is_cuda and gf.is_cuda and qf.device == gf.device
    Q, G = qf.shape[0], gf.shape[0]
    dev = qf.device
    block_bytes = block_bytes or int(os.environ.get("MPREID_BLOCK_BYTES", str(16 << 30)))
    rows = min(Q, _chunk_rows(G, block_bytes))
    norm = bool(feat_norm)
    g = E.prep_rows(gf, normalize=norm, precision=precision, keep_xn=False)
    q_pid_d, g_pid_d = E._labels(q_pids, dev), E._labels(g_pids, dev)
    junk_on = (junk or E.default_junk()).lower() != "none"
    q_cam_d = E._labels(q_camids, dev) if junk_on else None
    g_cam_d = E._labels(g_camids, dev) if junk_on else None
    block = E.alloc_dist(rows, G, dev)
    topk = torch.empty((Q, k), dtype=torch.int32, device=dev)
    fh_all = torch.empty((Q,), dtype=torch.int32, device=dev)
    ap_all = torch.empty((Q,), dtype=torch.float64, device=dev)
    nr_all = torch.empty((Q,), dtype=torch.int32, device=dev)
    for lo in range(0, Q, rows):
        hi = min(Q, lo + rows)
        q = E.prep_rows(qf[lo:hi], normalize=norm, precision=precision, keep_xn=False)
        d = E.dist_matrix(q, g, metric, precision, out=block[: hi - lo])
        topk[lo:hi] = E.row_topk(d, k)
        fh, ap, nr = E.rank_eval(d, q_pid_d[lo:hi], g_pid_d, None if q_cam_d is None else q_cam_d[lo:hi], g_cam_d, junk)
        fh_all[lo:hi], ap_all[lo:hi], nr_all[lo:hi] = fh, ap, nr
    out = dict(topk=topk, first_hit=fh_all, ap=ap_all, num_rel=nr_all, chunk_rows=rows)
    if return_device:
        return out
    fh, ap, nr = fh_all.cpu().numpy(), ap_all.cpu().numpy(), nr_all.cpu().numpy()
    cmc, mAP = E.reduce_cmc_map(fh, ap, nr, min(max_rank, G), G)
    out.update(topk=topk.cpu().numpy(), first_hit=fh, ap=ap, num_rel=nr, cmc=cmc, mAP=mAP)
    return out
