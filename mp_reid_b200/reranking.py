"""Drop-in for the reference's ``utils/reranking.py``.

    re_ranking(probFea, galFea, k1, k2, lambda_value, local_distmat=None, only_local=False)
        -> numpy float32 [query_num, gallery_num]                          utils/reranking.py:29-100
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import engine as E


def _device():
    E.require_cuda()
    return torch.device(os.environ.get("MPREID_DEVICE", f"cuda:{torch.cuda.current_device()}"))


def _all_pairs(prep: E.Prepared, precision=None, row_max: torch.Tensor | None = None) -> torch.Tensor:
    """utils/reranking.py:36-41: squared distances of the stacked features against themselves
    (+ the per-row maxima of :46, taken in the GEMM epilogue)."""
    n = prep.n
    ld = (n + 31) // 32 * 32
    buf = torch.empty((n, ld), dtype=torch.float32, device=prep.sqnorm.device)[:, :n]
    return E.dist_matrix_all_pairs(prep, precision, out=buf, row_max=row_max)


def _rerank_device(prep: E.Prepared, query_num: int, k1: int, k2: int, lambda_value: float, precision=None,
                   local_distmat: torch.Tensor | None = None) -> torch.Tensor:
    if local_distmat is not None:  # :43-44  (orientation: ours is the transpose of the reference's)
        dall = _all_pairs(prep, precision)
        dall.add_(local_distmat.t())
        return E.rerank_from_dist(dall, query_num, k1, k2, lambda_value)
    row_max = torch.empty((prep.n,), dtype=torch.float32, device=prep.sqnorm.device)
    dall = _all_pairs(prep, precision, row_max)
    return E.rerank_from_dist(dall, query_num, k1, k2, lambda_value, row_max=row_max)


def re_ranking(probFea, galFea, k1, k2, lambda_value, local_distmat=None, only_local=False, *, precision=None):
    dev = _device()
    query_num = int(probFea.shape[0])   # tensors and numpy arrays alike
    if only_local:  # :33-34
        d = torch.as_tensor(np.asarray(local_distmat), dtype=torch.float32).to(dev)
        out = E.rerank_from_dist(d.t().contiguous(), query_num, k1, k2, lambda_value)
        return out.cpu().numpy()
    to_dev = lambda x: (x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x))).to(dev, dtype=torch.float32)
    feat = torch.cat([to_dev(probFea), to_dev(galFea)])
    prep = E.prep_rows(feat, normalize=False, precision=precision, keep_xn=False)
    loc = None
    if local_distmat is not None:
        loc = torch.as_tensor(np.asarray(local_distmat), dtype=torch.float32).to(dev)
    out = _rerank_device(prep, query_num, k1, k2, lambda_value, precision, loc)
    return out.cpu().numpy()
