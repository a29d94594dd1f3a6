"""Forward pass of the reference's batch-hard triplet distance helpers (loss/triplet_loss.py:16-103),
SURVEY.md 8f-3: the same contraction as the evaluation path with the sqrt / clamp epilogue, plus the
hardest-positive / hardest-negative selection.  Forward only (no autograd): for monitoring / mining,
not a drop-in for the training loss.
"""
from __future__ import annotations

import torch

from . import _lib as L
from . import engine as E


def euclidean_dist(x: torch.Tensor, y: torch.Tensor, *, precision=None) -> torch.Tensor:
    """loss/triplet_loss.py:16-31: sqrt(clamp(|x|^2 + |y|^2 - 2 x.y, min=1e-12)), [m, n] on the device of x."""
    E.require_cuda()
    px = E.prep_rows(x.detach().cuda().float(), normalize=False, precision=precision, keep_xn=False)
    py = px if y is x else E.prep_rows(y.detach().cuda().float(), normalize=False, precision=precision, keep_xn=False)
    return E.dist_matrix(px, py, "sqrt_euclid", precision)


def hard_example_mining(dist_mat: torch.Tensor, labels, return_inds: bool = False):
    """loss/triplet_loss.py:50-103 -> (dist_ap, dist_an[, p_inds, n_inds])."""
    E.require_cuda()
    lib = L.load()
    assert dist_mat.dim() == 2 and dist_mat.size(0) == dist_mat.size(1)
    d = dist_mat.detach()
    if not d.is_cuda:
        d = d.cuda()
    if d.dtype != torch.float32 or d.stride(1) != 1:
        d = d.float().contiguous()
    N = d.shape[0]
    lab = E._labels(labels, d.device)
    ap = torch.empty((N,), dtype=torch.float32, device=d.device)
    an = torch.empty((N,), dtype=torch.float32, device=d.device)
    pi = torch.empty((N,), dtype=torch.int64, device=d.device) if return_inds else None
    ni = torch.empty((N,), dtype=torch.int64, device=d.device) if return_inds else None
    with torch.cuda.device(d.device):
        L.check(lib.mpreid_hard_example_mining(d.data_ptr(), d.stride(0), N, lab.data_ptr(), ap.data_ptr(), an.data_ptr(),
                                               E._ptr(pi), E._ptr(ni), E._stream()), "hard_example_mining")
    return (ap, an, pi, ni) if return_inds else (ap, an)
