"""The reference's batch-hard triplet loss (loss/triplet_loss.py) on the GPU, usable under autograd.

    TripletLoss(margin=None, hard_factor=0.0)(global_feat, labels, normalize_feature=False)
        -> (loss, dist_ap, dist_an)                                          loss/triplet_loss.py:106-135
    euclidean_dist(x, y)            sqrt(clamp(|x|^2 + |y|^2 - 2 x.y, 1e-12))   loss/triplet_loss.py:16-31
    hard_example_mining(dist_mat, labels, return_inds=False)                 loss/triplet_loss.py:50-103

The training path (`TripletLoss`) runs two hand-written kernels: a fused pairwise-distance + batch-hard mining forward
and an atomics-free backward (csrc/triplet.cu); `euclidean_dist` / `hard_example_mining` are the stand-alone forward
pieces (tensor-core distance kernel with the sqrt/clamp epilogue, mining kernel) for monitoring.
"""
from __future__ import annotations

import torch
from torch import nn

from . import _lib as L
from . import engine as E


def normalize(x, axis=-1):
    """loss/triplet_loss.py:5-13."""
    return 1. * x / (torch.norm(x, 2, axis, keepdim=True).expand_as(x) + 1e-12)


def euclidean_dist(x: torch.Tensor, y: torch.Tensor, *, precision=None) -> torch.Tensor:
    """loss/triplet_loss.py:16-31: sqrt(clamp(|x|^2 + |y|^2 - 2 x.y, min=1e-12)), [m, n] on the device of x (no autograd)."""
    E.require_cuda()
    px = E.prep_rows(x.detach().cuda().float(), normalize=False, precision=precision, keep_xn=False)
    py = px if y is x else E.prep_rows(y.detach().cuda().float(), normalize=False, precision=precision, keep_xn=False)
    return E.dist_matrix(px, py, "sqrt_euclid", precision)


def hard_example_mining(dist_mat: torch.Tensor, labels, return_inds: bool = False):
    """loss/triplet_loss.py:50-103 -> (dist_ap, dist_an[, p_inds, n_inds]) (no autograd)."""
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


class _BatchHardDistances(torch.autograd.Function):
    """(x [B, D], labels [B]) -> (dist_ap [B], dist_an [B]): euclidean_dist(x, x) + hard_example_mining in one kernel;
    backward = csrc/triplet.cu:k_triplet_backward."""

    @staticmethod
    def forward(ctx, x, labels):
        E.require_cuda()
        lib = L.load()
        xf = x.detach()
        if xf.dtype != torch.float32 or xf.stride(-1) != 1:
            xf = xf.float().contiguous()
        B, D = xf.shape
        dev = xf.device
        lab = E._labels(labels, dev)
        ap = torch.empty((B,), dtype=torch.float32, device=dev)
        an = torch.empty((B,), dtype=torch.float32, device=dev)
        pi = torch.empty((B,), dtype=torch.int64, device=dev)
        ni = torch.empty((B,), dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            L.check(lib.mpreid_triplet_forward(xf.data_ptr(), xf.stride(0), B, D, lab.data_ptr(), ap.data_ptr(), an.data_ptr(),
                                               pi.data_ptr(), ni.data_ptr(), E._stream()), "triplet_forward")
        ctx.save_for_backward(xf, pi, ni, ap, an)
        ctx.in_dtype = x.dtype
        ctx.mark_non_differentiable(pi, ni)
        return ap, an, pi, ni

    @staticmethod
    def backward(ctx, g_ap, g_an, _gp, _gn):
        lib = L.load()
        xf, pi, ni, ap, an = ctx.saved_tensors
        B, D = xf.shape
        g_ap = torch.zeros_like(ap) if g_ap is None else g_ap.float().contiguous()
        g_an = torch.zeros_like(an) if g_an is None else g_an.float().contiguous()
        grad = torch.empty_like(xf)
        with torch.cuda.device(xf.device):
            L.check(lib.mpreid_triplet_backward(xf.data_ptr(), xf.stride(0), B, D, pi.data_ptr(), ni.data_ptr(), ap.data_ptr(), an.data_ptr(),
                                                g_ap.data_ptr(), g_an.data_ptr(), grad.data_ptr(), grad.stride(0), E._stream()),
                    "triplet_backward")
        return grad.to(ctx.in_dtype), None


def batch_hard_distances(x: torch.Tensor, labels, return_inds: bool = False):
    """euclidean_dist(x, x) followed by hard_example_mining, differentiable with respect to x."""
    ap, an, pi, ni = _BatchHardDistances.apply(x, labels)
    return (ap, an, pi, ni) if return_inds else (ap, an)


class TripletLoss(object):
    """loss/triplet_loss.py:106-135, same constructor, call signature and return value."""

    def __init__(self, margin=None, hard_factor=0.0):
        self.margin = margin
        self.hard_factor = hard_factor
        if margin is not None:
            self.ranking_loss = nn.MarginRankingLoss(margin=margin)
        else:
            self.ranking_loss = nn.SoftMarginLoss()

    def __call__(self, global_feat, labels, normalize_feature=False):
        if normalize_feature:
            global_feat = normalize(global_feat, axis=-1)
        dist_ap, dist_an = batch_hard_distances(global_feat, labels)
        dist_ap = dist_ap * (1.0 + self.hard_factor)
        dist_an = dist_an * (1.0 - self.hard_factor)
        y = torch.ones_like(dist_an)
        if self.margin is not None:
            loss = self.ranking_loss(dist_an, dist_ap, y)
        else:
            loss = self.ranking_loss(dist_an - dist_ap, y)
        return loss, dist_ap, dist_an
