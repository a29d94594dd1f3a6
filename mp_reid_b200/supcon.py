"""Stage-1 cached-feature contrastive step: loss/supcontrast.py:SupConLoss and the pair of calls
processor/processor_uniprompt_stage1.py:88-93 makes with it,

    loss_i2t = xent(image_features, text_features, target, target)
    loss_t2i = xent(text_features, image_features, target, target)

on the GPU under autograd.  The similarity matrix text @ image^T comes from the distance kernels (metric "dot": the
tcgen05 GEMM once the batch spans more than one 128-row tile, the SIMT kernel below that); losses, dS and the feature
gradients are csrc/supcon.cu.  `stage1_contrastive_loss` computes both directions from ONE matrix.
"""
from __future__ import annotations

import torch
from torch import nn

from . import _lib as L
from . import engine as E


def _similarity(a: torch.Tensor, b: torch.Tensor, precision=None) -> torch.Tensor:
    """a @ b.T, fp32 [Ba, Bb] (loss/supcontrast.py:23)."""
    Ba, Bb = a.shape[0], b.shape[0]
    if precision is None:
        precision = "simt" if max(Ba, Bb) <= 256 else E.default_precision()
    pa = E.prep_rows(a, normalize=False, precision=precision, keep_xn=True)
    pb = E.prep_rows(b, normalize=False, precision=precision, keep_xn=True)
    return E.dist_matrix(pa, pb, "dot", precision)


class _SupCon(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, la, lb, temperature, both, precision):
        E.require_cuda()
        lib = L.load()
        af = a.detach().float().contiguous()
        bf = b.detach().float().contiguous()
        dev = af.device
        Ba, D = af.shape
        Bb = bf.shape[0]
        la_d, lb_d = E._labels(la, dev), E._labels(lb, dev)
        S = _similarity(af, bf, precision)
        need_a, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        ga = torch.empty_like(af) if need_a else None
        gb = torch.empty_like(bf) if need_b else None
        loss = torch.empty((3,), dtype=torch.float32, device=dev)
        nbytes = lib.mpreid_supcon_workspace_bytes(Ba, Bb)
        ws = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            L.check(lib.mpreid_supcon_step(S.data_ptr(), S.stride(0), Ba, Bb, la_d.data_ptr(), lb_d.data_ptr(), float(temperature), int(bool(both)),
                                           1.0, 1.0, af.data_ptr(), af.stride(0), bf.data_ptr(), bf.stride(0), D,
                                           loss.data_ptr(), E._ptr(ga), ga.stride(0) if need_a else D, E._ptr(gb), gb.stride(0) if need_b else D,
                                           ws.data_ptr(), nbytes, E._stream()), "supcon_step")
        ctx.grads = (ga, gb)
        ctx.dtypes = (a.dtype, b.dtype)
        total = (loss[2] if both else loss[0]).clone()
        ctx.mark_non_differentiable(loss)
        return total, loss

    @staticmethod
    def backward(ctx, g, _g_terms):
        ga, gb = ctx.grads
        return (None if ga is None else (ga * g).to(ctx.dtypes[0]), None if gb is None else (gb * g).to(ctx.dtypes[1]),
                None, None, None, None, None)


class SupConLoss(nn.Module):
    """loss/supcontrast.py:10-31 (same constructor and forward signature)."""

    def __init__(self, device=None):
        super(SupConLoss, self).__init__()
        self.device = device
        self.temperature = 1.0

    def forward(self, text_features, image_features, t_label, i_targets):
        loss, _ = _SupCon.apply(text_features, image_features, t_label, i_targets, self.temperature, False, None)
        return loss


def stage1_contrastive_loss(image_features, text_features, target, temperature: float = 1.0, precision=None, return_terms: bool = False):
    """processor/processor_uniprompt_stage1.py:88-93: SupConLoss(image, text) + SupConLoss(text, image) from one
    similarity matrix (its rows give the first term, its columns the second).  Gradients flow to whichever input needs them
    (the cached image features normally do not)."""
    loss, terms = _SupCon.apply(image_features, text_features, target, target, temperature, True, precision)
    return (loss, terms[0], terms[1]) if return_terms else loss
