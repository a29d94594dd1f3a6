"""Drop-in for the reference's ``utils/metrics.py`` — same names, signatures, prints and errors,
computed on the GPU through libmpreid_b200.so.

    R1_mAP_eval(num_query, max_rank=50, feat_norm=True, reranking=False)   utils/metrics.py:91-134
    eval_func(distmat, q_pids, g_pids, q_camids, g_camids, max_rank=50)    utils/metrics.py:28-88
    euclidean_distance(qf, gf)                                             utils/metrics.py:7-13
    cosine_similarity(qf, gf)                                              utils/metrics.py:15-25

Extra behaviour is reachable only through keyword-only arguments or environment variables
(MPREID_PRECISION = 3xtf32 | bf16 | simt, MPREID_JUNK = none | pid_cam, MPREID_DEVICE = cuda:N).
Tie contract: equal distances rank by ascending gallery index (np.argsort(kind='stable')); the
reference calls numpy's unstable default, so its own tie order is unspecified.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import engine as E
from .reranking import re_ranking, _rerank_device


def _device():
    E.require_cuda()
    return torch.device(os.environ.get("MPREID_DEVICE", f"cuda:{torch.cuda.current_device()}"))


def _to_device(x, dev):
    if not isinstance(x, torch.Tensor):
        x = torch.as_tensor(np.asarray(x))
    if x.dtype != torch.float32:
        x = x.float()
    return x.to(dev, non_blocking=True)


class LazyDistmat:
    """ndarray-like handle on the device-resident [Q, G] distance matrix.

    Every caller of R1_mAP_eval.compute() discards the matrix (processor/processor.py:154,
    processor_uniprompt_stage2.py:213,261); copying 3.8 GB (MSMT17 shape) to the host eagerly would
    dominate the evaluation, so the copy happens on first use (np.asarray(d), d[...], d.numpy()).
    """

    def __init__(self, dev_tensor: torch.Tensor):
        self._dev = dev_tensor
        self._host = None
        self.shape = tuple(dev_tensor.shape)
        self.dtype = np.dtype(np.float32)
        self.ndim = 2

    @property
    def device_tensor(self) -> torch.Tensor:
        return self._dev

    def numpy(self) -> np.ndarray:
        if self._host is None:
            self._host = self._dev.cpu().numpy()
        return self._host

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a if dtype is None else a.astype(dtype, copy=False)

    def __getitem__(self, item):
        return self.numpy()[item]

    def __len__(self):
        return self.shape[0]

    def __repr__(self):
        return f"LazyDistmat(shape={self.shape}, device={self._dev.device})"


def _distance(qf, gf, metric, precision=None, to_host=True):
    dev = _device()
    q = E.prep_rows(_to_device(qf, dev), normalize=False, precision=precision, keep_xn=False)
    g = E.prep_rows(_to_device(gf, dev), normalize=False, precision=precision, keep_xn=False)
    d = E.dist_matrix(q, g, metric, precision)
    return d.cpu().numpy() if to_host else d


def euclidean_distance(qf, gf, *, precision=None):
    """utils/metrics.py:7-13 — SQUARED euclidean distance, fp32, numpy [m, n]."""
    return _distance(qf, gf, "sqeuclid", precision)


def cosine_similarity(qf, gf, *, precision=None):
    """utils/metrics.py:15-25 — arccos of the clipped cosine, fp32 radians, numpy [m, n]."""
    return _distance(qf, gf, "arccos", precision)


def one_minus_cosine(qf, gf, *, precision=None):
    """processor/processor_uniprompt_stage2.py:466-468 — 1 - qf @ gf.T on already-normalised features."""
    return _distance(qf, gf, "one_minus_dot", precision)


def _eval_device(dist_dev, q_pids, g_pids, q_camids, g_camids, max_rank, junk, denominators="valid"):
    num_q, num_g = dist_dev.shape
    if num_g < max_rank:  # utils/metrics.py:36-38
        max_rank = num_g
        print("Note: number of gallery samples is quite small, got {}".format(num_g))
    first_hit, ap, num_rel = E.rank_eval(dist_dev, q_pids, g_pids, q_camids, g_camids, junk)
    return E.reduce_cmc_map(first_hit.cpu().numpy(), ap.cpu().numpy(), num_rel.cpu().numpy(), max_rank, num_g, denominators)


def eval_func(distmat, q_pids, g_pids, q_camids, g_camids, max_rank=50, *, junk=None):
    """utils/metrics.py:28-88 -> (all_cmc float32[max_rank], mAP float64).

    distmat: numpy [Q, G] (as the reference passes), a torch tensor on any device, or a LazyDistmat.
    """
    dev = _device()
    if isinstance(distmat, LazyDistmat):
        d = distmat.device_tensor
    else:
        d = distmat if isinstance(distmat, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(distmat))
        if d.dtype != torch.float32:
            d = d.float()  # the reference sorts whatever dtype it is given; the evaluator only ever passes fp32
        d = d.to(dev, non_blocking=True)
        if d.stride(1) != 1:
            d = d.contiguous()
    return _eval_device(d, q_pids, g_pids, q_camids, g_camids, max_rank, junk)


def clipstyle_eval(distmat, q_pids, g_pids, q_camids, g_camids):
    """processor/processor_uniprompt_stage2.py:471-509: junk rule always on, float64 CMC over all
    queries, mAP averaged over ALL queries.  Returns (all_cmc[:50] float64, mAP)."""
    dev = _device()
    d = distmat.device_tensor if isinstance(distmat, LazyDistmat) else _to_device(distmat, dev)
    return _eval_device(d, q_pids, g_pids, q_camids, g_camids, min(50, d.shape[1]), "pid_cam", "all")


class R1_mAP_eval():
    """utils/metrics.py:91-134.  Features stay on the GPU between update() and compute()."""

    def __init__(self, num_query, max_rank=50, feat_norm=True, reranking=False, *, precision=None, junk=None,
                 metric="sqeuclid"):
        super(R1_mAP_eval, self).__init__()
        self.num_query = num_query
        self.max_rank = max_rank
        self.feat_norm = feat_norm
        self.reranking = reranking
        self._precision = precision
        self._junk = junk
        self._metric = metric

    def reset(self):
        self.feats = []
        self.pids = []
        self.camids = []

    def update(self, output):  # called once for each batch
        feat, pid, camid = output
        feats = self.feats  # AttributeError before reset(), exactly like the reference (utils/metrics.py:99-106)
        dev = _device()
        if not isinstance(feat, torch.Tensor):
            feat = torch.as_tensor(np.asarray(feat))
        # the reference does feat.cpu() here (a synchronising D2H per batch, utils/metrics.py:106)
        feats.append(feat.detach().to(dev, dtype=torch.float32, non_blocking=True))
        self.pids.extend(np.asarray(pid))
        self.camids.extend(np.asarray(camid))

    def compute(self):  # called after each epoch
        feats = torch.cat(self.feats, dim=0)
        if self.feat_norm:
            print("The test feature is normalized")
        prep = E.prep_rows(feats, normalize=bool(self.feat_norm), precision=self._precision, keep_xn=True)
        nq = self.num_query
        q, g = prep.rows(0, nq), prep.rows(nq, prep.n)
        q_pids = np.asarray(self.pids[:nq])
        q_camids = np.asarray(self.camids[:nq])
        g_pids = np.asarray(self.pids[nq:])
        g_camids = np.asarray(self.camids[nq:])
        if self.reranking:
            print('=> Enter reranking')
            dist = _rerank_device(prep, nq, k1=50, k2=15, lambda_value=0.3, precision=self._precision)  # utils/metrics.py:127
        else:
            print('=> Computing DistMat with euclidean_distance')
            dist = E.dist_matrix(q, g, self._metric, self._precision)
        cmc, mAP = _eval_device(dist, q_pids, g_pids, q_camids, g_camids, 50, self._junk)  # :132 (max_rank is not forwarded)
        return cmc, mAP, LazyDistmat(dist), self.pids, self.camids, q.xn, g.xn
