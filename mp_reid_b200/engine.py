"""Device-side engine: thin, typed wrappers over the C ABI working on torch CUDA tensors.

torch is used for device memory, streams and (elsewhere) torch.distributed only.  Every function
launches on torch's current stream and never synchronises unless it says so.
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib as L


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("mp_reid_b200 needs a CUDA device (B200 / sm_100a); there is no CPU fallback")


# Optional stage timeline (bench.py): when `_timeline` is a list, mark(name) appends (name, CUDA event recorded on the
# current stream); the duration of a stage is the time between its mark and the previous one.
_timeline = None


def timeline_start():
    global _timeline
    _timeline = []
    mark("start")


def timeline_stop():
    """-> [(stage name, ms)] (synchronises)."""
    global _timeline
    tl, _timeline = _timeline, None
    if not tl:
        return []
    torch.cuda.synchronize()
    return [(tl[i][0], tl[i - 1][1].elapsed_time(tl[i][1])) for i in range(1, len(tl))]


def mark(name: str):
    if _timeline is not None:
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        _timeline.append((name, ev))


def default_precision() -> str:
    return os.environ.get("MPREID_PRECISION", "3xfp16").lower()


def default_junk() -> str:
    return os.environ.get("MPREID_JUNK", "none").lower()


@dataclass
class Prepared:
    """Feature rows ready for the distance kernels (all on one device)."""
    n: int
    D: int
    Dp: int
    xn: torch.Tensor | None      # [n, D] fp32 (normalised or raw copy)
    sqnorm: torch.Tensor         # [n]
    norm: torch.Tensor           # [n]
    hi: torch.Tensor | None      # [n, Dp] fp32 (TF32-exact)
    lo: torch.Tensor | None
    bf: torch.Tensor | None      # [n, Dp] bf16
    hh: torch.Tensor | None = None      # [n, Dp] fp16 hi plane of the 2^s-scaled row
    hl: torch.Tensor | None = None      # [n, Dp] fp16 lo plane
    hscale: torch.Tensor | None = None  # [n] fp32, 2^-s

    def take(self, ids: torch.Tensor) -> "Prepared":
        """Rows `ids` (int64 device tensor) gathered into a new Prepared (used by the row-sharded re-ranking)."""
        s = lambda t: None if t is None else t.index_select(0, ids)
        return Prepared(int(ids.numel()), self.D, self.Dp, s(self.xn), self.sqnorm.index_select(0, ids), self.norm.index_select(0, ids),
                        s(self.hi), s(self.lo), s(self.bf), s(self.hh), s(self.hl), s(self.hscale))

    def rows(self, a: int, b: int) -> "Prepared":
        s = lambda t: None if t is None else t[a:b]
        return Prepared(b - a, self.D, self.Dp, s(self.xn), self.sqnorm[a:b], self.norm[a:b], s(self.hi), s(self.lo), s(self.bf),
                        s(self.hh), s(self.hl), s(self.hscale))


def prep_rows(x: torch.Tensor, normalize: bool, precision: str | None = None, keep_xn: bool = True,
              xn_out: torch.Tensor | None = None) -> Prepared:
    """F.normalize + squared norms + operand planes in one pass (utils/metrics.py:10-11,114).
    `xn_out` ([n, D] fp32 rows, unit column stride) receives the (normalised) rows instead of a new tensor."""
    require_cuda()
    lib = L.load()
    precision = (precision or default_precision()).lower()
    prec = L.PRECISIONS[precision]
    assert x.is_cuda and x.dim() == 2
    if x.dtype != torch.float32:
        x = x.float()
    if x.stride(1) != 1:
        x = x.contiguous()
    n, D = x.shape
    dev = x.device
    pad = 64 if prec in (L.BF16, L.X3FP16, L.X2FP16) else 32
    Dp = (D + pad - 1) // pad * pad
    need_xn = keep_xn or prec == L.FP32_SIMT
    if xn_out is not None:
        assert xn_out.is_cuda and xn_out.dtype == torch.float32 and xn_out.shape == (n, D) and xn_out.stride(1) == 1
        xn = xn_out
    else:
        xn = torch.empty((n, D), dtype=torch.float32, device=dev) if need_xn else None
    sqnorm = torch.empty((n,), dtype=torch.float32, device=dev)
    norm = torch.empty((n,), dtype=torch.float32, device=dev)
    hi = lo = bf = hh = hl = hscale = None
    if prec in (L.X3FP16, L.X2FP16):
        hh = torch.empty((n, Dp), dtype=torch.float16, device=dev)
        hl = torch.empty((n, Dp), dtype=torch.float16, device=dev)
        hscale = torch.empty((n,), dtype=torch.float32, device=dev)
    elif prec == L.X3TF32:
        hi = torch.empty((n, Dp), dtype=torch.float32, device=dev)
        lo = torch.empty((n, Dp), dtype=torch.float32, device=dev)
    elif prec == L.BF16:
        bf = torch.empty((n, Dp), dtype=torch.bfloat16, device=dev)
    with torch.cuda.device(dev):
        L.check(lib.mpreid_prep_rows(x.data_ptr(), n, D, x.stride(0), int(bool(normalize)), _ptr(xn), xn.stride(0) if xn is not None else D, sqnorm.data_ptr(),
                                     norm.data_ptr(), _ptr(hi), _ptr(lo), _ptr(bf), _ptr(hh), _ptr(hl), _ptr(hscale), Dp,
                                     _stream()), "prep_rows")
    return Prepared(n, D, Dp, xn, sqnorm, norm, hi, lo, bf, hh, hl, hscale)


class GalleryPlanes:
    """Operand planes, norms and normalised rows of a whole gallery, filled block by block as the rows arrive, each block
    contracted against the prepared queries right away: prep_rows + dist_matrix on slices of buffers allocated ONCE.
    A block costs two C calls and no allocation (~30 us of host time instead of ~150 us through prep_rows / dist_matrix),
    which is what lets the sharded evaluator work in small sub-blocks (section 5 of DESIGN.md)."""

    def __init__(self, num_g: int, D: int, device, precision: str | None = None, keep_xn: bool = True):
        require_cuda()
        self.lib = L.load()
        self.precision = (precision or default_precision()).lower()
        self.prec = L.PRECISIONS[self.precision]
        if self.prec == L.FP32_SIMT:
            raise ValueError("GalleryPlanes needs a tensor-core precision mode")
        self.n, self.D, self.dev = num_g, D, device
        pad = 64 if self.prec in (L.BF16, L.X3FP16, L.X2FP16) else 32
        self.Dp = (D + pad - 1) // pad * pad
        e = torch.empty
        self.xn = e((num_g, D), dtype=torch.float32, device=device) if keep_xn else None
        self.sqnorm = e((num_g,), dtype=torch.float32, device=device)
        self.norm = e((num_g,), dtype=torch.float32, device=device)
        self.hi = self.lo = self.bf = self.hh = self.hl = self.hscale = None
        if self.prec in (L.X3FP16, L.X2FP16):
            self.hh = e((num_g, self.Dp), dtype=torch.float16, device=device)
            self.hl = e((num_g, self.Dp), dtype=torch.float16, device=device)
            self.hscale = e((num_g,), dtype=torch.float32, device=device)
            self.a, self.b, self.esz = self.hh, self.hl, 2
        elif self.prec == L.X3TF32:
            self.hi = e((num_g, self.Dp), dtype=torch.float32, device=device)
            self.lo = e((num_g, self.Dp), dtype=torch.float32, device=device)
            self.a, self.b, self.esz = self.hi, self.lo, 4
        else:
            self.bf = e((num_g, self.Dp), dtype=torch.bfloat16, device=device)
            self.a, self.b, self.esz = self.bf, None, 2
        self._p = {k: (None if t is None else t.data_ptr()) for k, t in dict(xn=self.xn, sq=self.sqnorm, nm=self.norm, hi=self.hi, lo=self.lo,
                                                                             bf=self.bf, hh=self.hh, hl=self.hl, sc=self.hscale).items()}

    def add_block(self, x: torch.Tensor, lo: int, normalize: bool, q: Prepared, metric: str, dmat: torch.Tensor):
        """Rows x [n, D] (fp32, unit column stride, on the device) are gallery rows lo .. lo+n: prepare them in place and write
        the distances of all queries to them into dmat[:, lo:lo+n]."""
        n = x.shape[0]
        with torch.cuda.device(self.dev):
            self._add_block(x, n, lo, normalize, q, metric, dmat)

    def _add_block(self, x, n, lo, normalize, q, metric, dmat):
        p, Dp, D, st = self._p, self.Dp, self.D, _stream()
        off = lambda base, per_row: None if base is None else base + lo * per_row
        met = L.METRICS[metric]
        L.check(self.lib.mpreid_prep_rows(x.data_ptr(), n, D, x.stride(0), int(bool(normalize)), off(p["xn"], 4 * D), D, off(p["sq"], 4), off(p["nm"], 4),
                                          off(p["hi"], 4 * Dp), off(p["lo"], 4 * Dp), off(p["bf"], 2 * Dp), off(p["hh"], 2 * Dp), off(p["hl"], 2 * Dp),
                                          off(p["sc"], 4), Dp, st), "prep_rows")
        if met == L.ARCCOS:
            qa, ga = q.norm.data_ptr(), off(p["nm"], 4)
        elif met in (L.ONE_MINUS_DOT, L.DOT):
            qa = ga = None
        else:
            qa, ga = q.sqnorm.data_ptr(), off(p["sq"], 4)
        if self.prec == L.X3TF32:
            qA, qB = q.hi, q.lo
        elif self.prec in (L.X3FP16, L.X2FP16):
            qA, qB = q.hh, q.hl
        else:
            qA, qB = q.bf, None
        L.check(self.lib.mpreid_dist_matrix(qA.data_ptr(), _ptr(qB), self.a.data_ptr() + lo * Dp * self.esz,
                                            None if self.b is None else self.b.data_ptr() + lo * Dp * self.esz, qa, ga, _ptr(q.hscale), off(p["sc"], 4),
                                            q.n, n, Dp, Dp, met, self.prec, dmat.data_ptr() + 4 * lo, dmat.stride(0), None, st), "dist_matrix")


def alloc_dist(Q: int, G: int, device) -> torch.Tensor:
    """[Q, G] fp32 view over a buffer whose leading dimension is padded to 128 B (vector stores, TMA)."""
    ld = (G + 31) // 32 * 32
    return torch.empty((Q, ld), dtype=torch.float32, device=device)[:, :G]


def dist_matrix(q: Prepared, g: Prepared, metric: str = "sqeuclid", precision: str | None = None,
                out: torch.Tensor | None = None, row_max: torch.Tensor | None = None) -> torch.Tensor:
    """utils/metrics.py:7-25, processor_uniprompt_stage2.py:466-468, utils/reranking.py:36-41."""
    require_cuda()
    lib = L.load()
    precision = (precision or default_precision()).lower()
    prec, met = L.PRECISIONS[precision], L.METRICS[metric]
    dev = q.sqnorm.device
    if out is None:
        out = alloc_dist(q.n, g.n, dev)
    assert out.shape == (q.n, g.n) and out.stride(1) == 1 and out.dtype == torch.float32
    if met == L.ARCCOS:
        qa, ga = q.norm, g.norm
    elif met in (L.ONE_MINUS_DOT, L.DOT):
        qa = ga = None
    else:
        qa, ga = q.sqnorm, g.sqnorm
    if prec == L.FP32_SIMT:
        if q.xn is None or g.xn is None or g.xn.stride(0) != q.xn.stride(0):
            raise ValueError("SIMT path needs the fp32 rows of both sides with equal leading dimensions")
        a, b, c, d, K, ldk = q.xn, None, g.xn, None, q.D, q.xn.stride(0)
    elif prec == L.X3TF32:
        a, b, c, d, K, ldk = q.hi, q.lo, g.hi, g.lo, q.Dp, q.Dp
    elif prec in (L.X3FP16, L.X2FP16):
        a, b, c, d, K, ldk = q.hh, q.hl, g.hh, g.hl, q.Dp, q.Dp
    else:
        a, b, c, d, K, ldk = q.bf, None, g.bf, None, q.Dp, q.Dp
    if a is None or c is None:
        raise ValueError(f"features were not prepared for precision '{precision}'")
    if row_max is not None:
        row_max.fill_(float("-inf"))
    with torch.cuda.device(dev):
        L.check(lib.mpreid_dist_matrix(_ptr(a), _ptr(b), _ptr(c), _ptr(d), _ptr(qa), _ptr(ga), _ptr(q.hscale), _ptr(g.hscale),
                                       q.n, g.n, K, ldk, met, prec,
                                       out.data_ptr(), out.stride(0), _ptr(row_max), _stream()), "dist_matrix")
    return out


def dist_matrix_all_pairs(x: Prepared, precision: str | None = None, out: torch.Tensor | None = None,
                          row_max: torch.Tensor | None = None, metric: str = "sqeuclid") -> torch.Tensor:
    """utils/reranking.py:36-41: the stacked features against themselves; upper-triangle tiles + mirrored stores."""
    require_cuda()
    lib = L.load()
    precision = (precision or default_precision()).lower()
    prec, met = L.PRECISIONS[precision], L.METRICS[metric]
    dev = x.sqnorm.device
    if out is None:
        out = alloc_dist(x.n, x.n, dev)
    assert out.shape == (x.n, x.n) and out.stride(1) == 1 and out.dtype == torch.float32
    aux = x.norm if met == L.ARCCOS else (None if met in (L.ONE_MINUS_DOT, L.DOT) else x.sqnorm)
    if prec == L.FP32_SIMT:
        a, b, K, ldk = x.xn, None, x.D, x.xn.stride(0)
    elif prec == L.X3TF32:
        a, b, K, ldk = x.hi, x.lo, x.Dp, x.Dp
    elif prec in (L.X3FP16, L.X2FP16):
        a, b, K, ldk = x.hh, x.hl, x.Dp, x.Dp
    else:
        a, b, K, ldk = x.bf, None, x.Dp, x.Dp
    if a is None:
        raise ValueError(f"features were not prepared for precision '{precision}'")
    if row_max is not None:
        row_max.fill_(float("-inf"))
    with torch.cuda.device(dev):
        L.check(lib.mpreid_dist_matrix_symmetric(_ptr(a), _ptr(b), _ptr(aux), _ptr(x.hscale), x.n, K, ldk, met, prec,
                                                 out.data_ptr(), out.stride(0), _ptr(row_max), _stream()), "dist_matrix_symmetric")
    return out


RANK_EVAL_LAUNCHES = 8   # kernels one mpreid_rank_eval call launches (6 label-index kernels, rank_count, ap_finalize)
_RESERVED_LABEL = np.iinfo(np.int64).min   # the empty-slot marker of the label hash table (include/mpreid_b200.h)


def _labels(x, device) -> torch.Tensor:
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=torch.int64, non_blocking=True)
    a = np.ascontiguousarray(np.asarray(x), dtype=np.int64)
    if a.size and a.min() == _RESERVED_LABEL:
        raise ValueError("pid / camid value INT64_MIN is reserved")
    return torch.from_numpy(a).to(device, non_blocking=True)


_pos_capacity_hint = {}


class RankResult:
    """Per-query outputs of the rank/AP kernels in ONE device allocation (so that one D2H copy brings
    everything back, status word included):  [ap f64 x Q | first_hit i32 x Q | num_rel i32 x Q | status i32 x 4]."""

    def __init__(self, Q: int, device):
        self.Q = Q
        self.buf = torch.empty((16 * Q + 16,), dtype=torch.uint8, device=device)
        self.ap = self.buf[: 8 * Q].view(torch.float64)
        self.first_hit = self.buf[8 * Q: 12 * Q].view(torch.int32)
        self.num_rel = self.buf[12 * Q: 16 * Q].view(torch.int32)
        self.status = self.buf[16 * Q:].view(torch.int32)

    def to_host(self):
        """-> (first_hit, ap, num_rel, status) numpy views of one pinned host copy (synchronises)."""
        host = torch.empty(self.buf.shape, dtype=torch.uint8, pin_memory=True)
        host.copy_(self.buf, non_blocking=True)
        torch.cuda.current_stream(self.buf.device).synchronize()   # the copy runs on the stream of the buffer's device
        h = host.numpy()
        Q = self.Q
        return (h[8 * Q: 12 * Q].view(np.int32), h[: 8 * Q].view(np.float64), h[12 * Q: 16 * Q].view(np.int32),
                h[16 * Q:].view(np.int32))


def rank_eval_async(dist: torch.Tensor, q_pid, g_pid, q_cam=None, g_cam=None, junk: str | None = None,
                    capacity: int | None = None) -> RankResult:
    """Launches the rank/AP kernels and returns without synchronising; check `status[0]` after the copy."""
    require_cuda()
    lib = L.load()
    junk_mode = L.JUNKS[(junk or default_junk()).lower()]
    assert dist.is_cuda and dist.dtype == torch.float32 and dist.stride(1) == 1
    Q, G = dist.shape
    dev = dist.device
    q_pid, g_pid = _labels(q_pid, dev), _labels(g_pid, dev)
    if junk_mode != L.JUNK_NONE:
        q_cam, g_cam = _labels(q_cam, dev), _labels(g_cam, dev)
    else:
        q_cam = g_cam = None
    res = RankResult(Q, dev)
    cap = capacity or max(_pos_capacity_hint.get((Q, G), 0), 64 * Q, 1 << 16)
    nbytes = lib.mpreid_rank_eval_workspace_bytes(Q, G, cap)
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        L.check(lib.mpreid_rank_eval(dist.data_ptr(), dist.stride(0), Q, G, q_pid.data_ptr(), g_pid.data_ptr(),
                                     _ptr(q_cam), _ptr(g_cam), junk_mode, res.first_hit.data_ptr(), res.ap.data_ptr(),
                                     res.num_rel.data_ptr(), ws.data_ptr(), nbytes, cap, res.status.data_ptr(), _stream()),
                "rank_eval")
    return res


def rank_eval_host(dist: torch.Tensor, q_pid, g_pid, q_cam=None, g_cam=None, junk: str | None = None):
    """eval_func's per-query part (utils/metrics.py:39-80) -> numpy (first_hit i32[Q], ap f64[Q], num_rel i32[Q]).
    One synchronising D2H copy; re-run with the exact workspace size if the positives workspace overflowed."""
    res = rank_eval_async(dist, q_pid, g_pid, q_cam, g_cam, junk)
    fh, ap, nr, st = res.to_host()
    if int(st[0]) != 0:
        need = int(st[1])
        if need >= 2**31 - 1:
            raise L.MpreidError("rank_eval: more than 2^31 same-pid (query, gallery) pairs")
        _pos_capacity_hint[tuple(dist.shape)] = need
        res = rank_eval_async(dist, q_pid, g_pid, q_cam, g_cam, junk, capacity=need)
        fh, ap, nr, st = res.to_host()
        if int(st[0]) != 0:
            raise L.MpreidError("rank_eval: positives workspace overflow after resize")
    return fh, ap, nr


def rank_eval(dist: torch.Tensor, q_pid, g_pid, q_cam=None, g_cam=None, junk: str | None = None):
    """Device-tensor flavour: -> (first_hit i32[Q], ap f64[Q], num_rel i32[Q]) on the device.
    Synchronises once on the 16-byte status word (workspace overflow -> re-run with the exact size)."""
    res = rank_eval_async(dist, q_pid, g_pid, q_cam, g_cam, junk)
    st = res.status.cpu()
    if int(st[0]) != 0:
        need = int(st[1])
        if need >= 2**31 - 1:
            raise L.MpreidError("rank_eval: more than 2^31 same-pid (query, gallery) pairs")
        _pos_capacity_hint[tuple(dist.shape)] = need
        res = rank_eval_async(dist, q_pid, g_pid, q_cam, g_cam, junk, capacity=need)
        if int(res.status.cpu()[0]) != 0:
            raise L.MpreidError("rank_eval: positives workspace overflow after resize")
    return res.first_hit, res.ap, res.num_rel


def reduce_cmc_map(first_hit, ap, num_rel, max_rank: int, num_g: int, denominators: str = "valid"):
    """utils/metrics.py:82-86 on the gathered per-query values, computed by numpy itself.

    denominators='valid' is eval_func; 'all' is the inline CLIP-style loop
    (processor/processor_uniprompt_stage2.py:508-509: mean over all queries, float64 CMC).
    """
    first_hit = np.asarray(first_hit)
    ap = np.asarray(ap, dtype=np.float64)
    num_rel = np.asarray(num_rel)
    valid = num_rel > 0
    n_valid = float(valid.sum())
    if denominators == "valid":   # utils/metrics.py:82; the inline CLIP-style loop has no such check (mAP = 0, CMC = 0)
        assert n_valid > 0, "Error: all query identities do not appear in gallery"
    fh = first_hit[valid].astype(np.int64)
    counts = np.bincount(np.minimum(fh, max_rank + 1), minlength=max_rank + 2)[1:max_rank + 1].cumsum()
    if denominators == "valid":
        cmc = counts.astype(np.float32) / n_valid
        mAP = np.mean(ap[valid])
    else:
        cmc = counts.astype(np.float64) / len(first_hit)
        mAP = np.where(valid, ap, 0.0).mean()
    return cmc, mAP


def row_topk(dist: torch.Tensor, k: int, row_scale: torch.Tensor | None = None, want_values: bool = False):
    require_cuda()
    lib = L.load()
    Q, G = dist.shape
    idx = torch.empty((Q, k), dtype=torch.int32, device=dist.device)
    val = torch.empty((Q, k), dtype=torch.float32, device=dist.device) if want_values else None
    with torch.cuda.device(dist.device):
        L.check(lib.mpreid_row_topk(dist.data_ptr(), dist.stride(0), Q, G, k, _ptr(row_scale), idx.data_ptr(), _ptr(val), _stream()),
                "row_topk")
    return (idx, val) if want_values else idx


def row_kth(dist: torch.Tensor, t: int, bound: bool = False) -> torch.Tensor:
    """The t-th smallest value (1-based) of every row of a short-row matrix (at most 4,096 columns); bound=True (t <= 128):
    a cheap upper bound of it instead (the t-th smallest of a subset of the row) -- all a threshold needs."""
    require_cuda()
    lib = L.load()
    R, S = dist.shape
    out = torch.empty((R,), dtype=torch.float32, device=dist.device)
    fn = lib.mpreid_row_kth_bound if (bound and t <= 128) else lib.mpreid_row_kth
    with torch.cuda.device(dist.device):
        L.check(fn(dist.data_ptr(), dist.stride(0), R, S, int(t), out.data_ptr(), _stream()), "row_kth")
    return out


def row_max(dist: torch.Tensor) -> torch.Tensor:
    require_cuda()
    lib = L.load()
    Q, G = dist.shape
    out = torch.empty((Q,), dtype=torch.float32, device=dist.device)
    with torch.cuda.device(dist.device):
        L.check(lib.mpreid_row_max(dist.data_ptr(), dist.stride(0), Q, G, out.data_ptr(), _stream()), "row_max")
    return out


def rerank_from_dist(dist_all: torch.Tensor, query_num: int, k1: int, k2: int, lambda_value: float,
                     out: torch.Tensor | None = None, row_max: torch.Tensor | None = None) -> torch.Tensor:
    """utils/reranking.py:45-99 on an all-pairs matrix in the orientation dist_all[i][j] = distmat[j][i]."""
    require_cuda()
    lib = L.load()
    N = dist_all.shape[0]
    assert dist_all.shape == (N, N) and dist_all.stride(1) == 1 and dist_all.dtype == torch.float32
    dev = dist_all.device
    G = N - query_num
    if out is None:
        out = alloc_dist(query_num, G, dev)
    nbytes = lib.mpreid_rerank_workspace_bytes(N, query_num, k1, k2)
    if nbytes == 0:
        raise ValueError(f"re_ranking: unsupported arguments N={N} Q={query_num} k1={k1} k2={k2}")
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        L.check(lib.mpreid_rerank(dist_all.data_ptr(), dist_all.stride(0), _ptr(row_max), N, query_num, k1, k2, float(lambda_value),
                                  out.data_ptr(), out.stride(0), ws.data_ptr(), nbytes, None, _stream()), "rerank")
    return out


# ------------------------------------------------------------------------------------------------
# staged re-ranking (row-sharded multi-GPU runs; the monolithic call above chains the same stages)
def rerank_neighbor_count(k1: int, k2: int) -> int:
    return int(L.load().mpreid_rerank_neighbor_count(k1, k2))


def rerank_build_v0(dist_rows: torch.Tensor, row_ids: torch.Tensor | None, N: int, k1: int, nbr_all: torch.Tensor,
                    row_max_rows: torch.Tensor):
    """utils/reranking.py:51-71 for the rows of one block of the all-pairs matrix -> (col i32, val fp16 bits, len) ELL."""
    require_cuda()
    lib = L.load()
    R = dist_rows.shape[0]
    C0 = int(lib.mpreid_rerank_v0_capacity(k1, N))
    if C0 == 0:
        raise ValueError(f"re_ranking: unsupported k1={k1}")
    dev = dist_rows.device
    v0_col = torch.empty((R, C0), dtype=torch.int32, device=dev)
    v0_val = torch.empty((R, C0), dtype=torch.float16, device=dev)  # fp16 V entries (NCCL has no int16)
    v0_len = torch.empty((R,), dtype=torch.int32, device=dev)
    assert nbr_all.dtype == torch.int32 and nbr_all.is_contiguous() and nbr_all.shape[0] == N
    with torch.cuda.device(dev):
        L.check(lib.mpreid_rerank_build_v0(dist_rows.data_ptr(), dist_rows.stride(0), _ptr(row_ids), R, N, k1, nbr_all.data_ptr(),
                                           nbr_all.shape[1], row_max_rows.data_ptr(), v0_col.data_ptr(), v0_val.data_ptr(),
                                           v0_len.data_ptr(), _stream()), "rerank_build_v0")
    return v0_col, v0_val, v0_len


# ------------------------------------------------------------------------------------------------
# fused all-pairs pass: top-(k1+1) candidates from the GEMM epilogue, the N x N matrix is never written
def _operands(x: Prepared, prec: int):
    if prec == L.X3TF32:
        return x.hi, x.lo
    if prec in (L.X3FP16, L.X2FP16):
        return x.hh, x.hl
    return x.bf, None


def dist_symmetric_topk(x: Prepared, thr: torch.Tensor, cand_cap: int, query_num: int, precision: str | None = None,
                        own_mod: int = 1, own_rank: int = 0):
    """utils/reranking.py:36-48 without the matrix: -> (cand int64 [N, cap], cand_cnt int32 [N], block fp32 [Q, G] view,
    col0 (column of gallery sample 0 in the block's buffer), row_max [N]).
    own_mod / own_rank: contract only the tiles of the 256-row blocks p with p % own_mod == own_rank (row-sharded runs)."""
    require_cuda()
    lib = L.load()
    prec = L.PRECISIONS[(precision or default_precision()).lower()]
    if prec == L.FP32_SIMT:
        raise ValueError("the fused all-pairs pass needs a tensor-core precision mode")
    a, b = _operands(x, prec)
    if a is None:
        raise ValueError("features were not prepared for this precision")
    N, dev = x.n, x.sqnorm.device
    G = N - query_num
    lead = query_num & 31
    ld = (lead + G + 31) // 32 * 32
    block = torch.empty((query_num, ld), dtype=torch.float32, device=dev)
    cand = torch.empty((N, cand_cap), dtype=torch.int64, device=dev)
    cnt = torch.zeros((N,), dtype=torch.int32, device=dev)
    row_max = torch.full((N,), float("-inf"), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        L.check(lib.mpreid_dist_symmetric_topk(_ptr(a), _ptr(b), x.sqnorm.data_ptr(), _ptr(x.hscale), N, x.Dp, x.Dp, prec,
                                               thr.data_ptr(), cand.data_ptr(), cnt.data_ptr(), cand_cap, query_num,
                                               block.data_ptr(), ld, row_max.data_ptr(), own_mod, own_rank, _stream()), "dist_symmetric_topk")
    return cand, cnt, block, lead, row_max


def cand_topk(cand: torch.Tensor, cnt: torch.Tensor, k: int, row_scale: torch.Tensor | None, thr: torch.Tensor, partial: bool = False):
    """-> (idx int32 [N, k], val fp32 [N, k] (divided values), status int32[4] on the device); partial=True (one rank's
    share of every row): -> (keys int64 [N, k], status)."""
    lib = L.load()
    N, cap = cand.shape
    dev = cand.device
    status = torch.empty((4,), dtype=torch.int32, device=dev)
    if partial:
        keys = torch.empty((N, k), dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            L.check(lib.mpreid_cand_topk(cand.data_ptr(), cnt.data_ptr(), cap, N, k, _ptr(row_scale), thr.data_ptr(), None, None,
                                         keys.data_ptr(), status.data_ptr(), _stream()), "cand_topk")
        return keys, status
    idx = torch.empty((N, k), dtype=torch.int32, device=dev)
    val = torch.empty((N, k), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        L.check(lib.mpreid_cand_topk(cand.data_ptr(), cnt.data_ptr(), cap, N, k, _ptr(row_scale), thr.data_ptr(), idx.data_ptr(),
                                     val.data_ptr(), None, status.data_ptr(), _stream()), "cand_topk")
    return idx, val, status


def merge_topk(keys_all: torch.Tensor, row_scale: torch.Tensor | None, thr: torch.Tensor):
    """keys_all int64 [P, N, k] (every rank's partial selection) -> (idx int32 [N, k], val fp32 [N, k], status int32[4])."""
    lib = L.load()
    P, N, k = keys_all.shape
    dev = keys_all.device
    idx = torch.empty((N, k), dtype=torch.int32, device=dev)
    val = torch.empty((N, k), dtype=torch.float32, device=dev)
    status = torch.empty((4,), dtype=torch.int32, device=dev)
    assert keys_all.is_contiguous()
    with torch.cuda.device(dev):
        L.check(lib.mpreid_merge_topk(keys_all.data_ptr(), P, N, k, _ptr(row_scale), thr.data_ptr(), idx.data_ptr(), val.data_ptr(),
                                      status.data_ptr(), _stream()), "merge_topk")
    return idx, val, status


def rerank_build_v0_sparse(row_ids: torch.Tensor | None, R: int, N: int, k1: int, nbr_all: torch.Tensor, nbr_val_all: torch.Tensor,
                           row_max_rows: torch.Tensor, xn: torch.Tensor, sqnorm: torch.Tensor):
    """utils/reranking.py:51-71 from neighbour lists + values and the feature rows (no matrix rows)."""
    lib = L.load()
    C0 = int(lib.mpreid_rerank_v0_capacity(k1, N))
    if C0 == 0:
        raise ValueError(f"re_ranking: unsupported k1={k1}")
    dev = nbr_all.device
    v0_col = torch.empty((R, C0), dtype=torch.int32, device=dev)
    v0_val = torch.empty((R, C0), dtype=torch.float16, device=dev)
    v0_len = torch.empty((R,), dtype=torch.int32, device=dev)
    assert nbr_all.is_contiguous() and nbr_val_all.is_contiguous() and xn.stride(1) == 1 and xn.dtype == torch.float32
    with torch.cuda.device(dev):
        L.check(lib.mpreid_rerank_build_v0_sparse(_ptr(row_ids), R, N, k1, nbr_all.data_ptr(), nbr_val_all.data_ptr(), nbr_all.shape[1],
                                                  row_max_rows.data_ptr(), xn.data_ptr(), xn.stride(0), xn.shape[1], sqnorm.data_ptr(),
                                                  v0_col.data_ptr(), v0_val.data_ptr(), v0_len.data_ptr(), _stream()),
                "rerank_build_v0_sparse")
    return v0_col, v0_val, v0_len


def rerank_finish_workspace(N: int, Q: int, k1: int, k2: int, device) -> torch.Tensor:
    nbytes = L.load().mpreid_rerank_finish_workspace_bytes(N, Q, k1, k2)
    if nbytes == 0:
        raise ValueError(f"re_ranking: unsupported arguments N={N} Q={Q} k1={k1} k2={k2}")
    return torch.empty((nbytes,), dtype=torch.uint8, device=device)


def rerank_finish_v_views(ws: torch.Tensor, N: int, Q: int, k1: int, k2: int):
    """Tensor views of the expanded V rows inside a finish workspace -> (v_col int32 [N, C1], v_val fp16 [N, C1], v_len int32 [N])."""
    import ctypes
    out = (ctypes.c_int64 * 4)()
    L.check(L.load().mpreid_rerank_finish_layout(N, Q, k1, k2, out), "rerank_finish_layout")
    o_col, o_val, o_len, C1 = [int(x) for x in out]
    v_col = ws[o_col: o_col + N * C1 * 4].view(torch.int32).view(N, C1)
    v_val = ws[o_val: o_val + N * C1 * 2].view(torch.float16).view(N, C1)
    v_len = ws[o_len: o_len + N * 4].view(torch.int32)
    return v_col, v_val, v_len


def rerank_finish(nbr_all: torch.Tensor, v0, dist_qrows: torch.Tensor, q_ids: torch.Tensor | None, row_max_q: torch.Tensor,
                  N: int, Q: int, k1: int, k2: int, lambda_value: float, out: torch.Tensor | None = None,
                  block_col0: int | None = None, rows_global: bool = False, stages: int = 7, ws: torch.Tensor | None = None,
                  qe_rows: tuple[int, int] = (0, 0)) -> torch.Tensor:
    """utils/reranking.py:73-99 for the query rows `dist_qrows` [Qs, N] -> final [Qs, N-Q].
    block_col0: `dist_qrows` is instead the [Qs, >= col0 + G] buffer of query-to-gallery distances, gallery sample 0 at
    column block_col0 (what the fused all-pairs pass keeps).  rows_global: dist_qrows / row_max_q are the full [Q, .] block
    and [N] maxima, addressed by the global indices q_ids (row-sharded runs).  The V0 arrays may be trimmed to any width."""
    require_cuda()
    lib = L.load()
    v0_col, v0_val, v0_len = v0
    Qs = int(q_ids.numel()) if rows_global else dist_qrows.shape[0]
    dev = dist_qrows.device
    if out is None:
        out = alloc_dist(Qs, N - Q, dev)
    nbytes = lib.mpreid_rerank_finish_workspace_bytes(N, Q, k1, k2)
    if nbytes == 0:
        raise ValueError(f"re_ranking: unsupported arguments N={N} Q={Q} k1={k1} k2={k2}")
    if ws is None:
        ws = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
    assert v0_col.is_contiguous() and v0_val.is_contiguous() and v0_col.shape[0] == N and v0_col.shape == v0_val.shape
    col0 = Q if block_col0 is None else int(block_col0)
    with torch.cuda.device(dev):
        parts = [stages] if _timeline is None else [p for p in (stages & 25, stages & 6) if p]   # timeline mode: separately timed calls
        for part in parts:
            L.check(lib.mpreid_rerank_finish_ex(nbr_all.data_ptr(), nbr_all.shape[1], v0_col.data_ptr(), v0_val.data_ptr(), v0_len.data_ptr(),
                                                dist_qrows.data_ptr(), dist_qrows.stride(0), col0, _ptr(q_ids), row_max_q.data_ptr(), N, Q, Qs,
                                                k1, k2, float(lambda_value), out.data_ptr(), out.stride(0), ws.data_ptr(), nbytes, part,
                                                v0_col.shape[1], int(bool(rows_global)), int(qe_rows[0]), int(qe_rows[1]), _stream()),
                    "rerank_finish")
            if part & 25:
                mark("rerank.expand_index")
            elif part & 2:
                mark("rerank.jaccard_blend")
    return out


def rerank_blend_default(dist_qrows: torch.Tensor, q_ids: torch.Tensor | None, row_max_q: torch.Tensor, N: int, Q: int, lambda_value: float,
                         out: torch.Tensor, block_col0: int | None = None, rows_global: bool = False, ctas_per_sm: int = 0):
    """The dense part of utils/reranking.py:95 (Jaccard distance 1 wherever the query shares no V column with the gallery
    sample): depends only on the distance block and the maxima, so it can run on a side stream while the sparse stages run."""
    lib = L.load()
    Qs = int(q_ids.numel()) if rows_global else dist_qrows.shape[0]
    col0 = Q if block_col0 is None else int(block_col0)
    with torch.cuda.device(out.device):
        L.check(lib.mpreid_rerank_blend_default(dist_qrows.data_ptr(), dist_qrows.stride(0), col0, _ptr(q_ids) if rows_global else None,
                                                row_max_q.data_ptr(), Qs, N - Q, float(lambda_value), out.data_ptr(), out.stride(0),
                                                int(ctas_per_sm), _stream()), "rerank_blend_default")
