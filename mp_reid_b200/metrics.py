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


def _devices():
    """Devices of the single-process multi-GPU mode: MPREID_DEVICES="0,1,2" or "all" (default: the one device).
    The first entry is where update() uploads; the evaluator is driven from ONE host thread, as the reference's
    do_inference is (under DIST_TRAIN only rank 0 evaluates, processor/processor.py:117-118)."""
    spec = os.environ.get("MPREID_DEVICES", "").strip().lower()
    d0 = _device()
    if not spec:
        return [d0]
    ids = list(range(torch.cuda.device_count())) if spec == "all" else [int(x) for x in spec.split(",") if x.strip() != ""]
    # the upload device comes first; an index listed twice gives two shards on that device (how the one-GPU test box
    # exercises this path)
    if d0.index in ids:
        ids.remove(d0.index)
    devs = [d0] + [torch.device(f"cuda:{i}") for i in ids]
    return devs


_COPY_STREAMS = {}
_SIDE_STREAMS = {}


def _side_stream(dev, tag):
    """Long-lived auxiliary streams (peer-to-peer fan-out / receive), one per (device, purpose)."""
    key = (dev.index, tag)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=dev)
    return _SIDE_STREAMS[key]


def _copy_stream(dev):
    """One upload stream per device for the life of the process (a fresh stream per evaluator would
    get a fresh caching-allocator pool, i.e. a cudaMalloc per evaluation)."""
    key = (dev.type, dev.index)
    if key not in _COPY_STREAMS:
        _COPY_STREAMS[key] = torch.cuda.Stream(device=dev)
    return _COPY_STREAMS[key]


def _to_device(x, dev):
    if not isinstance(x, torch.Tensor):
        x = torch.as_tensor(np.asarray(x))
    if x.dtype != torch.float32:
        x = x.float()
    return x.to(dev, non_blocking=True)


class LabelSeq:
    """The accumulated pids / camids.  The reference keeps Python lists of numpy scalars
    (utils/metrics.py:107-108: ``self.pids.extend(np.asarray(pid))``), which costs milliseconds per
    evaluation at MSMT17 size; this sequence stores the per-batch arrays and behaves like that list
    (len, indexing, slicing, iteration, ==, np.asarray) without materialising it."""

    def __init__(self):
        self._chunks = []
        self._flat = None

    def extend(self, values):
        self._chunks.append(np.asarray(values).reshape(-1))
        self._flat = None

    def append(self, value):
        self.extend([value])

    def array(self) -> np.ndarray:
        if self._flat is None:
            self._flat = np.concatenate(self._chunks) if self._chunks else np.zeros((0,), dtype=np.int64)
        return self._flat

    def __len__(self):
        return int(sum(c.shape[0] for c in self._chunks))

    def __getitem__(self, item):
        return self.array()[item]

    def __iter__(self):
        return iter(self.array())

    def __array__(self, dtype=None, copy=None):
        a = self.array()
        return a if dtype is None else a.astype(dtype, copy=False)

    def __eq__(self, other):
        return len(self) == len(other) and bool(np.all(self.array() == np.asarray(other)))

    def __repr__(self):
        return f"LabelSeq({self.array()!r})"


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

    def _row_blocks(self, rows):
        t = self.device_tensor
        for lo in range(0, self.shape[0], rows):
            yield lo, t[lo:lo + rows]

    def save(self, path: str, block_bytes: int = 256 << 20) -> str:
        """Persist the matrix as a .npy file (the reference's TEST.DIST_MAT = "dist_mat.npy", config/defaults.py:327,
        the on-disk format for offline analysis) without a second full-size host copy: the file is memory-mapped and
        filled block by block straight from the device through one pinned staging buffer."""
        if not path.endswith(".npy"):
            path += ".npy"
        out = np.lib.format.open_memmap(path, mode="w+", dtype=np.float32, shape=self.shape)
        if self._host is not None:
            out[:] = self._host
        else:
            rows = max(1, block_bytes // max(4 * self.shape[1], 1))
            stage = torch.empty((rows, self.shape[1]), dtype=torch.float32, pin_memory=True)
            for lo, blk in self._row_blocks(rows):
                n = blk.shape[0]
                stage[:n].copy_(blk, non_blocking=True)
                torch.cuda.current_stream(blk.device).synchronize()
                out[lo:lo + n] = stage[:n].numpy()
        out.flush()
        del out
        return path


def load_distmat(path: str, mmap: bool = True) -> np.ndarray:
    """Reads a matrix written by LazyDistmat.save / np.save (TEST.DIST_MAT)."""
    return np.load(path, mmap_mode="r" if mmap else None)


class LazyDistmatShards(LazyDistmat):
    """The same handle over query-row shards that live on several devices (single-process multi-GPU mode)."""

    def __init__(self, shards):
        self._shards = list(shards)
        self._dev = None
        self._host = None
        self.shape = (sum(t.shape[0] for t in self._shards), self._shards[0].shape[1])
        self.dtype = np.dtype(np.float32)
        self.ndim = 2

    @property
    def device_tensor(self) -> torch.Tensor:
        if self._dev is None:
            d0 = self._shards[0].device
            self._dev = torch.cat([t.to(d0) for t in self._shards], dim=0)
        return self._dev

    def numpy(self) -> np.ndarray:
        if self._host is None:
            self._host = np.concatenate([t.cpu().numpy() for t in self._shards], axis=0)
        return self._host

    def _row_blocks(self, rows):
        base = 0
        for t in self._shards:
            for lo in range(0, t.shape[0], rows):
                yield base + lo, t[lo:lo + rows]
            base += t.shape[0]

    def __repr__(self):
        return f"LazyDistmatShards(shape={self.shape}, devices={[str(t.device) for t in self._shards]})"


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
    first_hit, ap, num_rel = E.rank_eval_host(dist_dev, q_pids, g_pids, q_camids, g_camids, junk)   # one D2H copy
    return E.reduce_cmc_map(first_hit, ap, num_rel, max_rank, num_g, denominators)


def _dense_row_ranks(d: torch.Tensor) -> torch.Tensor:
    """fp32 matrix whose rows order exactly like the rows of `d` (any float dtype): equal values -> equal ranks."""
    assert d.shape[1] < (1 << 24), "dense ranks are exact in float32 only below 2^24 columns"
    vals, idx = torch.sort(d, dim=1, stable=True)
    new = torch.ones(vals.shape, dtype=torch.bool, device=d.device)
    new[:, 1:] = vals[:, 1:] != vals[:, :-1]
    new[:, 1:] &= ~(torch.isnan(vals[:, 1:]) & torch.isnan(vals[:, :-1]))   # numpy: all NaNs tie (and sort last)
    ranks = torch.cumsum(new, dim=1).to(torch.float32)
    return torch.empty_like(ranks).scatter_(1, idx, ranks)


def eval_func(distmat, q_pids, g_pids, q_camids, g_camids, max_rank=50, *, junk=None):
    """utils/metrics.py:28-88 -> (all_cmc float32[max_rank], mAP float64).

    distmat: numpy [Q, G] (as the reference passes), a torch tensor on any device, or a LazyDistmat.
    """
    dev = _device()
    if isinstance(distmat, LazyDistmat):
        d = distmat.device_tensor
    else:
        d = distmat if isinstance(distmat, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(distmat))
        d = d.to(dev, non_blocking=True)
        if d.dtype != torch.float32:
            # The reference sorts whatever dtype it is given; the kernels rank fp32.  Narrower types widen exactly; a
            # float64 matrix whose down-cast is not exact would get new ties, so it is replaced by its per-row dense
            # ranks (exact in fp32 for G < 2^24), which order identically under the stable tie rule.
            d32 = d.float()
            if d.dtype == torch.float64 and not torch.equal(d32.double(), d):
                d32 = _dense_row_ranks(d)
            d = d32
        if d.stride(1) != 1:
            d = d.contiguous()
    return _eval_device(d, q_pids, g_pids, q_camids, g_camids, max_rank, junk)


def clipstyle_eval(distmat, q_pids, g_pids, q_camids, g_camids):
    """processor/processor_uniprompt_stage2.py:471-509: junk rule always on, float64 CMC over all
    queries, mAP averaged over ALL queries.  Returns (all_cmc[:50] float64, mAP)."""
    dev = _device()
    d = distmat.device_tensor if isinstance(distmat, LazyDistmat) else _to_device(distmat, dev)
    return _eval_device(d, q_pids, g_pids, q_camids, g_camids, min(50, d.shape[1]), "pid_cam", "all")


def _fan_out(blk0, devs):
    """One copy of `blk0` (on devs[0]) per device: NCCL broadcast driven from this one process (torch.cuda.comm:
    ring / tree over NVLink instead of P-1 copies out of one GPU), plain peer-to-peer copies as the fallback."""
    if len(devs) == 1:
        return [blk0]
    if len({d.index for d in devs}) < len(devs):   # several shards on one device (tests): nothing to move for those
        return [blk0 if d == devs[0] else blk0.to(d, non_blocking=True) for d in devs]
    try:
        from torch.cuda import comm
        return list(comm.broadcast(blk0, devices=[d.index for d in devs]))
    except Exception:
        return [blk0] + [blk0.to(d, non_blocking=True) for d in devs[1:]]


def _take_rows(pend, rows_out):
    """Split the pending row views into the first `rows_out` rows (adjacent views merged) and the remainder."""
    take, got, rest = [], 0, []
    for t in pend:
        if got + t.shape[0] <= rows_out:
            take.append(t); got += t.shape[0]
        elif got < rows_out:
            take.append(t[: rows_out - got]); rest.append(t[rows_out - got:]); got = rows_out
        else:
            rest.append(t)
    return _merge_adjacent(take), rest


def _merge_adjacent(views):
    """Row views that follow each other in the same allocation (the pieces of one upload) -> one view."""
    out = []
    for v in views:
        if out:
            a = out[-1]
            if (v.dim() == 2 and a.stride() == v.stride() and v.stride(1) == 1 and v.stride(0) == v.shape[1]
                    and a.untyped_storage().data_ptr() == v.untyped_storage().data_ptr()
                    and a.storage_offset() + a.numel() == v.storage_offset()):
                out[-1] = torch.as_strided(a, (a.shape[0] + v.shape[0], a.shape[1]), a.stride(), a.storage_offset())
                continue
        out.append(v)
    return out


class R1_mAP_eval():
    """utils/metrics.py:91-134.  Features stay on the GPU between update() and compute().

    Host batches are uploaded on a side stream as they arrive (update), and compute() consumes the
    gallery in chunks of >= MPREID_CHUNK_ROWS rows: while chunk c is normalised and contracted against
    the queries, chunks c+1.. are still in flight over PCIe.
    """

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
        self.pids = LabelSeq()
        self.camids = LabelSeq()
        self._events = []

    def update(self, output):  # called once for each batch
        feat, pid, camid = output
        feats = self.feats  # AttributeError before reset(), exactly like the reference (utils/metrics.py:99-106)
        dev = _device()
        if not isinstance(feat, torch.Tensor):
            feat = torch.as_tensor(np.asarray(feat))
        feat = feat.detach()
        # the reference does feat.cpu() here (a synchronising D2H per batch, utils/metrics.py:106)
        if feat.is_cuda:
            # a snapshot, like the reference's feat.cpu(): the caller may overwrite its output buffer before compute()
            feats.append(feat.to(dev, dtype=torch.float32, copy=True))
            self._events.append(None)
        else:
            # upload on the side stream in pieces of <= MPREID_COPY_ROWS rows, one event each: compute() starts the
            # GEMM of a piece as soon as it has landed, so what is left after the final copy is one small chunk
            cs = _copy_stream(dev)
            piece = max(32, int(os.environ.get("MPREID_COPY_ROWS", "2048")))
            if feat.dtype != torch.float32:
                feat = feat.float()
            n = feat.shape[0]
            with torch.cuda.stream(cs):
                t = torch.empty(feat.shape, dtype=torch.float32, device=dev)
                for s0 in range(0, n, piece):
                    s1 = min(n, s0 + piece)
                    t[s0:s1].copy_(feat[s0:s1], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(cs)
                    feats.append(t[s0:s1])
                    self._events.append(ev)
        self.pids.extend(pid)
        self.camids.extend(camid)

    def _wait(self, lo, hi):
        cur = torch.cuda.current_stream(_device())   # the stream the kernels of compute() launch on (MPREID_DEVICE)
        for ev in self._events[lo:hi]:
            if ev is not None:
                cur.wait_event(ev)

    def _split_rows(self):
        """(batch index, tensor view) lists for the query rows and the gallery rows."""
        nq = self.num_query
        q_parts, g_parts, row = [], [], 0
        for bi, t in enumerate(self.feats):
            n = t.shape[0]
            if row + n <= nq:
                q_parts.append((bi, t))
            elif row >= nq:
                g_parts.append((bi, t))
            else:
                q_parts.append((bi, t[: nq - row]))
                g_parts.append((bi, t[nq - row:]))
            row += n
        return q_parts, g_parts

    def compute(self):  # called after each epoch
        with torch.cuda.device(_device()):   # MPREID_DEVICE may name a GPU other than torch's current one
            return self._compute()

    def _compute(self):
        if self.feat_norm:
            print("The test feature is normalized")
        nq = self.num_query
        q_pids = np.asarray(self.pids[:nq])
        q_camids = np.asarray(self.camids[:nq])
        g_pids = np.asarray(self.pids[nq:])
        g_camids = np.asarray(self.camids[nq:])
        norm = bool(self.feat_norm)
        if self.reranking:
            print('=> Enter reranking')
            self._wait(0, len(self.feats))
            prep = E.prep_rows(torch.cat(self.feats, dim=0), normalize=norm, precision=self._precision, keep_xn=True)
            q, g = prep.rows(0, nq), prep.rows(nq, prep.n)
            dist = _rerank_device(prep, nq, k1=50, k2=15, lambda_value=0.3, precision=self._precision)  # utils/metrics.py:127
            qf, gf = q.xn, g.xn
        else:
            print('=> Computing DistMat with euclidean_distance')
            devs = _devices()
            if len(devs) > 1 and nq >= len(devs):
                return self._compute_multi(devs, q_pids, g_pids, q_camids, g_camids, norm)
            q_parts, g_parts = self._split_rows()
            num_g = sum(t.shape[0] for _, t in g_parts)
            # (labels are uploaded by _eval_device AFTER the GEMMs: an early host->device copy on this stream would
            #  queue behind every feature piece on the copy engine and stall the whole pipeline)
            self._wait(0, (q_parts[-1][0] + 1) if q_parts else 0)
            q = E.prep_rows(torch.cat([t for _, t in q_parts], dim=0) if len(q_parts) != 1 else q_parts[0][1],
                            normalize=norm, precision=self._precision, keep_xn=True)
            dist = E.alloc_dist(nq, num_g, q.sqnorm.device)
            chunk_rows = max(32, int(os.environ.get("MPREID_CHUNK_ROWS", "8192")) // 32 * 32)
            tail_rows = min(chunk_rows, 2048)
            if all(self._events[bi] is None for bi, _ in g_parts):   # input already on the device: nothing to overlap with
                chunk_rows = tail_rows = max(num_g, 32)
            # gallery chunks: whatever has landed is contracted once it amounts to chunk_rows rows (tail_rows for
            # the last 2 * chunk_rows rows: the GEMM of a chunk can only start when its last row has arrived, so
            # the final chunk is what is left to do after the last host->device copy); cuts at multiples of 32
            # rows so that every column block of the distance matrix starts 128-byte aligned (vector stores)
            gf = torch.empty((num_g, q.D), dtype=torch.float32, device=q.sqnorm.device)   # normalised gallery rows (returned)
            off = 0
            pend, pend_rows, last_bi, arrived = [], 0, -1, 0

            def flush(rows_out):
                nonlocal pend, pend_rows, off
                take, rest = _take_rows(pend, rows_out)   # pieces of one upload are neighbouring views: no copy
                blk = take[0] if len(take) == 1 else torch.cat(take, dim=0)
                g = E.prep_rows(blk, normalize=norm, precision=self._precision, xn_out=gf[off:off + rows_out])
                E.dist_matrix(q, g, self._metric, self._precision, out=dist[:, off:off + rows_out])
                off += rows_out
                pend, pend_rows = rest, pend_rows - rows_out

            for bi, t in g_parts:
                self._wait(last_bi + 1, bi + 1)
                last_bi = bi
                pend.append(t); pend_rows += t.shape[0]
                arrived += t.shape[0]
                left = num_g - arrived
                if left == 0:
                    flush(pend_rows)
                elif pend_rows >= (chunk_rows if left > 2 * chunk_rows else tail_rows):
                    flush(pend_rows // 32 * 32)
            qf = q.xn
        cmc, mAP = _eval_device(dist, q_pids, g_pids, q_camids, g_camids, 50, self._junk)  # :132 (max_rank is not forwarded)
        lazy = LazyDistmat(dist)
        if os.environ.get("MPREID_DIST_MAT"):   # TEST.DIST_MAT (config/defaults.py:327): keep the matrix for offline analysis
            lazy.save(os.environ["MPREID_DIST_MAT"])
        return cmc, mAP, lazy, self.pids, self.camids, qf, gf

    def _compute_multi(self, devs, q_pids, g_pids, q_camids, g_camids, norm):
        """Single-process multi-GPU evaluation: query rows are sharded over `devs`, every gallery
        piece is fanned out from the upload device over NVLink as soon as it has landed, each device runs the same
        prep -> GEMM -> rank pipeline on its shard, and the per-query results are reduced by numpy on the host in
        query order -- bit-identical to the one-device result."""
        from .distributed import shard_bounds
        nq, P, d0 = self.num_query, len(devs), devs[0]
        q_parts, g_parts = self._split_rows()
        num_g = sum(t.shape[0] for _, t in g_parts)
        D = (q_parts[0][1] if q_parts else g_parts[0][1]).shape[1]
        bounds = [shard_bounds(nq, P, k) for k in range(P)]
        self._wait(0, (q_parts[-1][0] + 1) if q_parts else 0)
        qcat = _merge_adjacent([t for _, t in q_parts])
        qcat = qcat[0] if len(qcat) == 1 else torch.cat(qcat, dim=0)
        q, dist, qn = [], [], []
        for k, dev in enumerate(devs):
            lo, hi = bounds[k]
            with torch.cuda.device(dev):
                xq = qcat[lo:hi] if k == 0 else qcat[lo:hi].to(dev, non_blocking=True)
                q.append(E.prep_rows(xq, normalize=norm, precision=self._precision, keep_xn=True))
                dist.append(E.alloc_dist(hi - lo, num_g, dev))
        gf = torch.empty((num_g, D), dtype=torch.float32, device=d0)   # normalised gallery rows (returned)
        # fewer, larger chunks than on one device (every flush costs host time once per device); input that is already
        # on the device is contracted in one go
        streamed = any(self._events[bi] is not None for bi, _ in g_parts)
        base_rows = max(32, int(os.environ.get("MPREID_CHUNK_ROWS", "8192")) // 32 * 32)
        chunk_rows = 2 * base_rows if streamed else num_g
        tail_rows, tail_window = (min(chunk_rows, base_rows), 2 * base_rows) if streamed else (num_g, 0)
        pend, pend_rows, off, arrived, last_bi = [], 0, 0, 0, -1

        def flush(rows_out):
            # the chunk is assembled on device 0, broadcast to the peers over NVLink, then contracted everywhere
            nonlocal pend, pend_rows, off
            take, rest = _take_rows(pend, rows_out)
            blk0 = take[0] if len(take) == 1 else torch.cat(take, dim=0)
            blks = _fan_out(blk0, devs)
            for k, dev in enumerate(devs):
                with torch.cuda.device(dev):
                    g = E.prep_rows(blks[k], normalize=norm, precision=self._precision, keep_xn=False,
                                    xn_out=gf[off:off + rows_out] if k == 0 else None)
                    E.dist_matrix(q[k], g, self._metric, self._precision, out=dist[k][:, off:off + rows_out])
            off += rows_out
            pend, pend_rows = rest, pend_rows - rows_out

        for bi, t in g_parts:
            self._wait(last_bi + 1, bi + 1)
            last_bi = bi
            pend.append(t); pend_rows += t.shape[0]; arrived += t.shape[0]
            left = num_g - arrived
            if left == 0:
                flush(pend_rows)
            elif pend_rows >= (chunk_rows if left > tail_window else tail_rows):
                flush(pend_rows // 32 * 32)
        # rank / AP per shard, then ONE host-side reduction in query order
        results = []
        for k, dev in enumerate(devs):
            lo, hi = bounds[k]
            results.append(E.rank_eval_async(dist[k], q_pids[lo:hi], g_pids, q_camids[lo:hi], g_camids, self._junk))
        fh, ap, nr = [], [], []
        for k, dev in enumerate(devs):
            lo, hi = bounds[k]
            with torch.cuda.device(dev):
                a, b, c, st = results[k].to_host()
                if int(st[0]) != 0:   # positives workspace too small on this shard: exact re-run
                    a, b, c = E.rank_eval_host(dist[k], q_pids[lo:hi], g_pids, q_camids[lo:hi], g_camids, self._junk)
            fh.append(a); ap.append(b); nr.append(c)
        max_rank = 50
        if num_g < max_rank:
            max_rank = num_g
            print("Note: number of gallery samples is quite small, got {}".format(num_g))
        cmc, mAP = E.reduce_cmc_map(np.concatenate(fh), np.concatenate(ap), np.concatenate(nr), max_rank, num_g)
        qf = torch.cat([q[k].xn.to(d0) for k in range(P)], dim=0)
        return cmc, mAP, LazyDistmatShards(dist), self.pids, self.camids, qf, gf
