"""Multi-GPU plumbing for the hot path: one process per GPU, torch.distributed.

dist + rank + CMC/AP shards over QUERY rows — queries are independent (utils/metrics.py:47-80 keeps
no cross-query state) — so the data path needs no collective: the gallery is broadcast once (NCCL
over NVLink), every rank evaluates its own query rows, and the per-query (first_hit, AP, num_rel)
triples are all-gathered ONCE.  The final cmc / mAP are then computed by numpy on the gathered
arrays in global query order, which makes the N-GPU result bit-identical to the 1-GPU result
(an all-reduce of partial sums would be order dependent in the last bit of mAP).
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.distributed as dist

from . import engine as E


# ------------------------------------------------------------------------------------------------
# Collectives.  NCCL is the product path.  With any other backend (gloo: the CPU tests, and the two-ranks-on-ONE-GPU
# GPU test, which NCCL refuses) device tensors are staged through host memory, so the same code runs everywhere.
def _is_nccl(group=None) -> bool:
    return dist.get_backend(group) == "nccl"


class _StreamWork:
    """Stand-in for the Work handle of an asynchronous NCCL call: wait() makes the current stream wait for `event`."""

    def __init__(self, event=None):
        self.event = event

    def wait(self):
        if self.event is not None:
            torch.cuda.current_stream().wait_event(self.event)


def all_gather_into(out: torch.Tensor, inp: torch.Tensor, group=None):
    if inp.is_cuda and not _is_nccl(group):
        host = torch.empty(out.shape, dtype=out.dtype)
        dist.all_gather_into_tensor(host, inp.cpu(), group=group)
        out.copy_(host)
    else:
        dist.all_gather_into_tensor(out, inp, group=group)


def broadcast(t: torch.Tensor, src: int, group=None, async_op: bool = False):
    """dist.broadcast; returns an object with wait() when async_op (the current stream then waits for the data)."""
    if t.is_cuda and not _is_nccl(group):
        host = t.cpu()                                   # synchronises the current stream: the source rows are complete
        dist.broadcast(host, src=src, group=group)
        t.copy_(host)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        return _StreamWork(ev)
    work = dist.broadcast(t, src=src, group=group, async_op=async_op)
    return work if async_op else _StreamWork()


def shard_bounds(n: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous, balanced row shard [lo, hi) of n rows for `rank` of `world`."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def broadcast_gallery(gf: torch.Tensor | None, g_pid, g_cam, shape=None, src: int = 0, group=None, device=None):
    """Broadcast gallery features + labels from `src` (everyone else passes gf=None and `shape`)."""
    rank = dist.get_rank(group)
    meta = [None]
    if rank == src:
        meta = [(tuple(gf.shape), np.asarray(g_pid, dtype=np.int64), np.asarray(g_cam, dtype=np.int64))]
    dist.broadcast_object_list(meta, src=src, group=group)
    gshape, g_pid, g_cam = meta[0]
    if rank != src:
        gf = torch.empty(gshape, dtype=torch.float32, device=device)
    broadcast(gf, src, group)
    return gf, g_pid, g_cam


def gather_per_query(first_hit: torch.Tensor, ap: torch.Tensor, num_rel: torch.Tensor, counts: list[int] | None, group=None,
                     ids: torch.Tensor | None = None, total: int | None = None):
    """All-gather the per-query results of every rank (one collective) -> numpy arrays in global query order.

    counts[r] = number of queries of rank r when the shards are contiguous (shards may be ragged; the payload is padded
    to max(counts)).  ids (with total = number of queries overall): the GLOBAL query index of every local result instead,
    for shards that are not contiguous (block-cyclic re-ranking); results are scattered into place by it.
    first_hit / num_rel travel as float64 (exact for |x| < 2^53) next to the float64 AP.
    """
    world = dist.get_world_size(group)
    n = ap.shape[0]
    if ids is None:
        width, rows = max(counts), 3
    else:
        width, rows = (total + world - 1) // world + 512, 4     # an upper bound every rank can compute (256-row blocks, cyclic)
    packed = torch.zeros((rows, width), dtype=torch.float64, device=ap.device)
    packed[0, :n] = first_hit.to(torch.float64)
    packed[1, :n] = ap
    packed[2, :n] = num_rel.to(torch.float64)
    if ids is not None:
        packed[3].fill_(-1.0)
        packed[3, :n] = ids.to(torch.float64)
    out = torch.empty((world * rows, width), dtype=torch.float64, device=ap.device)  # concatenation along dim 0
    all_gather_into(out, packed, group)
    h = out.cpu().numpy().reshape(world, rows, width)
    if ids is None:
        fh = np.concatenate([h[r, 0, :counts[r]] for r in range(world)]).astype(np.int32)
        apv = np.concatenate([h[r, 1, :counts[r]] for r in range(world)])
        nr = np.concatenate([h[r, 2, :counts[r]] for r in range(world)]).astype(np.int32)
        return fh, apv, nr
    fh, apv, nr = np.zeros(total, np.int32), np.zeros(total, np.float64), np.zeros(total, np.int32)
    seen = 0
    for r in range(world):
        keep = h[r, 3] >= 0
        at = h[r, 3][keep].astype(np.int64)
        fh[at], apv[at], nr[at] = h[r, 0][keep].astype(np.int32), h[r, 1][keep], h[r, 2][keep].astype(np.int32)
        seen += int(keep.sum())
    assert seen == total, "sharded results do not cover every query exactly once"
    return fh, apv, nr


def sharded_reduce(first_hit, ap, num_rel, counts, max_rank: int, num_g: int, group=None, denominators="valid", ids=None, total=None):
    """gather_per_query + the host reduction of utils/metrics.py:82-86 (same on every rank)."""
    fh, apv, nr = gather_per_query(first_hit, ap, num_rel, counts, group, ids=ids, total=total)
    return E.reduce_cmc_map(fh, apv, nr, max_rank, num_g, denominators)


def evaluate_sharded(qf_local: torch.Tensor, q_pid_local, q_cam_local, gf: torch.Tensor, g_pid, g_cam, counts,
                     feat_norm=True, metric="sqeuclid", precision=None, junk=None, max_rank=50, group=None):
    """dist + rank + CMC/mAP with the query rows of this rank against the (replicated) gallery."""
    q = E.prep_rows(qf_local, normalize=bool(feat_norm), precision=precision, keep_xn=False)
    g = E.prep_rows(gf, normalize=bool(feat_norm), precision=precision, keep_xn=False)
    d = E.dist_matrix(q, g, metric, precision)
    fh, ap, nr = E.rank_eval(d, q_pid_local, g_pid, q_cam_local, g_cam, junk)
    num_g = gf.shape[0]
    return sharded_reduce(fh, ap, nr, counts, min(max_rank, num_g), num_g, group)


# ------------------------------------------------------------------------------------------------
# re-ranking across GPUs: the (Q+G)^2 matrix is row-sharded, the neighbour lists and
# the V0 rows are all-gathered (the one exchange step of the path), the cheap sparse stages run
# redundantly on every rank and the Jaccard / blend pass is sharded over the query rows.
def rerank_row_ids(nq: int, N: int, world: int, rank: int, device=None) -> torch.Tensor:
    """Global sample indices owned by `rank`: its share of the query rows followed by its share of the gallery rows."""
    q_lo, q_hi = shard_bounds(nq, world, rank)
    g_lo, g_hi = shard_bounds(N - nq, world, rank)
    return torch.cat([torch.arange(q_lo, q_hi, device=device), nq + torch.arange(g_lo, g_hi, device=device)])


def _allgather_rows(x_local: torch.Tensor, ids_all: list[torch.Tensor], N: int, group=None) -> torch.Tensor:
    """All-gather row blocks of unequal height and scatter them into global row order -> [N, ...]."""
    world = dist.get_world_size(group)
    width = max(int(i.numel()) for i in ids_all)
    pad = torch.zeros((width,) + tuple(x_local.shape[1:]), dtype=x_local.dtype, device=x_local.device)
    pad[: x_local.shape[0]] = x_local
    out = torch.empty((world * width,) + tuple(x_local.shape[1:]), dtype=x_local.dtype, device=x_local.device)
    all_gather_into(out, pad, group)
    full = torch.empty((N,) + tuple(x_local.shape[1:]), dtype=x_local.dtype, device=x_local.device)
    for r in range(world):
        n = int(ids_all[r].numel())
        full[ids_all[r]] = out[r * width: r * width + n]
    return full


def _allgather_contig(x_local: torch.Tensor, counts: list[int], group=None) -> torch.Tensor:
    """All-gather contiguous row shards of unequal height -> [sum(counts), ...] in rank order."""
    world = len(counts)
    width = max(counts)
    pad = x_local
    if x_local.shape[0] != width:
        pad = torch.zeros((width,) + tuple(x_local.shape[1:]), dtype=x_local.dtype, device=x_local.device)
        pad[: x_local.shape[0]] = x_local
    out = torch.empty((world * width,) + tuple(x_local.shape[1:]), dtype=x_local.dtype, device=x_local.device)
    all_gather_into(out, pad.contiguous(), group)
    if all(c == width for c in counts):
        return out
    return torch.cat([out[r * width: r * width + counts[r]] for r in range(world)], dim=0)


def rerank_owned_queries(nq: int, world: int, rank: int, device=None) -> torch.Tensor:
    """Global indices of the query rows `rank` finishes in the fused row-sharded re-ranking: the rows of the 256-row
    blocks p with p % world == rank (the blocks whose tiles that rank contracts: their [Q, G] block rows are local)."""
    ids = torch.arange(nq, device=device)
    return ids[(ids // 256) % world == rank]


def _rerank_sharded_fused(prep_all, nq, k1, k2, lambda_value, precision, group, world, rank):
    """The fused (no N x N matrix) pipeline of reranking._rerank_fused with the tiles of the symmetric all-pairs pass
    dealt out by 256-row block, cyclically.  Exchanges: thresholds [N] (all-gather), row maxima [N] (all-reduce max),
    per-rank partial top-K keys [N, K] (all-gather, merged on every rank), V0 rows trimmed to their longest length
    (all-gather).  Every rank finishes the query rows of its own blocks.  Same bits as the one-GPU fused pipeline.
    Returns None if some neighbour list cannot be decided from the candidates (every rank sees the same flag)."""
    from .reranking import FUSED_SAMPLE
    N = prep_all.n
    dev = prep_all.sqnorm.device
    K = E.rerank_neighbor_count(k1, k2)
    S = min(N, FUSED_SAMPLE)
    t = min(K + 2, S)
    counts = [shard_bounds(N, world, r)[1] - shard_bounds(N, world, r)[0] for r in range(world)]
    lo, hi = shard_bounds(N, world, rank)
    ids = (torch.arange(S, device=dev, dtype=torch.int64) * N) // S
    smp = prep_all.take(ids)
    dS = E.dist_matrix(prep_all.rows(lo, hi), smp, "sqeuclid", precision)
    thr = _allgather_contig(E.row_kth(dS, t, bound=True) + 1e-6 * (prep_all.sqnorm[lo:hi] + prep_all.sqnorm.max()), counts, group)
    E.mark("rerank.thresholds")
    expect = N * t / S
    cap = int(min(N, max(256, (int(3 * expect) + 256 + 255) // 256 * 256)))
    cand, cnt, block, col0, row_max = E.dist_symmetric_topk(prep_all, thr, cap, nq, precision, own_mod=world, own_rank=rank)
    red = torch.cat([row_max, (cnt > cap).any().to(torch.float32).view(1)])
    if _is_nccl(group):
        dist.all_reduce(red, op=dist.ReduceOp.MAX, group=group)
    else:
        h = red.cpu(); dist.all_reduce(h, op=dist.ReduceOp.MAX, group=group); red.copy_(h)
    row_max = red[:N].contiguous()
    E.mark("rerank.all_pairs_gemm")
    keys, _ = E.cand_topk(cand, cnt, K, row_max, thr, partial=True)
    del cand
    if _is_nccl(group) and os.environ.get("MPREID_SHARD_A2A", "1") != "0":
        # every rank merges the rows of ITS shard: the partial keys of those rows arrive with one all-to-all (1/N of an
        # all-gather), the merged lists (indices + values) are then all-gathered -- every rank needs all of them for V0
        recv = torch.empty((world, hi - lo, K), dtype=torch.int64, device=dev)
        dist.all_to_all_single(recv.view(-1), keys.view(-1), [(hi - lo) * K] * world, [c * K for c in counts], group=group)
        nbr_loc, val_loc, status = E.merge_topk(recv, row_max[lo:hi].contiguous(), thr[lo:hi].contiguous())
        nbr = _allgather_contig(nbr_loc, counts, group)
        nbr_val = _allgather_contig(val_loc, counts, group)
    else:
        keys_all = torch.empty((world, N, K), dtype=torch.int64, device=dev)
        all_gather_into(keys_all.view(-1), keys.view(-1), group)
        nbr, nbr_val, status = E.merge_topk(keys_all, row_max, thr)
    E.mark("rerank.topk")
    row_ids = torch.arange(lo, hi, device=dev, dtype=torch.int32)
    v0c, v0v, v0l = E.rerank_build_v0_sparse(row_ids, hi - lo, N, k1, nbr, nbr_val, row_max[lo:hi], prep_all.xn, prep_all.sqnorm)
    info = torch.stack([v0l.max().to(torch.float32), status[0].to(torch.float32), red[N]])
    if _is_nccl(group):
        dist.all_reduce(info, op=dist.ReduceOp.MAX, group=group)
        info_h = info.cpu()
    else:
        info_h = info.cpu(); dist.all_reduce(info_h, op=dist.ReduceOp.MAX, group=group)
    if os.environ.get("MPREID_DEBUG"):
        print(f"[rerank_sharded rank {rank}] max V0 length {float(info_h[0])}, undecided rows {float(info_h[1])}, list overflow {float(info_h[2])}, "
              f"longest list {int(cnt.max())} of {cap}", flush=True)
    if float(info_h[1]) != 0.0 or float(info_h[2]) != 0.0:
        return None
    W = max(8, (int(info_h[0]) + 7) // 8 * 8)
    W = min(W, v0c.shape[1])
    v0_all = (_allgather_contig(v0c[:, :W].contiguous(), counts, group), _allgather_contig(v0v[:, :W].contiguous(), counts, group),
              _allgather_contig(v0l, counts, group))
    E.mark("rerank.v0")
    q_ids = rerank_owned_queries(nq, world, rank, dev)
    # more ranks than 256-row query blocks: a rank may finish no query row.  It still expands ITS rows of V and takes part
    # in every exchange below (the other ranks need those rows); only the per-query stages are skipped.
    idle = q_ids.numel() == 0
    empty = torch.empty((0, N - nq), dtype=torch.float32, device=dev)
    if k2 == 1 or os.environ.get("MPREID_SHARD_QE", "1") == "0":
        if idle:
            return empty, q_ids
        q32 = q_ids.to(torch.int32).contiguous()
        final = E.rerank_finish(nbr, v0_all, block, q32, row_max, N, nq, k1, k2, lambda_value, block_col0=col0, rows_global=True)
        return final, q_ids
    # query expansion sharded like the V0 rows: expand the rows [lo, hi) into the finish workspace, all-gather the other ranks'
    # rows (trimmed to the longest one) into place, then inverted index + Jaccard + blend
    ws = E.rerank_finish_workspace(N, nq, k1, k2, dev)
    # stage 8 reads neither the query ids nor the output; an idle rank passes one placeholder row to satisfy the argument checks
    q32 = (torch.zeros(1, dtype=torch.int64, device=dev) if idle else q_ids).to(torch.int32).contiguous()
    final = E.alloc_dist(int(q32.numel()), N - nq, dev)
    args = (nbr, v0_all, block, q32, row_max, N, nq, k1, k2, lambda_value)
    E.rerank_finish(*args, out=final, block_col0=col0, rows_global=True, stages=8, ws=ws, qe_rows=(lo, hi))
    v_col, v_val, v_len = E.rerank_finish_v_views(ws, N, nq, k1, k2)
    wmax = v_len[lo:hi].max().to(torch.float32).view(1)
    if _is_nccl(group):
        dist.all_reduce(wmax, op=dist.ReduceOp.MAX, group=group)
        W1 = int(wmax.item())
    else:
        h = wmax.cpu(); dist.all_reduce(h, op=dist.ReduceOp.MAX, group=group); W1 = int(h.item())
    W1 = min(v_col.shape[1], max(8, (W1 + 7) // 8 * 8))
    v_col[:, :W1] = _allgather_contig(v_col[lo:hi, :W1].contiguous(), counts, group)
    v_val[:, :W1] = _allgather_contig(v_val[lo:hi, :W1].contiguous(), counts, group)
    v_len.copy_(_allgather_contig(v_len[lo:hi].contiguous(), counts, group))
    if idle:
        return empty, q_ids
    E.rerank_finish(*args, out=final, block_col0=col0, rows_global=True, stages=16 | 4 | 2, ws=ws)
    return final, q_ids


def _rerank_sharded_plain(prep_all, nq, k1, k2, lambda_value, precision, group, world, rank, exchange=None):
    """The materialising form: every rank contracts its own row block of the all-pairs matrix (its share of the query rows
    followed by its share of the gallery rows), neighbour lists and V0 rows are all-gathered."""
    N = prep_all.n
    dev = prep_all.sqnorm.device
    ids_all = [rerank_row_ids(nq, N, world, r, dev) for r in range(world)]
    row_ids = ids_all[rank]
    q_lo, q_hi = shard_bounds(nq, world, rank)
    nq_local = q_hi - q_lo
    local = prep_all.take(row_ids)
    R = int(row_ids.numel())
    ld = (N + 31) // 32 * 32
    rows = torch.empty((R, ld), dtype=torch.float32, device=dev)[:, :N]
    rm = torch.empty((R,), dtype=torch.float32, device=dev)
    E.dist_matrix(local, prep_all, "sqeuclid", precision, out=rows, row_max=rm)   # :36-41 + the maxima of :46
    E.mark("rerank.all_pairs_gemm")
    K = E.rerank_neighbor_count(k1, k2)
    nbr_local = E.row_topk(rows, K, rm)                                           # :46-48
    gather = exchange or (lambda x: _allgather_rows(x, ids_all, N, group))
    nbr_all = gather(nbr_local)
    E.mark("rerank.topk")
    ids32 = row_ids.to(torch.int32)
    v0 = E.rerank_build_v0(rows, ids32, N, k1, nbr_all, rm)                       # :51-71
    v0_all = tuple(gather(t) for t in v0)
    E.mark("rerank.v0")
    q_ids = ids32[:nq_local].contiguous()
    if nq_local == 0:      # fewer queries than ranks: nothing to finish here (every exchange is behind us)
        return torch.empty((0, N - nq), dtype=torch.float32, device=dev), torch.arange(q_lo, q_hi, device=dev)
    final_local = E.rerank_finish(nbr_all, v0_all, rows[:nq_local], q_ids, rm[:nq_local], N, nq, k1, k2, lambda_value)  # :73-99
    return final_local, torch.arange(q_lo, q_hi, device=dev)


def rerank_sharded(prep_all: "E.Prepared", nq: int, k1: int, k2: int, lambda_value: float, precision=None, group=None,
                   world: int | None = None, rank: int | None = None, exchange=None):
    """utils/reranking.py:29-100 with the work of the all-pairs pass and of the sparse stages sharded over the ranks of `group`.

    prep_all: the prepared stacked features (queries first) replicated on every rank.
    Returns (final_local [Qs, G], q_ids [Qs]): the re-ranked distances of the query rows this rank finished and their global
    indices (256-row blocks dealt out cyclically in the fused form, one contiguous shard in the materialising form).
    `exchange` lets the tests substitute the all-gather (single-device emulation of several ranks; materialising form only).
    """
    from .reranking import fused_enabled
    world = dist.get_world_size(group) if world is None else world
    rank = dist.get_rank(group) if rank is None else rank
    if exchange is None and prep_all.xn is not None and fused_enabled(prep_all.n, precision):
        out = _rerank_sharded_fused(prep_all, nq, k1, k2, lambda_value, precision, group, world, rank)
        if out is not None:
            return out
    return _rerank_sharded_plain(prep_all, nq, k1, k2, lambda_value, precision, group, world, rank, exchange)


# ------------------------------------------------------------------------------------------------
# Cooperative evaluation from host memory: every rank feeds ITS queries and ITS slice of the gallery
# (what a sharded feature extraction leaves on each rank), the slices travel once over NVLink.
def aligned_shard_bounds(n: int, world: int, rank: int, align: int = 32) -> tuple[int, int]:
    """Contiguous row shard [lo, hi) whose start is a multiple of `align` (column blocks of the distance
    matrix then start 128-byte aligned and keep the vector-store path); the last shards may be short or empty."""
    per = ((n + world - 1) // world + align - 1) // align * align
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def _metrics():
    from . import metrics   # late: metrics imports this module lazily too
    return metrics


# Peer-to-peer exchange of the gallery slices (NCCL runs only): every rank publishes its uploaded slice in a symmetric-memory
# buffer (torch.distributed._symmetric_memory: the same allocation mapped into every process of the node) and PULLS the other
# slices with plain device-to-device copies, which the copy engines move over NVLink -- no SM is taken from the persistent
# distance GEMM, unlike the NCCL broadcast kernels.  MPREID_SHARD_EXCHANGE=nccl keeps the broadcasts.
_SYMM = {}


def _symm_slice_buffer(rows: int, D: int, dev, group):
    """-> dict(buf [cap, D] fp32 symmetric, hdl, rows=cap) with cap >= rows, or None (every rank takes the same answer)."""
    if os.environ.get("MPREID_SHARD_EXCHANGE", "p2p").lower() != "p2p" or not _is_nccl(group):
        return None
    key = (id(group), int(D), dev.index)
    ent = _SYMM.get(key)
    if ent is False:
        return None
    if ent is not None and ent["rows"] >= rows:
        return ent
    ok, new = 1.0, None
    try:
        import torch.distributed._symmetric_memory as symm
        cap = (int(rows) + 31) // 32 * 32
        buf = symm.empty((cap, D), dtype=torch.float32, device=dev)
        hdl = symm.rendezvous(buf, group if group is not None else dist.group.WORLD)
        new = dict(buf=buf, hdl=hdl, rows=cap)
    except Exception as e:   # no P2P mapping on this box / build: fall back for good
        ok = 0.0
        if dist.get_rank(group) == 0:
            print(f"[mp_reid_b200] symmetric-memory exchange unavailable ({type(e).__name__}: {e}); using NCCL broadcasts", flush=True)
    flag = torch.tensor([ok], dtype=torch.float32, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)      # all ranks or none
    if float(flag.item()) < 1.0:
        _SYMM[key] = False
        return None
    _SYMM[key] = new
    return new


def sharded_evaluator(num_query_local: int, **kw):
    """R1_mAP_eval whose compute() is cooperative over the ranks of `group` (keyword, default WORLD): update() takes
    this rank's query batches followed by this rank's gallery batches.  See R1_mAP_eval_sharded.compute."""
    return _make_sharded_class()(num_query_local, **kw)


_SHARDED_CLS = None


def _make_sharded_class():
    global _SHARDED_CLS
    if _SHARDED_CLS is not None:
        return _SHARDED_CLS
    M = _metrics()

    class R1_mAP_eval_sharded(M.R1_mAP_eval):
        """Same surface as R1_mAP_eval (utils/metrics.py:91-134); every rank calls reset / update / compute.

        Rank r holds `num_query` of the queries and a contiguous slice of the gallery (slices in rank order form the
        gallery).  compute(): each slice is broadcast from its owner in sub-blocks (NCCL over NVLink, asynchronous,
        on a side stream) straight into its place in a [G, D] buffer, so the distance GEMM of one sub-block overlaps the
        upload and transfer of the next ones and the columns keep the global gallery order (ties break exactly as on
        one GPU); every rank ranks its own queries; the per-query results are all-gathered once and reduced by numpy in
        global query order: every rank returns the same (cmc, mAP) as a one-GPU evaluation of all queries.
        distmat / qf are this rank's rows, gf the full normalised gallery.
        """

        def __init__(self, num_query, *args, group=None, **kw):
            super().__init__(num_query, *args, **kw)
            self._group = group

        def reset(self):
            super().reset()
            self._lab_dev = []

        def update(self, output):
            super().update(output)
            # labels ride the upload stream right behind their features: a host->device copy issued later, from
            # compute(), would queue behind every feature piece on the copy engine and hold back the exchange
            _, pid, camid = output
            dev = M._device()
            lab = np.stack([np.asarray(pid, dtype=np.int64).reshape(-1), np.asarray(camid, dtype=np.int64).reshape(-1)])
            cs = M._copy_stream(dev)
            with torch.cuda.stream(cs):
                # pinned staging: a pageable source would make this call wait for the feature copies queued ahead of it
                self._lab_dev.append(torch.from_numpy(lab).pin_memory().to(dev, non_blocking=True))

        def compute(self):
            if self.reranking:
                raise NotImplementedError("sharded evaluator: use distributed.rerank_sharded for re-ranking")
            if self.feat_norm:
                print("The test feature is normalized")
            print('=> Computing DistMat with euclidean_distance')
            group, norm = self._group, bool(self.feat_norm)
            world, rank = dist.get_world_size(group), dist.get_rank(group)
            nq = self.num_query
            q_parts, g_parts = self._split_rows()
            dev = M._device()
            n_loc = sum(t.shape[0] for _, t in g_parts)
            D = (q_parts[0][1] if q_parts else g_parts[0][1]).shape[1]
            # sizes + labels of every slice: two small all-gathers (slice sizes may differ); no host->device copy here
            meta = torch.empty((2,), dtype=torch.int64, device=dev)
            meta[0].fill_(nq); meta[1].fill_(n_loc)
            metas = torch.empty((world, 2), dtype=torch.int64, device=dev)
            all_gather_into(metas.view(-1), meta, group)
            metas = metas.cpu().numpy()
            q_counts, g_counts = [int(x) for x in metas[:, 0]], [int(x) for x in metas[:, 1]]
            S = max(g_counts + [1])
            num_g = int(sum(g_counts))
            offs = np.concatenate([[0], np.cumsum(g_counts)]).astype(np.int64)
            # Every slice travels in sub-blocks of <= MPREID_SHARD_SUB_ROWS rows (multiples of 32: aligned columns), one
            # broadcast each, issued asynchronously in the order (sub-block, owner) on a side stream that waits only for
            # the upload pieces that sub-block needs: the GEMMs of early sub-blocks run while later ones are still on
            # PCIe / NVLink, and the compute stream never waits for more than it is about to use.
            # (sub-block height: ~8,192 rows for two ranks, ~2,048 from four ranks on -- with many owners a stage is world
            #  sub-blocks, and the GEMMs of the first stage should start soon after the queries have landed)
            target = int(os.environ.get("MPREID_SHARD_SUB_ROWS", "8192" if world <= 2 else "2048"))
            n_st = max(1, -(-S // max(32, target)))
            sub = max(32, (-(-S // n_st) + 31) // 32 * 32)
            gfull = torch.empty((num_g, D), dtype=torch.float32, device=dev)
            xs = M._side_stream(dev, "exchange")
            main = torch.cuda.current_stream(dev)
            xs.wait_stream(main)                       # gfull exists, labels gathered
            copied, next_piece = 0, 0
            n_sub = [(c + sub - 1) // sub for c in g_counts]
            stages = max(n_sub + [0])
            symm = _symm_slice_buffer(S, D, dev, group) if world > 1 else None

            def land_own(j):
                """(side stream) this rank's rows of sub-block j: wait for their upload pieces, move them into place."""
                nonlocal copied, next_piece
                lo = int(offs[rank]) + j * sub
                hi = min(int(offs[rank + 1]), lo + sub)
                need = hi - int(offs[rank])   # own rows that must have landed
                while copied < need:
                    bi, t = g_parts[next_piece]
                    ev = self._events[bi]
                    if ev is not None:
                        xs.wait_event(ev)
                    else:
                        xs.wait_stream(main)
                    gfull[int(offs[rank]) + copied: int(offs[rank]) + copied + t.shape[0]].copy_(t, non_blocking=True)
                    copied += t.shape[0]; next_piece += 1
                return lo, hi

            def exchange_nccl(j):
                """Enqueue (side stream) the broadcasts of sub-block j of every slice -> [(lo, hi, work)]."""
                out = []
                with torch.cuda.stream(xs):
                    for r in range(world):
                        if j >= n_sub[r]:
                            continue
                        lo = int(offs[r]) + j * sub
                        hi = min(int(offs[r + 1]), lo + sub)
                        if r == rank:
                            land_own(j)
                        src = dist.get_global_rank(group, r) if group is not None else r
                        out.append((lo, hi, broadcast(gfull[lo:hi], src, group, async_op=True)))
                return out

            def exchange_p2p(j):
                """Sub-block j of every slice by copy engine: publish the own rows in the symmetric buffer, one signal-pad
                barrier (a one-warp kernel, no shared memory: it fits next to the GEMM CTAs), then pull the peers' rows."""
                out = []
                hdl, sbuf = symm["hdl"], symm["buf"]
                with torch.cuda.stream(xs):
                    if j < n_sub[rank]:
                        lo, hi = land_own(j)
                        sbuf[j * sub: j * sub + (hi - lo)].copy_(gfull[lo:hi], non_blocking=True)
                        ev = torch.cuda.Event(); ev.record(xs)
                        out.append((lo, hi, _StreamWork(ev)))     # the own rows do not wait for anybody
                    hdl.barrier(channel=0)                        # every rank has published its sub-block j
                    for r in range(world):
                        if r == rank or j >= n_sub[r]:
                            continue
                        lo = int(offs[r]) + j * sub
                        hi = min(int(offs[r + 1]), lo + sub)
                        peer = hdl.get_buffer(r, (symm["rows"], D), torch.float32)
                        gfull[lo:hi].copy_(peer[j * sub: j * sub + (hi - lo)], non_blocking=True)
                        ev = torch.cuda.Event(); ev.record(xs)
                        out.append((lo, hi, _StreamWork(ev)))
                return out

            exchange = exchange_p2p if symm is not None else exchange_nccl

            # queries of this rank
            self._wait(0, (q_parts[-1][0] + 1) if q_parts else 0)
            qv = M._merge_adjacent([t for _, t in q_parts])
            q = E.prep_rows(qv[0] if len(qv) == 1 else torch.cat(qv, dim=0), normalize=norm, precision=self._precision, keep_xn=True)
            dmat = E.alloc_dist(nq, num_g, dev)
            simt = (self._precision or E.default_precision()).lower() in ("simt", "fp32_simt")
            planes = None if simt else E.GalleryPlanes(num_g, D, dev, self._precision, keep_xn=True)
            gf = torch.empty((num_g, D), dtype=torch.float32, device=dev) if simt else planes.xn
            # host-side software pipeline: the broadcasts of stage j+1 are enqueued before the GEMMs of stage j (an NCCL
            # call costs ~0.1 ms of host time; issuing all of them up front would delay the first GEMM by milliseconds)
            pending = exchange(0) if stages else []
            for j in range(stages):
                nxt = exchange(j + 1) if j + 1 < stages else []
                for lo, hi, work in pending:
                    work.wait()                        # the compute stream waits for this sub-block only
                    if planes is not None:             # two C calls on slices of buffers allocated once
                        planes.add_block(gfull[lo:hi], lo, norm, q, self._metric, dmat)
                    else:
                        g = E.prep_rows(gfull[lo:hi], normalize=norm, precision=self._precision, xn_out=gf[lo:hi])
                        E.dist_matrix(q, g, self._metric, self._precision, out=dmat[:, lo:hi])
                pending = nxt
            # labels last: they were uploaded behind their features, the late ones land when the last piece does
            main.wait_stream(M._copy_stream(dev))
            lab_all = torch.cat(self._lab_dev, dim=1) if self._lab_dev else torch.zeros((2, 0), dtype=torch.int64, device=dev)
            lab = torch.zeros((2, S), dtype=torch.int64, device=dev)
            lab[:, :n_loc] = lab_all[:, nq:nq + n_loc]
            labs = torch.empty((world * 2, S), dtype=torch.int64, device=dev)
            all_gather_into(labs, lab, group)
            labs = labs.view(world, 2, S)
            g_pid = torch.cat([labs[r, 0, :g_counts[r]] for r in range(world)])
            g_cam = torch.cat([labs[r, 1, :g_counts[r]] for r in range(world)])
            fh, ap, nr = E.rank_eval(dmat, lab_all[0, :nq].contiguous(), g_pid, lab_all[1, :nq].contiguous(), g_cam, self._junk)
            max_rank = 50
            if num_g < max_rank:
                max_rank = num_g
                print("Note: number of gallery samples is quite small, got {}".format(num_g))
            cmc, mAP = sharded_reduce(fh, ap, nr, q_counts, max_rank, num_g, group)
            return cmc, mAP, M.LazyDistmat(dmat), self.pids, self.camids, q.xn, gf

    _SHARDED_CLS = R1_mAP_eval_sharded
    return _SHARDED_CLS
