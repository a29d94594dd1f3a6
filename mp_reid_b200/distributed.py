"""Multi-GPU plumbing for the hot path: one process per GPU, torch.distributed.

dist + rank + CMC/AP shards over QUERY rows — queries are independent (utils/metrics.py:47-80 keeps
no cross-query state) — so the data path needs no collective: the gallery is broadcast once (NCCL
over NVLink), every rank evaluates its own query rows, and the per-query (first_hit, AP, num_rel)
triples are all-gathered ONCE.  The final cmc / mAP are then computed by numpy on the gathered
arrays in global query order, which makes the N-GPU result bit-identical to the 1-GPU result
(an all-reduce of partial sums would be order dependent in the last bit of mAP).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import engine as E


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
    dist.broadcast(gf, src=src, group=group)
    return gf, g_pid, g_cam


def gather_per_query(first_hit: torch.Tensor, ap: torch.Tensor, num_rel: torch.Tensor, counts: list[int], group=None):
    """All-gather the per-query results of every rank (one collective) -> numpy arrays in global query order.

    counts[r] = number of queries of rank r (shards may be ragged); the payload is padded to max(counts).
    first_hit / num_rel travel as float64 (exact for |x| < 2^53) next to the float64 AP.
    """
    world = dist.get_world_size(group)
    width = max(counts)
    packed = torch.zeros((3, width), dtype=torch.float64, device=ap.device)
    n = ap.shape[0]
    packed[0, :n] = first_hit.to(torch.float64)
    packed[1, :n] = ap
    packed[2, :n] = num_rel.to(torch.float64)
    out = torch.empty((world * 3, width), dtype=torch.float64, device=ap.device)  # concatenation along dim 0
    dist.all_gather_into_tensor(out, packed, group=group)
    h = out.cpu().numpy().reshape(world, 3, width)
    fh = np.concatenate([h[r, 0, :counts[r]] for r in range(world)]).astype(np.int32)
    apv = np.concatenate([h[r, 1, :counts[r]] for r in range(world)])
    nr = np.concatenate([h[r, 2, :counts[r]] for r in range(world)]).astype(np.int32)
    return fh, apv, nr


def sharded_reduce(first_hit, ap, num_rel, counts, max_rank: int, num_g: int, group=None, denominators="valid"):
    """gather_per_query + the host reduction of utils/metrics.py:82-86 (same on every rank)."""
    fh, apv, nr = gather_per_query(first_hit, ap, num_rel, counts, group)
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
    dist.all_gather_into_tensor(out, pad, group=group)
    full = torch.empty((N,) + tuple(x_local.shape[1:]), dtype=x_local.dtype, device=x_local.device)
    for r in range(world):
        n = int(ids_all[r].numel())
        full[ids_all[r]] = out[r * width: r * width + n]
    return full


def rerank_sharded(prep_all: "E.Prepared", nq: int, k1: int, k2: int, lambda_value: float, precision=None, group=None,
                   world: int | None = None, rank: int | None = None, exchange=None):
    """utils/reranking.py:29-100 with the rows of the all-pairs matrix sharded over the ranks of `group`.

    prep_all: the prepared stacked features (queries first) replicated on every rank.
    Returns (final_local [Qs, G], (q_lo, q_hi)): the re-ranked distances of this rank's query rows.
    `exchange` lets the tests substitute the all-gather (single-device emulation of several ranks).
    """
    N = prep_all.n
    dev = prep_all.sqnorm.device
    world = dist.get_world_size(group) if world is None else world
    rank = dist.get_rank(group) if rank is None else rank
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
    K = E.rerank_neighbor_count(k1, k2)
    nbr_local = E.row_topk(rows, K, rm)                                           # :46-48
    gather = exchange or (lambda x: _allgather_rows(x, ids_all, N, group))
    nbr_all = gather(nbr_local)
    ids32 = row_ids.to(torch.int32)
    v0 = E.rerank_build_v0(rows, ids32, N, k1, nbr_all, rm)                       # :51-71
    v0_all = tuple(gather(t) for t in v0)
    q_ids = ids32[:nq_local].contiguous()
    final_local = E.rerank_finish(nbr_all, v0_all, rows[:nq_local], q_ids, rm[:nq_local], N, nq, k1, k2, lambda_value)  # :73-99
    return final_local, (q_lo, q_hi)
