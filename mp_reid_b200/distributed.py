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
