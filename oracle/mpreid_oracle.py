"""CPU oracle for the MP-ReID evaluation / retrieval hot path.

TEST INFRASTRUCTURE ONLY.  This is a numpy (+ torch-CPU sgemm) restatement of the
reference algorithm; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import it.  The product path (mp_reid_b200/) never does.

Parity pinning: the reference ships no tests or golden vectors, so the oracle
is pinned against the reference ITSELF: oracle/make_golden.py imports the unmodified
/root/reference/utils/{metrics,reranking}.py in the build container, runs both on seeded
inputs and commits the reference outputs under tests/golden/; tests/test_oracle_golden.py
checks this file against those fixtures.

Every function cites the reference lines it restates (paths relative to /root/reference).
"""
from __future__ import annotations

import numpy as np

try:  # the reference distance is a torch-CPU sgemm; use the same BLAS entry when torch is present
    import torch
except Exception:  # pragma: no cover
    torch = None


# --------------------------------------------------------------------------------------
# distance matrices
# --------------------------------------------------------------------------------------
def l2_normalize(x: np.ndarray, eps: float = 1e-12) -> np.ndarray:
    """torch.nn.functional.normalize(x, dim=1, p=2) as called at utils/metrics.py:114."""
    if torch is not None:
        return torch.nn.functional.normalize(torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)), dim=1, p=2).numpy()
    n = np.sqrt((x.astype(np.float32) ** 2).sum(1, keepdims=True, dtype=np.float32))
    return (x / np.maximum(n, eps)).astype(np.float32)


def sq_euclidean(qf: np.ndarray, gf: np.ndarray) -> np.ndarray:
    """utils/metrics.py:7-13.  SQUARED distance ||q||^2 + ||g||^2 - 2 q.g, fp32, no sqrt, no clamp.

    The reference forms the [m,n] sum of squared norms first and then accumulates -2*q@g^T
    into it with one BLAS call (addmm_ with beta=1, alpha=-2).
    """
    qf = np.ascontiguousarray(qf, dtype=np.float32)
    gf = np.ascontiguousarray(gf, dtype=np.float32)
    if torch is not None:
        q, g = torch.from_numpy(qf), torch.from_numpy(gf)
        base = (q * q).sum(dim=1, keepdim=True) + (g * g).sum(dim=1, keepdim=True).t()
        return torch.addmm(base, q, g.t(), beta=1, alpha=-2).numpy()
    base = (qf * qf).sum(1, keepdims=True) + (gf * gf).sum(1)[None, :]
    return (base + np.float32(-2) * (qf @ gf.T)).astype(np.float32)


def arccos_cosine(qf: np.ndarray, gf: np.ndarray, epsilon: float = 1e-5) -> np.ndarray:
    """utils/metrics.py:15-25.  arccos(clip(q.g * (1 / (||q|| ||g||)), -1+eps, 1-eps)), fp32 radians."""
    qf = np.ascontiguousarray(qf, dtype=np.float32)
    gf = np.ascontiguousarray(gf, dtype=np.float32)
    if torch is not None:
        q, g = torch.from_numpy(qf), torch.from_numpy(gf)
        dots = q.mm(g.t())
        nn_ = torch.norm(q, p=2, dim=1, keepdim=True).mm(torch.norm(g, p=2, dim=1, keepdim=True).t())
        cosv = dots.mul(1 / nn_).numpy()
    else:
        nq = np.sqrt((qf * qf).sum(1, keepdims=True))
        ng = np.sqrt((gf * gf).sum(1, keepdims=True))
        cosv = ((qf @ gf.T) * (np.float32(1) / (nq @ ng.T))).astype(np.float32)
    return np.arccos(np.clip(cosv, -1 + epsilon, 1 - epsilon))


def one_minus_cosine(qf: np.ndarray, gf: np.ndarray) -> np.ndarray:
    """processor/processor_uniprompt_stage2.py:466-468.  1 - q.g on already-normalised features."""
    qf = np.ascontiguousarray(qf, dtype=np.float32)
    gf = np.ascontiguousarray(gf, dtype=np.float32)
    if torch is not None:
        return (1 - torch.matmul(torch.from_numpy(qf), torch.from_numpy(gf).t())).numpy()
    return (np.float32(1) - qf @ gf.T).astype(np.float32)


def sqrt_euclidean(x: np.ndarray, y: np.ndarray) -> np.ndarray:
    """loss/triplet_loss.py:16-31: sqrt(clamp(||x||^2+||y||^2-2x.y, 1e-12))."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.ascontiguousarray(y, dtype=np.float32)
    d = (x * x).sum(1, keepdims=True) + (y * y).sum(1)[None, :] - np.float32(2) * (x @ y.T)
    return np.sqrt(np.maximum(d.astype(np.float32), np.float32(1e-12)))


# --------------------------------------------------------------------------------------
# ranking + CMC / mAP
# --------------------------------------------------------------------------------------
def rank_eval(distmat, q_pids, g_pids, q_camids, g_camids, max_rank=50, sort_kind="stable",
              junk="none", verbose=False):
    """utils/metrics.py:28-88 (eval_func), returning the per-query quantities too.

    sort_kind: 'stable' is the tie contract of the GPU path (SURVEY §7 hard part 1); pass
      None for numpy's default (unstable) kind, which is what the reference literally calls (:39).
    junk: 'none' reproduces HEAD, where the removal rule is commented out (:54-55);
      'pid_cam' applies the commented rule  (g_pid == q_pid) & (g_camid == q_camid)  (:54),
      i.e. the classic Market-1501 protocol that is live in
      processor/processor_uniprompt_stage2.py:483-488.

    Returns dict(cmc float32[max_rank], mAP float64, first_hit int32[Q] (1-based rank of the first
    match among kept gallery entries, 0 = query has no match), ap float64[Q], num_rel int32[Q]).
    """
    distmat = np.asarray(distmat)
    q_pids, g_pids = np.asarray(q_pids), np.asarray(g_pids)
    q_camids, g_camids = np.asarray(q_camids), np.asarray(g_camids)
    num_q, num_g = distmat.shape
    if num_g < max_rank:  # :36-38
        max_rank = num_g
        if verbose:
            print("Note: number of gallery samples is quite small, got {}".format(num_g))
    order_all = np.argsort(distmat, axis=1) if sort_kind is None else np.argsort(distmat, axis=1, kind=sort_kind)
    hits_all = (g_pids[order_all] == q_pids[:, None]).astype(np.int32)  # :42

    first_hit = np.zeros(num_q, np.int32)
    ap = np.zeros(num_q, np.float64)
    num_rel = np.zeros(num_q, np.int32)
    cmc_rows = []
    for qi in range(num_q):
        order = order_all[qi]
        if junk == "pid_cam":
            drop = (g_pids[order] == q_pids[qi]) & (g_camids[order] == q_camids[qi])  # :54
            keep = np.invert(drop)
        else:
            keep = np.invert(False)  # :55-56 -> numpy True scalar, the row keeps a leading axis of 1
        row = hits_all[qi][keep]  # :60
        if not np.any(row):  # :61-63
            continue
        run = row.cumsum()  # :65 (flattens)
        first_hit[qi] = int(np.argmax(run > 0)) + 1
        capped = run.copy()
        capped[capped > 1] = 1  # :66
        cmc_rows.append(capped[:max_rank])  # :68
        rel = row.sum()  # :73
        prec = row.cumsum() / (np.arange(1, run.shape[0] + 1) * 1.0)  # :74-77
        ap[qi] = (np.asarray(prec) * row).sum() / rel  # :78-79 (numpy pairwise float64 sum)
        num_rel[qi] = rel
    n_valid = float(len(cmc_rows))
    assert n_valid > 0, "Error: all query identities do not appear in gallery"  # :82
    cmc = np.asarray(cmc_rows).astype(np.float32).sum(0) / n_valid  # :84-85
    mAP = np.mean([ap[i] for i in range(num_q) if num_rel[i] > 0])  # :86
    return dict(cmc=cmc, mAP=mAP, first_hit=first_hit, ap=ap, num_rel=num_rel)


def eval_func(distmat, q_pids, g_pids, q_camids, g_camids, max_rank=50, sort_kind="stable", junk="none"):
    """Reference signature (utils/metrics.py:28) -> (all_cmc, mAP)."""
    r = rank_eval(distmat, q_pids, g_pids, q_camids, g_camids, max_rank, sort_kind, junk, verbose=True)
    return r["cmc"], r["mAP"]


def clipstyle_eval(distmat, q_pids, g_pids, q_camids, g_camids, sort_kind="stable"):
    """processor/processor_uniprompt_stage2.py:471-509: junk removal always on, CMC summed in
    float64 over ALL gallery ranks and divided by len(q_pids), mAP = mean over ALL queries
    (queries without a match contribute 0)."""
    distmat = np.asarray(distmat)
    nq, ng = distmat.shape
    cmc = np.zeros(ng)
    ap = np.zeros(nq)
    for qi in range(nq):
        order = np.argsort(distmat[qi]) if sort_kind is None else np.argsort(distmat[qi], kind=sort_kind)
        drop = (g_pids[order] == q_pids[qi]) & (g_camids[order] == q_camids[qi])  # :484
        hits = (g_pids[order][~drop] == q_pids[qi]).astype(np.int32)  # :491
        pos = np.where(hits == 1)[0]
        if len(pos) == 0:
            continue
        step = np.zeros(ng)
        step[pos[0]:] = 1  # :496-497
        cmc += step
        prec = np.asarray([x / (i + 1.) for i, x in enumerate(hits.cumsum())]) * hits  # :502-504
        ap[qi] = prec.sum() / hits.sum()
    return cmc / nq, ap.mean()


# --------------------------------------------------------------------------------------
# k-reciprocal re-ranking
# --------------------------------------------------------------------------------------
def pairwise_sq_all(feat: np.ndarray) -> np.ndarray:
    """utils/reranking.py:36-41: the (Q+G)^2 squared-distance matrix of the stacked features."""
    return sq_euclidean(feat, feat)


def re_ranking_from_dist(distmat_all, query_num, k1, k2, lambda_value, sort_kind="stable", return_parts=False):
    """utils/reranking.py:45-99 on a given all-pairs matrix (what :41-44 produce).

    Rounding points follow the reference exactly: V / V_qe / the Jaccard accumulator are float16
    (:47,74,84,87), the kernel weights are fp32 exp / fp32 sum (:70-71), (1-lambda) multiplies in
    float16 (:95).
    """
    all_num = distmat_all.shape[0]
    # :46  column-max normalise, then transpose: row i holds d(., i) / max_k d(k, i)
    dn = np.transpose(distmat_all / np.max(distmat_all, axis=0))
    rank = (np.argsort(dn) if sort_kind is None else np.argsort(dn, kind=sort_kind)).astype(np.int32)  # :48
    V = np.zeros_like(dn).astype(np.float16)  # :47
    half = int(np.around(k1 / 2)) + 1  # :60 (banker's rounding)
    for i in range(all_num):
        fwd = rank[i, :k1 + 1]  # :53
        back = rank[fwd, :k1 + 1]  # :54
        recip = fwd[np.where(back == i)[0]]  # :55-56
        grown = recip
        for cand in recip:  # :58-67
            cf = rank[cand, :half]
            cb = rank[cf, :half]
            crec = cf[np.where(cb == cand)[0]]
            if len(np.intersect1d(crec, recip)) > 2 / 3 * len(crec):
                grown = np.append(grown, crec)
        grown = np.unique(grown)  # :69
        w = np.exp(-dn[i, grown])  # :70
        V[i, grown] = w / np.sum(w)  # :71
    V0 = V.copy() if return_parts else None
    dq = dn[:query_num, ]  # :72
    if k2 != 1:  # :73-78
        Vq = np.zeros_like(V, dtype=np.float16)
        for i in range(all_num):
            Vq[i, :] = np.mean(V[rank[i, :k2], :], axis=0)
        V = Vq
    inv = [np.where(V[:, j] != 0)[0] for j in range(all_num)]  # :80-82
    jac = np.zeros_like(dq, dtype=np.float16)  # :84
    for i in range(query_num):  # :86-93
        acc = np.zeros((1, all_num), dtype=np.float16)
        nz = np.where(V[i, :] != 0)[0]
        for k in nz:
            rows = inv[k]
            acc[0, rows] = acc[0, rows] + np.minimum(V[i, k], V[rows, k])
        jac[i] = 1 - acc / (2 - acc)
    final = jac * (1 - lambda_value) + dq * lambda_value  # :95
    out = final[:query_num, query_num:]  # :99
    if return_parts:
        return out, dict(dn=dn, rank=rank, V0=V0, V=V, jaccard=jac)
    return out


def re_ranking(probFea, galFea, k1, k2, lambda_value, local_distmat=None, only_local=False,
               sort_kind="stable", return_parts=False):
    """utils/reranking.py:29-100, reference signature (features as numpy or torch)."""
    p = probFea.numpy() if hasattr(probFea, "numpy") else np.asarray(probFea)
    g = galFea.numpy() if hasattr(galFea, "numpy") else np.asarray(galFea)
    query_num = p.shape[0]
    if only_local:  # :33-34
        dall = np.asarray(local_distmat)
    else:
        dall = pairwise_sq_all(np.concatenate([p, g], 0).astype(np.float32))
        if local_distmat is not None:  # :43-44
            dall = dall + local_distmat
    return re_ranking_from_dist(dall, query_num, k1, k2, lambda_value, sort_kind, return_parts)


# --------------------------------------------------------------------------------------
# evaluator object (utils/metrics.py:91-134), numpy-only restatement used as the CPU baseline
# --------------------------------------------------------------------------------------
class R1_mAP_eval:
    def __init__(self, num_query, max_rank=50, feat_norm=True, reranking=False, sort_kind="stable"):
        self.num_query, self.max_rank, self.feat_norm, self.reranking = num_query, max_rank, feat_norm, reranking
        self.sort_kind = sort_kind

    def reset(self):
        self.feats, self.pids, self.camids = [], [], []

    def update(self, output):
        feat, pid, camid = output
        self.feats.append(np.asarray(feat.cpu() if hasattr(feat, "cpu") else feat, dtype=np.float32))
        self.pids.extend(np.asarray(pid))
        self.camids.extend(np.asarray(camid))

    def compute(self):
        feats = np.concatenate(self.feats, 0)
        if self.feat_norm:
            feats = l2_normalize(feats)
        nq = self.num_query
        qf, gf = feats[:nq], feats[nq:]
        q_pids, g_pids = np.asarray(self.pids[:nq]), np.asarray(self.pids[nq:])
        q_cam, g_cam = np.asarray(self.camids[:nq]), np.asarray(self.camids[nq:])
        if self.reranking:
            distmat = re_ranking(qf, gf, k1=50, k2=15, lambda_value=0.3, sort_kind=self.sort_kind)  # :127
        else:
            distmat = sq_euclidean(qf, gf)  # :131
        cmc, mAP = eval_func(distmat, q_pids, g_pids, q_cam, g_cam, sort_kind=self.sort_kind)  # :132 (default max_rank)
        return cmc, mAP, distmat, self.pids, self.camids, qf, gf
