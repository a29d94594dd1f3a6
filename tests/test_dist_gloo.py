"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: sharding, gallery broadcast, the single
all-gather of per-query results and the host reduction.  The per-query inputs come from the oracle;
the N>1 result must be bit-identical to the single-process reduction."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, golden, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mp_reid_b200 import distributed as D
    from oracle import mpreid_oracle as orc
    g = dict(np.load(golden))
    Q = len(g["q_pid"])
    counts = [D.shard_bounds(Q, world, r)[1] - D.shard_bounds(Q, world, r)[0] for r in range(world)]
    lo, hi = D.shard_bounds(Q, world, rank)
    # gallery broadcast from rank 0
    gf_src = torch.from_numpy(g["gf"]) if rank == 0 else None
    gf, g_pid, g_cam = D.broadcast_gallery(gf_src, g["g_pid"] if rank == 0 else None, g["g_cam"] if rank == 0 else None,
                                           src=0, device="cpu")
    assert torch.equal(gf, torch.from_numpy(g["gf"])) and np.array_equal(g_pid, g["g_pid"])
    # each rank evaluates ITS query rows (oracle stands in for the CUDA kernels on this CPU box)
    r = orc.rank_eval(g["dist_euclid"][lo:hi], g["q_pid"][lo:hi], g_pid, g["q_cam"][lo:hi], g_cam) if hi > lo else None
    fh = torch.from_numpy(r["first_hit"]); ap = torch.from_numpy(r["ap"]); nr = torch.from_numpy(r["num_rel"])
    cmc, mAP = D.sharded_reduce(fh, ap, nr, counts, 50, g["dist_euclid"].shape[1])
    # block-cyclic ownership (the fused row-sharded re-ranking): results travel with their global query index; with fewer
    # than 256 queries rank 1 owns none and still takes part in the collective
    full = orc.rank_eval(g["dist_euclid"], g["q_pid"], g_pid, g["q_cam"], g_cam)
    for blk in (256, 16):     # 16: pretend blocks of 16 rows, so that both ranks own some
        ids = torch.arange(Q)
        ids = ids[(ids // blk) % world == rank]
        if blk == 256:
            assert torch.equal(ids, D.rerank_owned_queries(Q, world, rank))
        sel = ids.numpy()
        cmc2, mAP2 = D.sharded_reduce(torch.from_numpy(full["first_hit"][sel]), torch.from_numpy(full["ap"][sel]),
                                      torch.from_numpy(full["num_rel"][sel]), None, 50, g["dist_euclid"].shape[1], ids=ids, total=Q)
        assert np.array_equal(cmc2, cmc) and mAP2 == mAP
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), cmc=cmc, mAP=mAP)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["small_eval", "no_match"])
def test_two_rank_sharded_eval_is_bit_identical(golden_dir, tmp_path, name):
    golden = os.path.join(golden_dir, name + ".npz")
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), golden, str(tmp_path)), nprocs=world, join=True)
    g = dict(np.load(golden))
    for r in range(world):
        out = np.load(tmp_path / f"r{r}.npz")
        assert np.array_equal(out["cmc"], g["ref_stable_cmc"])
        assert out["mAP"] == g["ref_stable_mAP"]


def test_shard_bounds_cover_and_balance():
    from mp_reid_b200.distributed import shard_bounds
    for n in [0, 1, 7, 11659, 100000]:
        for w in [1, 2, 3, 8]:
            b = [shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def test_dropin_seeds_only_the_two_hot_path_modules(tmp_path):
    """A fake reference tree: utils/{__init__,meter}.py + a script importing like processor/processor.py:6-7."""
    import subprocess
    (tmp_path / "utils").mkdir()
    (tmp_path / "utils" / "__init__.py").write_text("")
    (tmp_path / "utils" / "meter.py").write_text("class AverageMeter:\n    pass\n")
    (tmp_path / "utils" / "metrics.py").write_text("raise RuntimeError('the reference module must not be imported')\n")
    (tmp_path / "script.py").write_text(
        "from utils.meter import AverageMeter\nfrom utils.metrics import R1_mAP_eval\nfrom utils.reranking import re_ranking\n"
        "import utils.metrics as m\nprint('META', m.__name__, AverageMeter.__module__, R1_mAP_eval.__module__, re_ranking.__module__)\n")
    out = subprocess.run([sys.executable, "-m", "mp_reid_b200.dropin", str(tmp_path / "script.py")], cwd=ROOT,
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    assert "META mp_reid_b200.metrics utils.meter mp_reid_b200.metrics mp_reid_b200.reranking" in out.stdout


def test_dropin_can_also_serve_the_loss_modules(tmp_path):
    """MPREID_DROPIN_LOSSES=1: `from .triplet_loss import TripletLoss` inside the reference's loss package (loss/make_loss.py:7)
    and `from loss.supcontrast import SupConLoss` (processor/processor_uniprompt_stage1.py:9) resolve to this package."""
    import subprocess
    (tmp_path / "loss").mkdir()
    (tmp_path / "loss" / "__init__.py").write_text("from .make_loss import make_loss\n")
    (tmp_path / "loss" / "make_loss.py").write_text("from .triplet_loss import TripletLoss\n\ndef make_loss():\n    return TripletLoss(0.3)\n")
    (tmp_path / "loss" / "triplet_loss.py").write_text("raise RuntimeError('the reference module must not be imported')\n")
    (tmp_path / "loss" / "supcontrast.py").write_text("raise RuntimeError('the reference module must not be imported')\n")
    (tmp_path / "script.py").write_text(
        "from loss import make_loss\nfrom loss.supcontrast import SupConLoss\n"
        "print('META', type(make_loss()).__module__, make_loss().margin, SupConLoss.__module__)\n")
    out = subprocess.run([sys.executable, "-m", "mp_reid_b200.dropin", str(tmp_path / "script.py")], cwd=ROOT,
                         capture_output=True, text=True, timeout=300, env=dict(os.environ, MPREID_DROPIN_LOSSES="1"))
    assert out.returncode == 0, out.stderr
    assert "META mp_reid_b200.triplet 0.3 mp_reid_b200.supcon" in out.stdout


def test_aligned_shard_bounds_cover_and_align():
    from mp_reid_b200.distributed import aligned_shard_bounds
    for n, w in [(82161, 8), (82161, 2), (100, 8), (0, 4), (33, 2), (9003, 2), (15913, 4)]:
        b = [aligned_shard_bounds(n, w, r) for r in range(w)]
        assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        assert all(lo % 32 == 0 or lo == n for lo, _ in b)


# ------------------------------------------------------------------------------------------------
# Collective sequence of the fused row-sharded re-ranking (distributed._rerank_sharded_fused): every rank must issue the
# same exchanges in the same order, also a rank that owns no 256-row query block (fewer than world * 256 queries).
# The CUDA calls are replaced by shape-only stand-ins, the collectives are the real ones (gloo).
class _FakePrep:
    def __init__(self, n, d):
        self.n, self.sqnorm, self.xn = n, torch.ones(n), torch.zeros(n, d)

    def take(self, ids):
        return _FakePrep(int(ids.numel()), self.xn.shape[1])

    def rows(self, lo, hi):
        return _FakePrep(hi - lo, self.xn.shape[1])


class _FakeEngine:
    """Stand-in for mp_reid_b200.engine: tensors of the right shape and dtype, a log of the finish calls."""

    def __init__(self, rank):
        self.rank, self.finish_calls, self.K, self.C0, self.C1 = rank, [], 27, 40, 96

    def rerank_neighbor_count(self, k1, k2):
        return self.K

    def mark(self, name):
        pass

    def dist_matrix(self, a, b, metric, precision):
        return torch.zeros(a.n, b.n)

    def row_kth(self, d, t, bound=False):
        return torch.ones(d.shape[0])

    def dist_symmetric_topk(self, x, thr, cap, nq, precision, own_mod=1, own_rank=0):
        assert thr.shape == (x.n,) and (own_mod, own_rank) == (dist.get_world_size(), self.rank)
        return (torch.zeros((x.n, cap), dtype=torch.int64), torch.zeros(x.n, dtype=torch.int32), torch.zeros(nq, x.n - nq), 0,
                torch.full((x.n,), 1.0 + self.rank))

    def cand_topk(self, cand, cnt, k, row_scale, thr, partial=False):
        assert partial and bool((row_scale == float(dist.get_world_size())).all())    # maxima were max-reduced over the ranks
        return torch.zeros((cand.shape[0], k), dtype=torch.int64), torch.zeros(4, dtype=torch.int32)

    def merge_topk(self, keys_all, row_scale, thr):
        P, n, k = keys_all.shape
        return torch.zeros((n, k), dtype=torch.int32), torch.zeros((n, k)), torch.zeros(4, dtype=torch.int32)

    def rerank_build_v0_sparse(self, row_ids, R, N, k1, nbr, nbr_val, row_max_rows, xn, sqnorm):
        return (torch.zeros((R, self.C0), dtype=torch.int32), torch.zeros((R, self.C0), dtype=torch.float16),
                torch.full((R,), 9 + self.rank, dtype=torch.int32))

    def rerank_finish_workspace(self, N, Q, k1, k2, device):
        self.v = (torch.zeros((N, self.C1), dtype=torch.int32), torch.zeros((N, self.C1), dtype=torch.float16), torch.zeros(N, dtype=torch.int32))
        return torch.zeros(1, dtype=torch.uint8)

    def rerank_finish_v_views(self, ws, N, Q, k1, k2):
        return self.v

    def alloc_dist(self, Q, G, device):
        return torch.zeros(Q, G)

    def rerank_finish(self, nbr, v0, block, q_ids, row_max, N, Q, k1, k2, lam, out=None, block_col0=None, rows_global=False, stages=7,
                      ws=None, qe_rows=(0, 0)):
        assert q_ids.numel() >= 1 and v0[0].shape[0] == N and v0[2].shape == (N,)      # the C entry rejects an empty query list
        self.finish_calls.append((stages, int(q_ids.numel()), tuple(qe_rows)))
        if stages == 8:
            self.v[2][qe_rows[0]: qe_rows[1]] = 20 + self.rank                        # lengths of the rows this rank expanded
            self.v[0][qe_rows[0]: qe_rows[1], :21] = self.rank + 1
        return out if out is not None else torch.zeros(int(q_ids.numel()), N - Q)


def _fused_worker(rank, world, port, nq, out_dir):
    import datetime
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world, timeout=datetime.timedelta(seconds=60))
    from mp_reid_b200 import distributed as D
    N = 3001
    log = {}
    for k2 in (6, 1):
        fake = _FakeEngine(rank)
        D.E = fake
        final, q_ids = D._rerank_sharded_fused(_FakePrep(N, 8), nq, 20, k2, 0.3, None, None, world, rank)
        own = D.rerank_owned_queries(nq, world, rank)
        assert torch.equal(q_ids, own) and tuple(final.shape) == (int(own.numel()), N - nq)
        if k2 != 1:   # every rank holds every rank's expanded rows, trimmed to the longest one
            lens = torch.cat([torch.full((D.shard_bounds(N, world, r)[1] - D.shard_bounds(N, world, r)[0],), 20 + r) for r in range(world)])
            assert torch.equal(fake.v[2], lens.to(torch.int32))
            assert bool((fake.v[0][:, :21] == (lens - 19).to(torch.int32)[:, None]).all()) and bool((fake.v[0][:, 21:] == 0).all())
        log[k2] = fake.finish_calls
    np.save(os.path.join(out_dir, f"calls{rank}.npy"), np.array([repr(log)]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("nq", [200, 700])
def test_fused_sharded_rerank_same_collectives_on_every_rank(tmp_path, nq):
    """nq = 200: rank 1 owns no query block -- it must still expand its V rows and join every exchange (a rank that returned
    early once left the others waiting in the all-reduce).  nq = 700: blocks 0, 2 on rank 0 and block 1 on rank 1."""
    world = 2
    mp.spawn(_fused_worker, args=(world, _free_port(), nq, str(tmp_path)), nprocs=world, join=True)
    calls = [eval(str(np.load(tmp_path / f"calls{r}.npy")[0])) for r in range(world)]
    own = [nq if nq <= 256 else 256 + (nq - 512), 0 if nq <= 256 else 256]
    for r in range(world):
        lo = r * 1501 if r else 0
        hi = 1501 if r == 0 else 3001
        if own[r]:
            assert calls[r][6] == [(8, own[r], (lo, hi)), (16 | 4 | 2, own[r], (0, 0))]
            assert calls[r][1] == [(7, own[r], (0, 0))]
        else:     # idle rank: the expansion of its own rows only, with the placeholder query row
            assert calls[r][6] == [(8, 1, (lo, hi))] and calls[r][1] == []
