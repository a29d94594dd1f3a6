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
