"""CPU-only: the C-ABI library loads, exports everything include/mpreid_b200.h declares, and the
host-side logic (AP arithmetic shared with the kernels, final CMC/mAP reduction, lazy distmat,
evaluator bookkeeping) matches numpy / the oracle.  No compute calls that need a GPU."""
import ctypes
import os

import numpy as np
import pytest
import torch

from mp_reid_b200 import _lib as L
from mp_reid_b200 import engine as E
from oracle import mpreid_oracle as orc


def test_library_exports_every_declared_symbol():
    lib = L.load()
    declared = L.declared_symbols()
    assert len(declared) >= 13
    for name in declared:
        assert hasattr(lib, name), name
        assert name in L.SIGNATURES, f"{name} declared in the header but not bound in _lib.py"
    assert set(L.SIGNATURES) == set(declared)
    assert lib.mpreid_abi_version() == 1


def test_header_cites_reference_lines():
    text = open(L.HEADER).read()
    for cite in ["utils/metrics.py:7-13", "utils/metrics.py:15-25", "utils/metrics.py:28-88", "utils/reranking.py:29-100"]:
        assert cite in text


def test_workspace_queries_are_pure_host_calls():
    lib = L.load()
    assert lib.mpreid_rank_eval_workspace_bytes(3368, 15913, 1 << 18) > 0
    assert lib.mpreid_rerank_workspace_bytes(19281, 3368, 20, 6) > 0
    assert lib.mpreid_rerank_workspace_bytes(100, 10, 500, 6) == 0  # k1 out of range


def test_host_average_precision_bit_exact_vs_numpy():
    lib = L.load()
    rng = np.random.RandomState(0)
    for t in range(1500):
        n = int(rng.randint(1, 200000)) if t % 3 else int(rng.randint(1, 400))
        m = int(rng.randint(1, min(n, 80) + 1))
        pos = np.sort(rng.choice(n, m, replace=False))
        row = np.zeros((1, n), np.int32)
        row[0, pos] = 1
        tmp = row.cumsum() / (np.arange(1, n + 1) * 1.0)   # utils/metrics.py:74-77
        ref = (np.asarray(tmp) * row).sum() / row.sum()     # :78-79
        ranks = (pos + 1).astype(np.int32)
        got = lib.mpreid_host_average_precision(ranks.ctypes.data, m, n)
        assert got == ref, (n, m)


def test_host_average_precision_dense_runs_and_block_edges():
    """Runs of consecutive hits, hits on both sides of the 8/128-element block edges, every element a hit."""
    lib = L.load()
    rng = np.random.RandomState(5)
    for t in range(160):
        n = int(rng.randint(1, 1500000)) if t % 2 else int(rng.randint(1, 5000))
        mode = t % 4
        if mode == 0:
            pos = np.sort(rng.choice(n, int(rng.randint(1, min(n, 3000) + 1)), replace=False))
        elif mode == 1:
            m = int(rng.randint(1, min(n, 500) + 1)); st = int(rng.randint(0, n - m + 1)); pos = np.arange(st, st + m)
        elif mode == 2:
            c = rng.choice(n, min(n, int(rng.randint(1, 200))), replace=False)
            pos = np.unique(np.clip(np.concatenate([c, c + 1, c + 7, c + 8, c + 127, c + 128]), 0, n - 1))
        else:
            pos = np.arange(n) if n < 20000 else np.sort(rng.choice(n, 20000, replace=False))
        row = np.zeros(n, np.int32); row[pos] = 1
        tmp = row.cumsum() / (np.arange(1, n + 1) * 1.0)
        ref = (tmp * row).sum() / row.sum()
        ranks = (pos + 1).astype(np.int32)
        assert lib.mpreid_host_average_precision(ranks.ctypes.data, len(pos), n) == ref, (n, mode)


def test_host_order_keys_sort_like_numpy_stable():
    lib = L.load()
    rng = np.random.RandomState(1)
    v = rng.randn(5000).astype(np.float32)
    v[::7] = v[3]  # ties
    v[10] = 0.0; v[11] = -0.0; v[12] = np.inf; v[13] = -np.inf; v[14] = np.nan; v[15] = -np.nan
    keys = np.zeros(v.shape, np.uint32)
    lib.mpreid_host_order_keys(v.ctypes.data, v.size, keys.ctypes.data)
    mine = np.argsort((keys.astype(np.uint64) << np.uint64(32)) | np.arange(v.size, dtype=np.uint64), kind="stable")
    assert np.array_equal(mine, np.argsort(v, kind="stable"))


@pytest.mark.parametrize("name", ["small_eval", "ties_eval", "no_match", "cctv_small"])
def test_reduce_cmc_map_equals_reference_reduction(golden_dir, name):
    g = dict(np.load(os.path.join(golden_dir, name + ".npz")))
    r = orc.rank_eval(g["dist_euclid"], g["q_pid"], g["g_pid"], g["q_cam"], g["g_cam"])
    cmc, mAP = E.reduce_cmc_map(r["first_hit"], r["ap"], r["num_rel"], 50, g["dist_euclid"].shape[1])
    assert cmc.dtype == np.float32 and np.array_equal(cmc, g["ref_stable_cmc"])
    assert mAP == g["ref_stable_mAP"]
    # CLIP-style denominators (processor_uniprompt_stage2.py:508-509)
    rj = orc.rank_eval(g["dist_1mcos"], g["q_pid"], g["g_pid"], g["q_cam"], g["g_cam"], junk="pid_cam")
    cmc2, mAP2 = E.reduce_cmc_map(rj["first_hit"], rj["ap"], rj["num_rel"], 50, g["dist_1mcos"].shape[1], "all")
    assert np.array_equal(cmc2, g["ref_clip_cmc"]) and mAP2 == g["ref_clip_mAP"]


def test_reduce_raises_like_reference_when_no_query_is_valid():
    with pytest.raises(AssertionError, match="all query identities do not appear in gallery"):
        E.reduce_cmc_map(np.zeros(3, np.int32), np.zeros(3), np.zeros(3, np.int32), 50, 100)


def test_compute_path_fails_loudly_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mp_reid_b200 import metrics
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        metrics.euclidean_distance(torch.randn(4, 8), torch.randn(5, 8))


def test_evaluator_surface_matches_reference():
    import inspect
    from mp_reid_b200 import metrics, reranking
    sig = inspect.signature(metrics.R1_mAP_eval.__init__)
    names = list(sig.parameters)[:5]
    assert names == ["self", "num_query", "max_rank", "feat_norm", "reranking"]
    assert sig.parameters["max_rank"].default == 50 and sig.parameters["feat_norm"].default is True
    assert list(inspect.signature(metrics.eval_func).parameters)[:6] == ["distmat", "q_pids", "g_pids", "q_camids", "g_camids", "max_rank"]
    assert list(inspect.signature(reranking.re_ranking).parameters)[:7] == [
        "probFea", "galFea", "k1", "k2", "lambda_value", "local_distmat", "only_local"]
    ev = metrics.R1_mAP_eval(10)
    with pytest.raises(AttributeError):  # update before reset, as in the reference (lists undefined)
        ev.update((torch.zeros(2, 4), (1, 2), (0, 0)))


def test_merge_adjacent_views_is_copy_free_and_exact():
    """Pieces of one upload are neighbouring row views of one allocation: they merge into one view (no copy);
    views from different allocations, gaps or reordered pieces stay separate."""
    import torch
    from mp_reid_b200.metrics import _merge_adjacent
    t = torch.arange(60.).reshape(15, 4)
    u = torch.zeros(3, 4)
    out = _merge_adjacent([t[0:3], t[3:5], t[5:9], u, t[9:12], t[13:15], t[12:13]])
    assert [tuple(o.shape) for o in out] == [(9, 4), (3, 4), (3, 4), (2, 4), (1, 4)]
    assert out[0].data_ptr() == t.data_ptr() and torch.equal(out[0], t[0:9])      # a view, not a copy
    assert torch.equal(torch.cat(out), torch.cat([t[0:9], u, t[9:12], t[13:15], t[12:13]]))
    cols = t[:, :2]                                                                 # column slices are not contiguous rows
    assert len(_merge_adjacent([cols[0:2], cols[2:4]])) == 2
    assert _merge_adjacent([]) == []


def test_take_rows_splits_pending_views_exactly():
    import torch
    from mp_reid_b200.metrics import _take_rows
    t = torch.arange(80.).reshape(20, 4)
    u = torch.ones(5, 4)
    for cut in (0, 1, 7, 8, 12, 13, 20, 25):
        pend = [t[0:8], t[8:12], u, t[12:20]]
        take, rest = _take_rows(pend, cut)
        whole = torch.cat(pend)
        got = torch.cat(take) if take else whole[:0]
        left = torch.cat(rest) if rest else whole[:0]
        assert torch.equal(got, whole[:cut]) and torch.equal(left, whole[cut:]), cut
    take, rest = _take_rows([t[0:8], t[8:12]], 12)
    assert len(take) == 1 and take[0].data_ptr() == t.data_ptr() and rest == []
