"""torchrun check of the multi-rank paths: distributed.sharded_evaluator (ragged query / gallery shards fed per rank
from host memory) and distributed.rerank_sharded (row-sharded k-reciprocal re-ranking) must give what ONE GPU gives,
bit for bit.  Prints 'SHARDED_EVAL_OK <world>' on rank 0.

    MPREID_CHECK_BACKEND=nccl|gloo   (default nccl; gloo stages device tensors through the host)
    MPREID_CHECK_ONE_DEVICE=1        every rank uses cuda:0 (two ranks on a one-GPU box; needs gloo: NCCL refuses
                                     two ranks on one device)
"""
import contextlib, datetime, io, os, sys
import numpy as np, torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mp_reid_b200 import engine as E, metrics, distributed as MD
from mp_reid_b200.reranking import _rerank_device

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
backend = os.environ.get("MPREID_CHECK_BACKEND", "nccl")
if os.environ.get("MPREID_CHECK_ONE_DEVICE", "0") == "1":
    local = 0
torch.cuda.set_device(local)
os.environ["MPREID_DEVICE"] = f"cuda:{local}"
# a rank that misses a collective must fail the check in minutes, not hold the GPUs for the default 10-30 minutes
limit = datetime.timedelta(seconds=int(os.environ.get("MPREID_CHECK_TIMEOUT_S", "300")))
if backend == "nccl":
    dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=limit)
else:
    dist.init_process_group(backend, timeout=limit)
rng = np.random.RandomState(21)
Q, G, D = 1501, 9003, 320
x = torch.from_numpy(rng.randn(Q + G, D).astype(np.float32))
pid = rng.randint(0, 120, Q + G); cam = rng.randint(0, 6, Q + G)
q_lo, q_hi = MD.shard_bounds(Q, world, rank)
g_lo, g_hi = MD.aligned_shard_bounds(G, world, rank)
for junk in ("none", "pid_cam"):
    ev = MD.sharded_evaluator(q_hi - q_lo, junk=junk); ev.reset()
    for s in range(q_lo, q_hi, 400):
        e = min(q_hi, s + 400); ev.update((x[s:e].pin_memory(), pid[s:e], cam[s:e]))
    for s in range(Q + g_lo, Q + g_hi, 1000):
        e = min(Q + g_hi, s + 1000); ev.update((x[s:e], pid[s:e], cam[s:e]))
    with contextlib.redirect_stdout(io.StringIO()):
        cmc, mAP, dmat, *_, qf, gf = ev.compute()
    ref = metrics.R1_mAP_eval(Q, junk=junk); ref.reset(); ref.update((x, pid, cam))
    with contextlib.redirect_stdout(io.StringIO()):
        cmc0, mAP0, d0, *_, qf0, gf0 = ref.compute()
    assert np.array_equal(cmc, cmc0) and mAP == mAP0, (rank, junk, mAP, mAP0)
    assert np.array_equal(np.asarray(dmat), np.asarray(d0)[q_lo:q_hi]) and torch.equal(gf, gf0) and torch.equal(qf, qf0[q_lo:q_hi])

# ---- row-sharded re-ranking: this rank's query rows of final_dist == the same rows of the one-GPU result
dev = torch.device("cuda", local)
centers = rng.randn(40, 96).astype(np.float32)
NS = 2301          # odd: the row shards are ragged
lab = rng.randint(0, 40, NS)
feats = torch.from_numpy(centers[lab] + 1.3 * rng.randn(NS, 96).astype(np.float32)).to(dev)
nq = 301
prep = E.prep_rows(feats, normalize=True, keep_xn=True)   # the fused pipelines read feature rows
for fused in ("1", "0"):          # the fused (no N x N matrix) pipeline and the materialising one
    os.environ["MPREID_RERANK_FUSED"] = fused
    for (k1, k2, lam) in [(20, 6, 0.3), (7, 1, 0.5)]:
        want = _rerank_device(prep, nq, k1, k2, lam)
        got, ids = MD.rerank_sharded(prep, nq, k1, k2, lam)
        if fused == "1":
            assert torch.equal(ids, MD.rerank_owned_queries(nq, world, rank, dev))
        else:
            lo, hi = MD.shard_bounds(nq, world, rank)
            assert torch.equal(ids.cpu(), torch.arange(lo, hi))
        assert torch.equal(got, want[ids]), (rank, fused, k1, k2, float((got - want[ids]).abs().max()))
        # per-query results gathered into global query order == one-GPU evaluation of the whole matrix
        q_pid = torch.from_numpy(lab[:nq]).to(dev); g_pid = torch.from_numpy(lab[nq:]).to(dev)
        if ids.numel():
            fh, ap, nr = E.rank_eval(got, q_pid[ids], g_pid)
        else:   # more ranks than 256-row query blocks: this rank finishes no query row
            fh, ap, nr = (torch.zeros(0, dtype=torch.int32, device=dev), torch.zeros(0, dtype=torch.float64, device=dev),
                          torch.zeros(0, dtype=torch.int32, device=dev))
        cmc, mAP = MD.sharded_reduce(fh, ap, nr, None, 50, NS - nq, ids=ids, total=nq)
        fh0, ap0, nr0 = E.rank_eval(want, q_pid, g_pid)
        cmc0, mAP0 = E.reduce_cmc_map(fh0.cpu().numpy(), ap0.cpu().numpy(), nr0.cpu().numpy(), 50, NS - nq)
        assert mAP == mAP0 and np.array_equal(cmc, cmc0)
os.environ.pop("MPREID_RERANK_FUSED")
# ---- the C-ABI communicator (mpreid_comm_*): the three collectives of the path between the ranks, without torch.distributed
if backend == "nccl":
    import ctypes
    from mp_reid_b200 import _lib as L
    lib = L.load()
    uid = ctypes.create_string_buffer(128)
    if rank == 0:
        L.check(lib.mpreid_comm_unique_id(uid), "comm_unique_id")
    box = [uid.raw]
    dist.broadcast_object_list(box, src=0)          # the 128 bytes travel out of band
    uid = ctypes.create_string_buffer(box[0], 128)
    comm = ctypes.c_void_p()
    L.check(lib.mpreid_comm_init(ctypes.byref(comm), world, rank, uid), "comm_init")
    st = torch.cuda.current_stream().cuda_stream
    g = torch.full((4096,), float(rank + 1), device=dev) if rank == 0 else torch.zeros(4096, device=dev)
    L.check(lib.mpreid_comm_broadcast(comm, g.data_ptr(), g.numel() * 4, 0, st), "comm_broadcast")
    mine = torch.full((512,), float(rank), device=dev)
    allv = torch.empty((world, 512), device=dev)
    L.check(lib.mpreid_comm_allgather(comm, mine.data_ptr(), allv.data_ptr(), 512 * 4, st), "comm_allgather")
    mx = torch.arange(100, dtype=torch.float32, device=dev) * (1.0 if rank == 0 else -1.0)
    L.check(lib.mpreid_comm_allreduce_max_f32(comm, mx.data_ptr(), 100, st), "comm_allreduce_max_f32")
    torch.cuda.synchronize()
    assert bool((g == 1.0).all()) and all(bool((allv[r] == float(r)).all()) for r in range(world))
    assert torch.equal(mx, torch.arange(100, dtype=torch.float32, device=dev))
    L.check(lib.mpreid_comm_destroy(comm), "comm_destroy")
dist.barrier()
if rank == 0:
    print("SHARDED_EVAL_OK", world)
dist.destroy_process_group()
