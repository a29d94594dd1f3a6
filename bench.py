#!/usr/bin/env python
"""bench.py — query x gallery pairs/s through distance + ranking + CMC/mAP on the MSMT17-shaped set
(BASELINE.json: 11,659 x 82,161 x 1280-d), plus re-rank ms, on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload msmt17|market|cctv|c5] [--scaling weak|strong]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...      (one rank per GPU, NCCL)

A "step" is one pass of the hot path over one batch of synthetic features:
  value : features already resident in HBM -> normalise + operand planes -> distance matrix ->
          rank / CMC / AP kernels -> per-query results to the host -> (cmc, mAP)
  e2e   : the same through the reference-facing API (R1_mAP_eval.reset/update/compute) fed from
          PINNED HOST batches, host->device copies and the result read-back inside the timed region
Workloads (BASELINE.json configs): msmt17 (the metric's named config, default), market (config 1/3), cctv
(config 2: arccos-cosine distance, pid+camera junk rule, bf16 delta reported), c5 (config 5: 100k x 1M x 768
top-100 retrieval, gallery broadcast from rank 0 inside the e2e region).
Multi-GPU: query rows are sharded.  --scaling weak (default): every rank evaluates its own MSMT17-sized query shard
against the replicated gallery; --scaling strong: the ONE named problem, its query rows split over the ranks
(gallery preparation repeated on every rank).  Per-query results are gathered with one collective, rank 0 reduces.
The reference arm (--impl reference) times the reference's CPU algorithm (oracle port: torch-CPU sgemm + numpy
argsort + Python loop; the Python reference itself cannot travel to the GPU box) on a bounded sample with all
host threads.
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from mp_reid_b200 import synth  # noqa: E402

METRIC_NAME = "query x gallery pairs/sec (dist+rank+mAP)"


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], bf16=d["bf16_tflops"], bf16_sustained=d["bf16_tflops_sustained"], source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


# ------------------------------------------------------------------------------------------ reference arm
def cpu_reference_pass(qf, gf, q_pid, g_pid, q_cam, g_cam, metric="sqeuclid", junk="none"):
    """One pass of the reference's CPU algorithm (oracle port) -> (cmc, mAP)."""
    from oracle import mpreid_oracle as orc
    feats = orc.l2_normalize(np.concatenate([qf, gf]))
    q, g = feats[: len(qf)], feats[len(qf):]
    d = orc.arccos_cosine(q, g) if metric == "arccos" else (orc.one_minus_cosine(q, g) if metric == "one_minus_dot" else orc.sq_euclidean(q, g))
    # numpy's default sort, as the reference calls it (the junk rule needs the stable kind the goldens were made with)
    return orc.eval_func(d, q_pid, g_pid, q_cam, g_cam, sort_kind=None if junk == "none" else "stable", junk=junk)


def cpu_sample(data, n_queries):
    qf, gf, q_pid, g_pid, q_cam, g_cam = data
    return qf[:n_queries].numpy(), gf.numpy(), q_pid[:n_queries], g_pid, q_cam[:n_queries], g_cam


def cpu_rerank_sample(data, n_all, k1, k2):
    """The reference's re_ranking (oracle port of utils/reranking.py:29-100) on a stated sub-shape -> dict."""
    from oracle import mpreid_oracle as orc
    qf, gf, q_pid, g_pid, q_cam, g_cam = data
    Q, G = qf.shape[0], gf.shape[0]
    nq = max(8, min(Q, int(round(n_all * Q / (Q + G)))))
    ng = max(k1 + 2, min(G, n_all - nq))
    feats = orc.l2_normalize(np.concatenate([qf[:nq].numpy(), gf[:ng].numpy()]))
    t0 = time.perf_counter()
    fd = orc.re_ranking(feats[:nq], feats[nq:], k1, k2, 0.3, sort_kind=None)
    dt = time.perf_counter() - t0
    return {"ms": 1e3 * dt, "Q": nq, "G": ng, "N": nq + ng, "k1": k1, "k2": k2, "cores": os.cpu_count(), "kind": "port",
            "note": "oracle port of utils/reranking.py on the first rows of the same synthetic set; the reference's cost grows "
                    "~N^2 (N x N float32/int64/fp16 temporaries: 175 GB at the MSMT17 shape, infeasible on this host)",
            "final_sum": float(fd.astype(np.float64).sum())}


def run_reference(args, data, workload):
    if env_int("RANK", 0) != 0:
        return
    torch.set_num_threads(os.cpu_count())
    G = data[1].shape[0]
    nq = max(8, min(data[0].shape[0], int(args.ref_queries)))
    sample = cpu_sample(data, nq)
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        cmc, mAP = quiet(cpu_reference_pass, *sample, metric=args.metric, junk=args.junk)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = nq * G * len(times) / total
    line = {
        "impl": "reference", "metric": METRIC_NAME, "value": value, "unit": "pairs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload, "Q_sample": nq, "G": G, "D": int(data[0].shape[1]), "distance": args.metric,
                   "feat_norm": True, "junk": args.junk},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": os.cpu_count(), "kind": "port",
                         "sample": f"first {nq} queries x full gallery per step (torch-CPU sgemm, numpy argsort, Python loop)"},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "mAP_sample": float(mAP),
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ distributed plumbing
def init_dist(dev):
    import torch.distributed as dist
    # keep stdout to the ONE JSON line: NCCL prints its version banner to fd 1 when the communicator is created,
    # so fd 1 points at stderr until the first collective has run
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        dist.init_process_group("nccl", device_id=dev)
        warm = torch.zeros(1, device=dev)
        dist.all_reduce(warm)
        torch.cuda.synchronize()
    finally:
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)
    return dist


def max_over_ranks(ms, dist, dev):
    if dist is None:
        return ms
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def kernel_ms(fn, reps):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def dtype_name(prec):
    return {"bf16": "bf16", "3xtf32": "f32 (3xTF32 tensor-core split)", "2xfp16": "f32 (2xFP16 fast split)"}.get(prec, "f32 (3xFP16 scaled tensor-core split)")


def gemm_peak(prec, pk, dev):
    if prec == "bf16":
        return pk["bf16"], f"{pk['source']} cuBLAS bf16 burst"
    if prec == "2xfp16":
        return pk["bf16"] / 2.0, f"{pk['source']} cuBLAS bf16 burst {pk['bf16']:.0f} TFLOP/s / 2 MMAs per product (fast mode)"
    if prec in ("3xfp16", "fp32"):
        # fp32-accurate mode on the fp16 pipe: 3 MMAs per product -> denominator = dense 16-bit peak / 3
        return pk["bf16"] / 3.0, (f"{pk['source']} cuBLAS bf16 burst {pk['bf16']:.0f} TFLOP/s (same pipe and rate as fp16) / 3 MMAs per product; "
                                  f"sustained figure {pk['bf16_sustained']:.0f}/3 = {pk['bf16_sustained'] / 3:.0f}")
    # 3xTF32: denominator = dense TF32 peak / 3, TF32 peak measured here
    torch.backends.cuda.matmul.allow_tf32 = True
    a = torch.randn(8192, 8192, device=dev); b = torch.randn(8192, 8192, device=dev)
    t_tf32 = min(kernel_ms(lambda: torch.matmul(a, b), 5) for _ in range(3))
    tf32_peak = 2 * 8192 ** 3 / (t_tf32 * 1e-3) / 1e12
    return tf32_peak / 3.0, f"cuBLAS TF32 8192^3 measured in this run ({tf32_peak:.0f} TFLOP/s) / 3 MMAs per product"


def traffic_of(kernel, workload_key, prec):
    """DRAM bytes per launch of `kernel` from this round's `ncu --set full` capture, recorded in profiles/r02/traffic.json
    together with the .ncu-rep summary it was read from -> (bytes or None, source description)."""
    p = os.path.join(ROOT, "profiles", "r02", "traffic.json")
    if os.path.exists(p):
        rec = json.load(open(p)).get(f"{kernel}:{workload_key}:{prec}")
        if rec:
            return rec["dram_bytes_read"] + rec["dram_bytes_write"], "profiles/r02/traffic.json <- " + rec.get("source", "?")
    return None, None


# ------------------------------------------------------------------------------------------ our arm: evaluator workloads
def run_ours(args, data, workload, wkey):
    from mp_reid_b200 import engine as E
    from mp_reid_b200 import metrics
    from mp_reid_b200 import distributed as MD
    from mp_reid_b200.reranking import _rerank_device

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    distributed = world > 1
    dist = init_dist(dev) if distributed else None
    qf, gf, q_pid, g_pid, q_cam, g_cam = data
    Qall, G, D = qf.shape[0], gf.shape[0], qf.shape[1]
    prec, junk, metric = args.precision, args.junk, args.metric
    strong = distributed and args.scaling == "strong"
    qf_all, q_pid_all, q_cam_all = qf, q_pid, q_cam
    if strong:   # the ONE named problem: this rank's slice of its query rows
        q_lo, q_hi = MD.shard_bounds(Qall, world, rank)
        qf, q_pid, q_cam = qf[q_lo:q_hi], q_pid[q_lo:q_hi], q_cam[q_lo:q_hi]
    Q = qf.shape[0]                                   # queries of THIS rank per step
    Qmax = (Qall + world - 1) // world if strong else Q
    total_pairs = (Qall if strong else world * Q) * G  # pairs per step over all ranks

    # ---- inputs resident in HBM (device-timed `value`) and in pinned host memory (`e2e`)
    feats_dev = torch.cat([qf, gf]).to(dev)
    lab = dict(q_pid=torch.from_numpy(q_pid).to(dev), g_pid=torch.from_numpy(g_pid).to(dev),
               q_cam=torch.from_numpy(q_cam).to(dev), g_cam=torch.from_numpy(g_cam).to(dev))
    batch = 8192
    dist_buf = E.alloc_dist(Q, G, dev)
    launches = [0]
    n_slots = max(args.steps, args.warmup)
    # per-step result slots: the packed per-query buffer of every timed step.  N = 1: copied to a pinned host slot as
    # soon as the step's kernels are queued and reduced by numpy one step later (while the GPU runs the next step).
    # N > 1: the slots stay on the device and travel with ONE all-gather after the last step -- no collective and no
    # host synchronisation between the steps, so a rank never waits for a slower one inside the loop.
    slot_bytes = 16 * Qmax + 16
    dev_slots = torch.zeros((n_slots, slot_bytes), dtype=torch.uint8, device=dev)
    host_slots = torch.empty((n_slots, slot_bytes), dtype=torch.uint8, pin_memory=True) if not distributed else None
    copy_done = [torch.cuda.Event() for _ in range(n_slots)]
    stage_events = []   # per step: events around prep | distance GEMM | rank/AP kernels, on the launching stream

    def unpack(h, nq):
        """packed RankResult bytes of one rank -> (first_hit, ap, num_rel, status)"""
        return (h[8 * nq: 12 * nq].view(np.int32), h[: 8 * nq].view(np.float64), h[12 * nq: 16 * nq].view(np.int32),
                h[16 * nq: 16 * nq + 16].view(np.int32))

    def reduce_host(parts):
        fh = np.concatenate([p[0] for p in parts]); apv = np.concatenate([p[1] for p in parts]); nr = np.concatenate([p[2] for p in parts])
        assert all(int(p[3][0]) == 0 for p in parts), "positives workspace overflow"
        return E.reduce_cmc_map(fh, apv, nr, 50, G)

    def step_kernels(slot):
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        evs[0].record()
        prep = E.prep_rows(feats_dev, normalize=True, precision=prec, keep_xn=False)
        q, g = prep.rows(0, Q), prep.rows(Q, Q + G)
        evs[1].record()
        d = E.dist_matrix(q, g, metric, prec, out=dist_buf)
        evs[2].record()
        res = E.rank_eval_async(d, lab["q_pid"], lab["g_pid"], lab["q_cam"], lab["g_cam"], junk)
        evs[3].record()
        stage_events.append(evs)
        launches[0] += 1 + 1 + E.RANK_EVAL_LAUNCHES
        if slot is not None:
            dev_slots[slot, : res.buf.numel()].copy_(res.buf, non_blocking=True)
            if host_slots is not None:
                host_slots[slot].copy_(dev_slots[slot], non_blocking=True)
                copy_done[slot].record()
        return res

    t_begin = [0.0]

    def run_steps(steps):
        """`steps` passes; returns the (cmc, mAP) of every pass (rank 0; None elsewhere)."""
        results = []
        t_begin[0] = time.perf_counter()
        for s in range(steps):
            step_kernels(s)
            if not distributed and s > 0:           # reduce the previous step while this one runs
                copy_done[s - 1].synchronize()
                results.append(reduce_host([unpack(host_slots[s - 1].numpy(), Q)]))
        if not distributed:
            copy_done[steps - 1].synchronize()
            results.append(reduce_host([unpack(host_slots[steps - 1].numpy(), Q)]))
            return results
        t_loop = time.perf_counter()
        if os.environ.get("MPREID_BENCH_DEBUG"):
            torch.cuda.synchronize(); t_sync = time.perf_counter()
        # all slots travel, whatever `steps` is: the warm-up then runs the collectives at the timed run's sizes
        out = torch.empty((world, n_slots, slot_bytes), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(out.view(-1), dev_slots.view(-1))
        # every rank now holds every step's per-query results: the host reductions (numpy, global query order) are dealt out
        # over the ranks (step s on rank s % N) instead of queueing on rank 0's one host thread; the (cmc, mAP) of all
        # steps then meet on every rank with one tiny all-reduce
        mine = [s for s in range(steps) if s % world == rank]
        res = torch.zeros((n_slots, 51), dtype=torch.float64, device=dev)
        if mine:
            h = out[:, mine].cpu().numpy()
            counts = [MD.shard_bounds(Qall, world, r)[1] - MD.shard_bounds(Qall, world, r)[0] for r in range(world)] if strong else [Q] * world
            vals = np.zeros((len(mine), 51))
            for k, s in enumerate(mine):
                cmc, mAP = reduce_host([unpack(h[r, k], counts[r]) for r in range(world)])
                vals[k, : len(cmc)] = cmc; vals[k, 50] = mAP
            res[mine] = torch.from_numpy(vals).to(dev)
        t_red = time.perf_counter()
        dist.all_reduce(res)
        rh = res.cpu().numpy()
        if os.environ.get("MPREID_BENCH_DEBUG"):
            print(f"[bench rank {rank}] steps queued {1e3 * (t_loop - t_begin[0]):.2f} ms, own GPU done +{1e3 * (t_sync - t_loop):.2f}, gather + D2H + "
                  f"{len(mine)} reductions +{1e3 * (t_red - t_sync):.2f}, all-reduce + D2H +{1e3 * (time.perf_counter() - t_red):.2f}", file=sys.stderr, flush=True)
        return [(rh[s, :50].astype(np.float32), np.float64(rh[s, 50])) for s in range(steps)]

    # ---- e2e inputs
    if distributed:
        # cooperative evaluation: this rank feeds its queries and ITS slice of the gallery (what a sharded feature
        # extraction leaves on each rank); the slices travel once over NVLink (distributed.sharded_evaluator)
        g_lo, g_hi = MD.aligned_shard_bounds(G, world, rank)
        e2e_batches = []
        for s0 in range(0, Q, batch):
            e2e_batches.append((qf[s0:s0 + batch].clone().pin_memory(), q_pid[s0:s0 + batch], q_cam[s0:s0 + batch]))
        for s0 in range(g_lo, g_hi, batch):
            s1 = min(g_hi, s0 + batch)
            e2e_batches.append((gf[s0:s1].clone().pin_memory(), g_pid[s0:s1], g_cam[s0:s1]))
        n_q_total = Qall if strong else world * Q
        h2d_step_bytes = int((n_q_total + G) * D * 4 + (n_q_total + G) * 16)
        e2e_api = ("distributed.sharded_evaluator(...).reset/update/compute: every rank uploads its queries and its 1/N slice of "
                   "the gallery from pinned host batches, slices are broadcast over NVLink, per-query results all-gathered")
    else:
        allf = torch.cat([qf, gf])
        pids_all, cams_all = np.concatenate([q_pid, g_pid]), np.concatenate([q_cam, g_cam])
        e2e_batches = [(allf[s:s + batch].clone().pin_memory(), pids_all[s:s + batch], cams_all[s:s + batch]) for s in range(0, allf.shape[0], batch)]
        del allf
        n_q_total = Q
        h2d_step_bytes = int((Q + G) * D * 4 + (Q + G) * 16)
        e2e_api = "R1_mAP_eval.reset/update/compute from pinned host batches"

    def step_e2e():
        if distributed:
            ev = MD.sharded_evaluator(Q, max_rank=50, feat_norm=True, precision=prec, junk=junk, metric=metric)
        else:
            ev = metrics.R1_mAP_eval(Q, max_rank=50, feat_norm=True, precision=prec, junk=junk, metric=metric)
        ev.reset()
        for f, p, c in e2e_batches:
            ev.update((f, p, c))
        cmc, mAP, *_ = quiet(ev.compute)
        return cmc, mAP

    def timed_region(fn, warm_fn, sample_clocks=False):
        warm_fn()
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()
        sampler = ClockSampler(local) if (sample_clocks and rank == 0) else None
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res = fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop() if sampler else None
        if distributed:
            dist.barrier()
        return max_over_ranks(ms, dist, dev), res, clocks

    # ---- `value`: W warm-up steps, then exactly K timed steps
    def warm():
        run_steps(args.warmup)      # the same code path as the timed steps (first use of a collective shape costs milliseconds)
        torch.cuda.synchronize()
    ms_total, results, clocks = timed_region(lambda: run_steps(args.steps), warm, sample_clocks=True)
    n_launch = launches[0] - (2 + E.RANK_EVAL_LAUNCHES) * args.warmup
    live = stage_events[-args.steps:]
    prep_ms, gemm_ms, rank_ms = [sum(e[i].elapsed_time(e[i + 1]) for e in live) / len(live) for i in range(3)]
    ms_step = ms_total / args.steps
    value = total_pairs / (ms_step * 1e-3)

    # ---- `e2e`
    e2e_steps = max(1, min(args.steps, 5))
    hb = e2e_batches[0][0]
    db = torch.empty_like(hb, device=dev)
    db.copy_(hb, non_blocking=True); torch.cuda.synchronize()
    h2d_gbs = hb.numel() * 4 / (kernel_ms(lambda: db.copy_(hb, non_blocking=True), 8) * 1e-3) / 1e9   # raw host->device rate of this box
    del db
    e2e_last = [None]

    def e2e_loop():
        for _ in range(e2e_steps):
            e2e_last[0] = step_e2e()
        return e2e_last[0]
    ms_e2e_total, res_e2e, _ = timed_region(e2e_loop, step_e2e)
    ms_e2e = ms_e2e_total / e2e_steps
    e2e_value = total_pairs / (ms_e2e * 1e-3)

    # ---- dominant kernel (distance GEMM) roofline from the in-step events; HBM roofline of the rank stage
    prep = E.prep_rows(feats_dev, normalize=True, precision=prec, keep_xn=False)
    q, g = prep.rows(0, Q), prep.rows(Q, Q + G)
    gemm_iso_ms = kernel_ms(lambda: E.dist_matrix(q, g, metric, prec, out=dist_buf), max(3, args.steps))   # back-to-back launches
    if os.environ.get("MPREID_BENCH_PROFILE"):   # per-kernel device times of the rank stage in this process (stderr)
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(5):
                E.rank_eval_async(dist_buf, lab["q_pid"], lab["g_pid"], lab["q_cam"], lab["g_cam"], junk)
            torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=10, max_name_column_width=50), file=sys.stderr)
    pk = peaks()
    flops = 2.0 * Q * G * D
    achieved_tf = flops / (gemm_ms * 1e-3) / 1e12
    peak_tf, peak_note = gemm_peak(prec, pk, dev)
    traffic, traffic_src = traffic_of("k_dist_tc", wkey, prec) if not strong else (None, None)
    roofline = {"bound": "tensor", "kernel": "k_dist_tc", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": achieved_tf / peak_tf, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_note,
                "algorithmic": "2*Q*G*D flops per launch", "ms_per_launch": gemm_ms,
                "ms_per_launch_back_to_back": gemm_iso_ms,
                "timing": "CUDA events around the launch inside the timed steps (average over the timed region)"}
    rank_gbs = (4.0 * Q * G) / (rank_ms * 1e-3) / 1e9
    plane_bytes = {"3xtf32": 8.0, "bf16": 2.0}.get(prec, 4.0)
    prep_gbs = (Q + G) * (4.0 * D + plane_bytes * prep.Dp) / (prep_ms * 1e-3) / 1e9
    stages = {"prep_ms": prep_ms, "dist_ms": gemm_ms, "rank_eval_ms": rank_ms,
              "rank_eval_roofline": {"bound": "hbm", "achieved": rank_gbs, "peak": pk["hbm"], "unit": "GB/s",
                                     "frac": rank_gbs / pk["hbm"], "algorithmic": "4*Q*G bytes (distance matrix read once)",
                                     "peak_source": pk["source"]},
              "prep_roofline": {"bound": "hbm", "achieved": prep_gbs, "peak": pk["hbm"], "unit": "GB/s", "frac": prep_gbs / pk["hbm"],
                                "algorithmic": "rows * (4*D bytes read + operand planes written)"}}
    del q, g, prep

    # ---- re-rank ms (second half of the headline metric): prep + (Q+G)^2 distances + k-reciprocal re-ranking +
    #      rank/CMC/mAP on the re-ranked matrix.  Fixed problem (strong scaling): with N ranks the rows of the
    #      all-pairs matrix are sharded and the neighbour lists / V rows are all-gathered (distributed.rerank_sharded).
    rerank = None
    if args.rerank != "none":
        rq, rg = (Qall, G) if args.rerank == "full" else (min(Qall, 3368), min(G, 15913))
        sub = torch.cat([qf_all[:rq].to(dev) if strong else feats_dev[:rq], feats_dev[Q:Q + rg]])
        rr_rows = [rq]
        rr_q_pid = torch.from_numpy(q_pid_all[:rq]).to(dev)
        rr_q_cam = torch.from_numpy(q_cam_all[:rq]).to(dev)

        def rr():
            p = E.prep_rows(sub, normalize=True, precision=prec, keep_xn=True)   # the fused all-pairs pass reads feature rows
            E.mark("prep")
            if distributed:
                dfin, q_ids = MD.rerank_sharded(p, rq, args.k1, args.k2, 0.3, prec)    # this rank's share of the query rows
                rr_rows[0] = int(q_ids.numel())
                fh, ap, nr = E.rank_eval(dfin, rr_q_pid[q_ids], lab["g_pid"][:rg], rr_q_cam[q_ids], lab["g_cam"][:rg], junk)
                E.mark("rank_eval")
                return MD.sharded_reduce(fh, ap, nr, None, 50, rg, ids=q_ids, total=rq)   # gathered into global query order
            dfin = _rerank_device(p, rq, args.k1, args.k2, 0.3, prec)
            fh, ap, nr = E.rank_eval(dfin, rr_q_pid, lab["g_pid"][:rg], rr_q_cam, lab["g_cam"][:rg], junk)
            E.mark("rank_eval")
            return E.reduce_cmc_map(fh.cpu().numpy(), ap.cpu().numpy(), nr.cpu().numpy(), 50, rg)

        rr(); torch.cuda.synchronize()
        if distributed:
            dist.barrier()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        r0.record()
        for _ in range(reps):
            _, rr_map = rr()
        r1.record(); torch.cuda.synchronize()
        rr_ms = max_over_ranks(r0.elapsed_time(r1) / reps, dist, dev)
        # stage timeline (separate pass: CUDA events between the stages; the two halves of `finish` run as separate calls)
        E.timeline_start()
        rr()
        tl = dict()
        for name, ms in E.timeline_stop():
            tl[name] = tl.get(name, 0.0) + ms
        Nn = rq + rg
        hbm = pk["hbm"]
        pk3 = pk["bf16_sustained"] / 3.0 if prec in ("3xfp16", "fp32") else peak_tf
        rows_here = rr_rows[0]
        st_roof = {}
        if "rerank.all_pairs_gemm" in tl:
            t = tl["rerank.all_pairs_gemm"] * 1e-3
            ach_alg = 2.0 * Nn * Nn * D / world / t / 1e12
            st_roof["all_pairs_gemm"] = {"ms": tl["rerank.all_pairs_gemm"], "bound": "tensor", "unit": "TFLOP/s", "achieved": ach_alg, "peak": pk3,
                                         "frac": ach_alg / pk3, "issued_frac": 0.5 * ach_alg / pk3,
                                         "algorithmic": "2*N^2*D flops (SURVEY 8d) / ranks; only upper-triangle tiles are contracted, so the "
                                                        "rate the tensor pipe actually issues is half of `achieved` (issued_frac)",
                                         "peak_source": "sustained bf16 figure / 3 MMAs per product (the launch runs for tens of ms at the power cap)"}
        for key, note in (("thresholds", "N x 2048 sampled-column GEMM + (K+2)-th smallest per row"),
                          ("topk", "candidate lists (about N * 1,000 entries of 8 B) read once"),
                          ("v0", "k-reciprocal sets, 2/3 rule, Gaussian kernel rows (sparse gathers; latency bound)"),
                          ("expand_index", "query expansion + inverted index (sort-merge of k2 sparse rows; latency bound)"),
                          ("exchange", "all-gathers of neighbour lists / V rows (NVLink)")):
            if "rerank." + key in tl:
                st_roof[key] = {"ms": tl["rerank." + key], "note": note}
        if "rerank.jaccard_blend" in tl:
            gbs = 8.0 * rows_here * rg / (tl["rerank.jaccard_blend"] * 1e-3) / 1e9
            st_roof["jaccard_blend"] = {"ms": tl["rerank.jaccard_blend"], "bound": "hbm", "unit": "GB/s", "achieved": gbs, "peak": hbm, "frac": gbs / hbm,
                                        "algorithmic": "8*Q*G bytes (distance block read + final written)"}
        if "rank_eval" in tl:
            gbs = 4.0 * rows_here * rg / (tl["rank_eval"] * 1e-3) / 1e9
            st_roof["rank_eval"] = {"ms": tl["rank_eval"], "bound": "hbm", "unit": "GB/s", "achieved": gbs, "peak": hbm, "frac": gbs / hbm}
        rerank = {"ms": rr_ms, "unit": "ms", "n_gpus": world, "scaling": "strong", "Q": rq, "G": rg, "k1": args.k1, "k2": args.k2,
                  "lambda": 0.3, "mAP": float(rr_map), "stages_ms": tl, "stage_rooflines": st_roof,
                  "pipeline": "fused (no N x N matrix)" if os.environ.get("MPREID_RERANK_FUSED", "auto").lower() not in ("0", "off", "no") else "materialising",
                  "includes": "prep + (Q+G)^2 distances (upper-triangle tiles; top-k candidates from the GEMM epilogue) + k-reciprocal "
                              "re-ranking + rank/CMC/mAP, features resident in HBM, result on the host"}

    if rank == 0:
        # ---- CPU baselines on bounded samples (rank 0, N=1 only)
        cpu = None
        if world == 1 and args.cpu_queries > 0:
            torch.set_num_threads(os.cpu_count())
            nq = min(Q, args.cpu_queries)
            t0 = time.perf_counter()
            _, cpu_map = quiet(cpu_reference_pass, *cpu_sample(data, nq), metric=metric, junk=junk)
            dt = time.perf_counter() - t0
            cpu = {"value": nq * G / dt, "unit": "pairs/s", "cores": os.cpu_count(), "kind": "port",
                   "sample": f"first {nq} of {Q} queries x full gallery, one pass, {dt:.1f} s "
                             "(torch-CPU sgemm + numpy argsort + Python loop, oracle port of utils/metrics.py)"}
            if rerank is not None and args.cpu_rerank_n > 0:
                cpu["rerank"] = cpu_rerank_sample(data, args.cpu_rerank_n, args.k1, args.k2)
                rerank["cpu_same_run"] = {k: cpu["rerank"][k] for k in ("ms", "N", "Q", "G", "cores", "kind")}
        cmc, mAP = results[-1]
        assert all(r[1] == mAP for r in results), "the timed steps must all produce the same mAP"
        line = {
            "metric": METRIC_NAME, "value": value, "unit": "pairs/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": dtype_name(prec), "data": "synthetic",
            "config": {"workload": workload, "Q_per_gpu": Q, "Q_total": Qall if strong else world * Q, "G": G, "D": D, "distance": metric,
                       "precision": prec, "feat_norm": True, "junk": junk,
                       "l2": f"inputs larger than L2 (features {(Q + G) * D * 4 / 1e9:.2f} GB, distance matrix {Q * G * 4 / 1e9:.2f} GB per pass)",
                       "sharding": ("strong: the named problem, query rows split over the ranks, gallery prepared on every rank" if strong else
                                    "query rows per GPU; value: gallery resident on every GPU") +
                                   "; e2e: gallery slices uploaded per rank and broadcast over NVLink; per-query results of all timed steps "
                                   "travel with one all-gather after the last step (N>1) / one pinned copy per step (N=1)"},
            "e2e": {"value": e2e_value, "unit": "pairs/s", "ms_per_step": ms_e2e, "steps": e2e_steps,
                    "h2d_bytes_per_step": h2d_step_bytes, "d2h_bytes_per_step": int(n_q_total * 24),
                    "api": e2e_api, "mAP": float(res_e2e[1]),
                    "h2d_gbs_measured": h2d_gbs, "h2d_floor_ms": (h2d_step_bytes / world) / (h2d_gbs * 1e9) * 1e3},
            "gpu_launches": int(n_launch), "clocks": clocks, "roofline": roofline, "stages": stages, "cpu_baseline": cpu,
            "rerank": rerank, "mAP": float(mAP), "rank1": float(cmc[0]),
        }
        if distributed:
            try:
                # SURVEY 8d "multi-GPU comm cost": bytes every rank RECEIVES per step (from the shapes; NVLink 5 = 900 GB/s per direction)
                nvl = 900e9
                g_own = MD.aligned_shard_bounds(G, world, 0)[1]
                ex = float(G - g_own) * D * 4
                lab_b = float(world) * 2 * g_own * 8
                res_b = float(world) * 3 * max(Q, 1) * 8
                comm = {"unit": "bytes received per rank per step", "nvlink_gbs_per_direction": 900,
                        "value_arm": {"per_step": 0, "after_last_step": float(world) * n_slots * slot_bytes + n_slots * 51 * 8,
                                      "note": "no collective inside the timed loop; one all-gather of the packed per-query results of all steps + one all-reduce of the (cmc, mAP) rows"},
                        "e2e_arm": {"gallery_exchange": ex, "labels_allgather": lab_b, "per_query_results_allgather": res_b,
                                    "nvlink_floor_ms": (ex + lab_b + res_b) / nvl * 1e3,
                                    "exchange": os.environ.get("MPREID_SHARD_EXCHANGE", "p2p") + " (p2p: copy-engine pulls from symmetric memory; nccl: broadcasts)"}}
                if rerank is not None:
                    Nn, Kn = rerank["Q"] + rerank["G"], int(E.rerank_neighbor_count(args.k1, args.k2))
                    keys_b = float(Nn) * Kn * 8                      # all-to-all: world pieces of (N / world) rows x K keys
                    lists_b = float(Nn) * Kn * 8                     # merged neighbour lists: int32 index + fp32 value
                    comm["rerank"] = {"thresholds_allgather": 4.0 * Nn, "row_max_allreduce": 4.0 * (Nn + 1), "partial_topk_keys_all_to_all": keys_b,
                                      "neighbour_lists_allgather": lists_b,
                                      "v0_rows_allgather": "N * W0 * 6 + 4 N, W0 = longest V0 row (<= (k1+1)(round(k1/2)+2) entries), decided at run time",
                                      "expanded_rows_allgather": "N * W1 * 6 + 4 N, W1 = longest expanded row", "K": Kn,
                                      "nvlink_floor_ms_fixed_part": (8.0 * Nn + 4 + keys_b + lists_b) / nvl * 1e3}
                line["comm"] = comm
            except Exception as e:      # reporting only: never lose the measured line over it
                line["comm"] = {"error": f"{type(e).__name__}: {e}"}
        if args.bf16_delta and prec != "bf16":
            # stated low-precision mode (BASELINE config 2): one extra pass, outside every timed region
            pb = E.prep_rows(feats_dev, normalize=True, precision="bf16", keep_xn=False)
            dbf = E.dist_matrix(pb.rows(0, Q), pb.rows(Q, Q + G), metric, "bf16", out=dist_buf)
            fh, apv, nr = E.rank_eval_host(dbf, lab["q_pid"], lab["g_pid"], lab["q_cam"], lab["g_cam"], junk)
            cb, mb = E.reduce_cmc_map(fh, apv, nr, 50, G)
            bf_ms = kernel_ms(lambda: E.dist_matrix(pb.rows(0, Q), pb.rows(Q, Q + G), metric, "bf16", out=dist_buf), 5)
            line["bf16_mode"] = {"mAP": float(mb), "rank1": float(cb[0]), "delta_mAP": float(mb - mAP), "delta_rank1": float(cb[0] - cmc[0]),
                                 "dist_ms": bf_ms, "tflops": flops / (bf_ms * 1e-3) / 1e12, "frac_of_bf16_peak": flops / (bf_ms * 1e-3) / 1e12 / pk["bf16"]}
        print(json.dumps(line))
    if distributed:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------ our arm: config 5 (retrieval)
def run_c5(args):
    """BASELINE config 5: 100k queries x 1M gallery x 768-d, top-100 + CMC/mAP; queries sharded over the ranks.
    value: gallery resident on every GPU.  e2e: the gallery starts in rank 0's pinned host memory and is uploaded and
    broadcast over NVLink (NCCL) in chunks inside the timed region; every rank uploads its own queries."""
    from mp_reid_b200 import engine as E, retrieval
    from mp_reid_b200 import distributed as MD
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    distributed = world > 1
    dist = init_dist(dev) if distributed else None
    s = synth.SHAPES["retrieval"]
    Q, G, D = max(world, int(s.Q * args.scale)), max(256, int(s.G * args.scale)), s.D
    n_id = max(2, int(s.n_id * args.scale))
    prec, k = args.precision, 100
    # per-rank seeded draws (a 3 GB host randn per rank would take longer than the benchmark): identity centres are the
    # same on every rank, rank 0 draws the gallery, every rank draws its own query shard
    gen = torch.Generator("cpu").manual_seed(s.seed)
    centers = torch.randn(n_id, D, generator=gen)
    lo, hi = MD.shard_bounds(Q, world, rank)
    gq = torch.Generator("cpu").manual_seed(1000 + s.seed + rank)
    q_pid = torch.randint(0, n_id, (hi - lo,), generator=gq)
    qf_host = (centers[q_pid] + s.sigma * torch.randn(hi - lo, D, generator=gq)).pin_memory()
    g_pid = torch.empty((G,), dtype=torch.int64)
    gf_host = None
    if rank == 0:
        g_pid = torch.randint(0, n_id, (G,), generator=gen)
        gf_host = torch.empty((G, D), dtype=torch.float32).pin_memory()
        for c0 in range(0, G, 65536):
            c1 = min(G, c0 + 65536)
            gf_host[c0:c1] = centers[g_pid[c0:c1]] + s.sigma * torch.randn(c1 - c0, D, generator=gen)
    del centers
    g_pid_d = g_pid.to(dev)
    q_pid_d = q_pid.to(dev)
    gf = torch.empty((G, D), dtype=torch.float32, device=dev)
    chunk = 65536
    copy_stream = torch.cuda.Stream(device=dev)

    def load_gallery():
        """rank 0: pinned host -> device in chunks on a copy stream; every chunk is broadcast as soon as it has landed"""
        works = []
        for c0 in range(0, G, chunk):
            c1 = min(G, c0 + chunk)
            if rank == 0:
                with torch.cuda.stream(copy_stream):
                    gf[c0:c1].copy_(gf_host[c0:c1], non_blocking=True)
                    ev = torch.cuda.Event(); ev.record(copy_stream)
                torch.cuda.current_stream().wait_event(ev)
            if distributed:
                works.append(dist.broadcast(gf[c0:c1], src=0, async_op=True))
        for w in works:
            w.wait()

    copy_stream.wait_stream(torch.cuda.current_stream())
    load_gallery()
    if distributed:
        dist.broadcast(g_pid_d, src=0)
    qf = qf_host.to(dev)
    torch.cuda.synchronize()
    counts = [MD.shard_bounds(Q, world, r)[1] - MD.shard_bounds(Q, world, r)[0] for r in range(world)]

    def finish(r):
        if distributed:
            fh, apv, nr = MD.gather_per_query(r["first_hit"], r["ap"], r["num_rel"], counts)
        else:
            fh, apv, nr = r["first_hit"].cpu().numpy(), r["ap"].cpu().numpy(), r["num_rel"].cpu().numpy()
        return E.reduce_cmc_map(fh, apv, nr, 50, G) if rank == 0 else None

    def step():
        return finish(retrieval.retrieve(qf, gf, q_pid_d, g_pid_d, k=k, precision=prec, return_device=True))

    def step_e2e():
        copy_stream.wait_stream(torch.cuda.current_stream())   # the previous step has finished reading the gallery buffer
        load_gallery()
        q_dev = qf_host.to(dev, non_blocking=True)
        out = retrieval.retrieve(q_dev, gf, q_pid_d, g_pid_d, k=k, precision=prec, return_device=True)
        top_host = out["topk"].cpu()        # the retrieval result itself: top-100 indices of this rank's queries
        return finish(out), top_host

    def timed(fn, steps, warmup, clocks=False):
        res = None
        for _ in range(warmup):
            res = fn()
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()
        sampler = ClockSampler(local) if (clocks and rank == 0) else None
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            res = fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        ck = sampler.stop() if sampler else None
        return max_over_ranks(ms, dist, dev), res, ck

    steps = max(1, min(args.steps, 3))
    ms_step, res, clocks = timed(step, steps, args.warmup, clocks=True)
    ms_e2e, res_e2e, _ = timed(step_e2e, max(1, min(steps, 2)), 1)
    # dominant kernel: the distance GEMM of one query chunk
    rows = min(hi - lo, retrieval._chunk_rows(G, int(os.environ.get("MPREID_BLOCK_BYTES", str(16 << 30)))))
    gp = E.prep_rows(gf, normalize=True, precision=prec, keep_xn=False)
    qp = E.prep_rows(qf[:rows], normalize=True, precision=prec, keep_xn=False)
    blk = E.alloc_dist(rows, G, dev)
    gemm_ms = kernel_ms(lambda: E.dist_matrix(qp, gp, "sqeuclid", prec, out=blk), 3)
    topk_ms = kernel_ms(lambda: E.row_topk(blk, k), 3)
    rank_ms = kernel_ms(lambda: E.rank_eval_async(blk, q_pid_d[:rows], g_pid_d), 3)
    pk = peaks()
    peak_tf, peak_note = gemm_peak(prec, pk, dev)
    ach = 2.0 * rows * G * D / (gemm_ms * 1e-3) / 1e12
    if rank == 0:
        cmc, mAP = res
        n_chunks = (hi - lo + rows - 1) // rows
        line = {"metric": METRIC_NAME, "value": Q * G / (ms_step * 1e-3), "unit": "pairs/s", "n_gpus": world, "steps": steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": dtype_name(prec), "data": "synthetic (per-rank seeded draws)",
                "config": {"workload": f"retrieval {Q} x {G} x {D}, euclidean, top-{k} + rank + CMC/mAP (BASELINE config 5)", "Q_total": Q, "G": G, "D": D,
                           "precision": prec, "chunk_rows": rows, "l2": "inputs larger than L2 (gallery planes 3 GB, up to 16 GB distance block per chunk)",
                           "sharding": "queries split over the ranks, gallery replicated; the Q x G matrix (400 GB) never exists: query chunks of <= 16 GB"},
                "e2e": {"value": Q * G / (ms_e2e * 1e-3), "unit": "pairs/s", "ms_per_step": ms_e2e,
                        "h2d_bytes_per_step": int(G * D * 4 + Q * D * 4), "d2h_bytes_per_step": int(Q * k * 4 + Q * 24),
                        "nvlink_broadcast_bytes": int(G * D * 4) if distributed else 0,
                        "api": "retrieval.retrieve after uploading the gallery from rank 0's pinned host memory (+ NCCL broadcast) and the "
                               "rank's queries; top-100 indices and per-query results read back", "mAP": float(res_e2e[0][1])},
                "gpu_launches": int(steps * (n_chunks * (3 + E.RANK_EVAL_LAUNCHES) + 1)),
                "clocks": clocks,
                "roofline": {"bound": "tensor", "kernel": "k_dist_tc", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
                             "traffic": None, "peak_source": peak_note, "algorithmic": "2*rows*G*D flops per launch (one query chunk)",
                             "ms_per_launch": gemm_ms, "timing": "CUDA events, back-to-back launches of one chunk"},
                "stages": {"dist_ms_per_chunk": gemm_ms, "topk_ms_per_chunk": topk_ms, "rank_eval_ms_per_chunk": rank_ms, "chunks_per_rank": n_chunks,
                           "topk_roofline": {"bound": "hbm", "achieved": 4.0 * rows * G / (topk_ms * 1e-3) / 1e9, "peak": pk["hbm"], "unit": "GB/s",
                                             "frac": 4.0 * rows * G / (topk_ms * 1e-3) / 1e9 / pk["hbm"]},
                           "rank_eval_roofline": {"bound": "hbm", "achieved": 4.0 * rows * G / (rank_ms * 1e-3) / 1e9, "peak": pk["hbm"], "unit": "GB/s",
                                                  "frac": 4.0 * rows * G / (rank_ms * 1e-3) / 1e9 / pk["hbm"]}},
                "cpu_baseline": None, "mAP": float(mAP), "rank1": float(cmc[0])}
        if world == 1 and args.cpu_queries > 0:
            from oracle import mpreid_oracle as orc
            torch.set_num_threads(os.cpu_count())
            nq = min(hi - lo, 64)
            t0 = time.perf_counter()
            fq = orc.l2_normalize(qf_host[:nq].numpy()); fg = orc.l2_normalize(gf_host.numpy())
            dmat = orc.sq_euclidean(fq, fg)
            quiet(orc.eval_func, dmat, q_pid[:nq].numpy(), g_pid.numpy(), np.zeros(nq, np.int64), np.ones(G, np.int64), sort_kind=None)
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": nq * G / dt, "unit": "pairs/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"first {nq} queries x full gallery, one pass, {dt:.1f} s (gallery normalisation included)"}
        print(json.dumps(line))
    if distributed:
        dist.barrier()
        dist.destroy_process_group()


WORKLOADS = {
    # name: (synth shape, metric, junk, description)
    "msmt17": ("msmt17", "sqeuclid", "none", "euclidean, dist+rank+CMC/mAP"),
    "market": ("market", "sqeuclid", "none", "euclidean, dist+rank+CMC/mAP"),
    "cctv": ("cctv", "arccos", "pid_cam", "arccos-cosine distance, pid+camera junk rule, dist+rank+CMC/mAP"),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="msmt17", choices=["msmt17", "market", "cctv", "c5"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--precision", default=os.environ.get("MPREID_PRECISION", "3xfp16"))
    ap.add_argument("--metric", default=None)
    ap.add_argument("--junk", default=None)
    ap.add_argument("--rerank", default="full", choices=["none", "market", "full"])
    ap.add_argument("--k1", type=int, default=20)
    ap.add_argument("--k2", type=int, default=6)
    ap.add_argument("--cpu-queries", type=int, default=1500, help="queries in the bounded CPU-baseline sample")
    ap.add_argument("--cpu-rerank-n", type=int, default=12000, help="samples (Q+G) of the bounded CPU re-ranking baseline (0 = skip)")
    ap.add_argument("--ref-queries", type=int, default=600, help="queries per step of the reference arm")
    ap.add_argument("--bf16-delta", action="store_true", help="also report mAP / rank-1 of the stated bf16 mode (default on for cctv)")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the workload (debugging only)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    world = env_int("WORLD_SIZE", 1)
    if args.impl == "reference" and env_int("RANK", 0) != 0:
        return
    if world == 1 and args.gpus > 1 and args.impl == "ours":
        # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    if args.workload == "c5":
        if args.impl == "reference":
            print(json.dumps({"impl": "reference", "unavailable": "config 5 has no reference equivalent (the reference would need a 400 GB matrix); "
                                                                    "see cpu_baseline of the c5 line for the oracle on a 64-query slice"}))
            return
        return run_c5(args)
    shape, metric, junk, desc = WORKLOADS[args.workload]
    args.metric = args.metric or metric
    args.junk = args.junk or junk
    if args.workload == "cctv":
        args.bf16_delta = True
    s = synth.SHAPES[shape]
    workload = f"{s.name} shape {int(round(s.Q * args.scale))} x {int(round(s.G * args.scale))} x {s.D}, {desc}"
    data = synth.make_shape(shape, args.scale)
    if args.impl == "reference":
        run_reference(args, data, workload)
    else:
        run_ours(args, data, workload, args.workload)


if __name__ == "__main__":
    main()
