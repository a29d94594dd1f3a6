"""Synthetic ReID feature sets of the shapes BASELINE.json names.

There is no network for datasets, so every config is exercised on clustered
synthetic features: ``n_id`` identity centres, each query / gallery sample is
its centre plus isotropic noise.  The draw order is fixed (it pins the golden
numbers under tests/golden/), all draws come from one CPU generator.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch


@dataclass(frozen=True)
class Shape:
    name: str
    Q: int
    G: int
    D: int
    n_id: int
    n_cam: int
    seed: int
    sigma: float


# C1..C5 of SURVEY.md §8 (dataset shapes: datasets/market1501.py:24, datasets/msmt17.py:21)
SHAPES = {
    "market": Shape("market1501", 3368, 15913, 1280, 751, 6, 0, 3.0),
    "cctv": Shape("cctv_ir_rgb", 12000, 16000, 1280, 600, 6, 2, 3.0),
    "msmt17": Shape("msmt17", 11659, 82161, 1280, 3060, 15, 1, 3.0),
    "retrieval": Shape("retrieval_100k_1m", 100000, 1000000, 768, 50000, 8, 3, 2.0),
}


def make_set(Q, G, D, n_id, n_cam, seed, sigma, cross_modality=False):
    """Returns (qf, gf, q_pid, g_pid, q_cam, g_cam); features torch fp32 CPU, labels numpy int64."""
    gen = torch.Generator("cpu").manual_seed(seed)
    centers = torch.randn(n_id, D, generator=gen)
    q_pid = torch.randint(0, n_id, (Q,), generator=gen)
    g_pid = torch.randint(0, n_id, (G,), generator=gen)
    qf = centers[q_pid] + sigma * torch.randn(Q, D, generator=gen)
    gf = centers[g_pid] + sigma * torch.randn(G, D, generator=gen)
    q_cam = torch.randint(0, n_cam, (Q,), generator=gen)
    g_cam = torch.randint(0, n_cam, (G,), generator=gen)
    q_cam, g_cam = q_cam.numpy(), g_cam.numpy()
    if cross_modality:
        # MP-ReID camid = last digit of the 2-digit camera folder (datasets/mmmp.py:128):
        # IR cams 07..12 -> {7,8,9,0,1,2}, RGB cams 01..06 -> {1..6}
        q_cam = np.array([7, 8, 9, 0, 1, 2], dtype=np.int64)[q_cam % 6]
        g_cam = np.array([1, 2, 3, 4, 5, 6], dtype=np.int64)[g_cam % 6]
    return qf, gf, q_pid.numpy(), g_pid.numpy(), q_cam.astype(np.int64), g_cam.astype(np.int64)


def make_shape(name: str, scale: float = 1.0):
    s = SHAPES[name]
    Q = max(1, int(round(s.Q * scale)))
    G = max(1, int(round(s.G * scale)))
    n_id = max(2, int(round(s.n_id * scale)))
    return make_set(Q, G, s.D, n_id, s.n_cam, s.seed, s.sigma, cross_modality=(name == "cctv"))
