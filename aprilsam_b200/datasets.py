"""Seeded synthetic Manhattan-world pose graphs (BASELINE.json configs 4 and 5).

A robot walks an integer grid inside a square world of side floor(0.9*sqrt(N)) cells with unit
forward steps, turning +-90 degrees with probability 0.15 each (and rotating until the step
stays in bounds).  Odometry edges i -> i+1; loop-closure candidates for pose i are earlier
poses j < i-10 in the same or a 4-adjacent cell.

  manhattan_dense(N, seed)   config 4: every candidate kept with the probability that makes the
                             total factor count ~ 4N
  manhattan_sparse(N, seed)  config 5: odometry + for each pose i >= 12, with probability 0.05,
                             one closure j -> i (a random spatial candidate, else a random
                             earlier pose j < i-10)

Measurements are the exact relative pose plus N(0, sigma^2) noise with sigma = 1/sqrt(2000) on
x, y, theta; W = diag(44.72136) (M3500's convention); the VERTEX2 initial estimate is the
dead-reckoned noisy odometry; edges are sorted by (max id, min id).  The output is a
PoseGraphData, i.e. exactly what the reference demo's text loader would produce.
"""
from __future__ import annotations

import numpy as np

from .harness import PoseGraphData

SIGMA = 1.0 / np.sqrt(2000.0)
W_DIAG = 44.721360


def _trajectory(N: int, rng: np.random.Generator):
    side = int(np.floor(0.9 * np.sqrt(N)))
    side = max(side, 2)
    dirs = np.array([[1, 0], [0, 1], [-1, 0], [0, -1]])
    xy = np.zeros((N, 2), dtype=np.int64)
    hd = np.zeros(N, dtype=np.int64)
    x, y, h = side // 2, side // 2, 0
    turn = rng.random(N)
    for i in range(N):
        if i > 0:
            if turn[i] < 0.15:
                h = (h + 1) % 4
            elif turn[i] < 0.30:
                h = (h + 3) % 4
            for _ in range(4):
                nx, ny = x + dirs[h][0], y + dirs[h][1]
                if 0 <= nx < side and 0 <= ny < side:
                    break
                h = (h + 1) % 4
            x, y = nx, ny
        xy[i] = (x, y)
        hd[i] = h
    truth = np.column_stack([xy[:, 0].astype(float), xy[:, 1].astype(float), hd * (np.pi / 2)])
    return truth, xy, side


def _rel(truth, a, b):
    pa, pb = truth[a], truth[b]
    c, s = np.cos(pa[:, 2]), np.sin(pa[:, 2])
    dx, dy = pb[:, 0] - pa[:, 0], pb[:, 1] - pa[:, 1]
    dt = pb[:, 2] - pa[:, 2]
    dt = (dt + np.pi) % (2 * np.pi) - np.pi
    return np.column_stack([c * dx + s * dy, -s * dx + c * dy, dt])


def _finish(truth, ea, eb, rng) -> PoseGraphData:
    N = len(truth)
    ea = np.asarray(ea, dtype=np.int32)
    eb = np.asarray(eb, dtype=np.int32)
    key = np.maximum(ea, eb).astype(np.int64) * (N + 1) + np.minimum(ea, eb)
    order = np.argsort(key, kind="stable")
    ea, eb = ea[order], eb[order]
    z = _rel(truth, ea, eb) + rng.normal(0.0, SIGMA, size=(len(ea), 3))
    W = np.zeros((len(ea), 9))
    W[:, 0] = W[:, 4] = W[:, 8] = W_DIAG
    # dead-reckoned initial estimate from the noisy odometry edges (a = i, b = i+1)
    init = np.zeros((N, 3))
    odo = {int(a): k for k, (a, b) in enumerate(zip(ea, eb)) if b == a + 1}
    for i in range(N - 1):
        k = odo[i]
        p = init[i]
        c, s = np.cos(p[2]), np.sin(p[2])
        init[i + 1] = (p[0] + c * z[k, 0] - s * z[k, 1], p[1] + s * z[k, 0] + c * z[k, 1], p[2] + z[k, 2])
    t0 = truth - np.array([truth[0, 0], truth[0, 1], 0.0])  # pose 0 sits at the origin (heading 0), like the prior
    return PoseGraphData(init, ea, eb, np.ascontiguousarray(z), W, t0)


def _candidates(xy, side):
    """For each pose i: earlier poses j < i-10 in the same or a 4-adjacent cell."""
    cell_of = xy[:, 0] * side + xy[:, 1]
    cells: dict[int, list[int]] = {}
    out = []
    for i in range(len(xy)):
        x, y = xy[i]
        cand = []
        for dx, dy in ((0, 0), (1, 0), (-1, 0), (0, 1), (0, -1)):
            nx, ny = x + dx, y + dy
            if 0 <= nx < side and 0 <= ny < side:
                lst = cells.get(int(nx * side + ny))
                if lst:
                    cand.extend(j for j in lst if j < i - 10)
        out.append(cand)
        cells.setdefault(int(cell_of[i]), []).append(i)
    return out


def manhattan_dense(N: int, seed: int = 1) -> PoseGraphData:
    rng = np.random.default_rng(seed)
    truth, xy, side = _trajectory(N, rng)
    cand = _candidates(xy, side)
    total = sum(len(c) for c in cand)
    p = min(1.0, 3.0 * N / max(total, 1))
    ea = list(range(N - 1))
    eb = list(range(1, N))
    for i, c in enumerate(cand):
        if not c:
            continue
        keep = rng.random(len(c)) < p
        for j, k in zip(c, keep):
            if k:
                ea.append(j)
                eb.append(i)
    return _finish(truth, ea, eb, rng)


def manhattan_sparse(N: int, seed: int = 1, p_closure: float = 0.05) -> PoseGraphData:
    rng = np.random.default_rng(seed)
    truth, xy, side = _trajectory(N, rng)
    cand = _candidates(xy, side)
    ea = list(range(N - 1))
    eb = list(range(1, N))
    u = rng.random(N)
    for i in range(12, N):
        if u[i] >= p_closure:
            continue
        c = cand[i]
        j = int(c[rng.integers(len(c))]) if c else int(rng.integers(0, i - 10))
        ea.append(j)
        eb.append(i)
    return _finish(truth, ea, eb, rng)
