"""Shared helpers of the test-suite (fixtures, synthetic inputs, a host mirror of K3a)."""
import os
import random
from typing import Dict, List

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name: str):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def scenario(fx, prefix: str) -> Dict[str, np.ndarray]:
    return {k[len(prefix) + 1:]: fx[k] for k in fx.files if k.startswith(prefix + "/")}


def synth_u8(b, c, h, w, salt=0) -> np.ndarray:
    """Closed-form deterministic uint8 images (same formula as tests/golden/make_golden.py)."""
    bb, cc, yy, xx = np.meshgrid(np.arange(b), np.arange(c), np.arange(h), np.arange(w), indexing="ij", sparse=True)
    v = xx * 131 + yy * 71 + cc * 29 + bb * 17 + (xx * yy) % 251 + ((xx ^ yy) * 7) % 13 + salt * 101
    return (v % 256).astype(np.uint8)


def to_f32(u8: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(u8)).float() / 255


def random_boxes(rng, b, n, h, w, max_side):
    boxes = np.zeros((b, n, 4), dtype=np.int64)
    for i in range(b):
        for j in range(n):
            bw, bh = rng.integers(2, max_side, size=2)
            x1 = rng.integers(0, max(w - bw, 1)); y1 = rng.integers(0, max(h - bh, 1))
            boxes[i, j] = (x1, y1, min(x1 + bw, w - 1), min(y1 + bh, h - 1))
    return boxes


def focus_restatement(x: torch.Tensor) -> torch.Tensor:
    """YOLOX Focus stem slicing (un-vendored `yolox` package, network_blocks.Focus): TL, BL, TR, BR
    concatenated on the channel axis.  x is [..., C, H, W]."""
    tl = x[..., ::2, ::2]
    bl = x[..., 1::2, ::2]
    tr = x[..., ::2, 1::2]
    br = x[..., 1::2, 1::2]
    return torch.cat((tl, bl, tr, br), dim=-3)


_DELTA = ((0, -1), (0, 1), (-1, 0), (1, 0), (-1, -1), (-1, 1), (1, -1), (1, 1), (0, 0))


def _heading(sy, sx, ty, tx):
    dy, dx = ty - sy, tx - sx
    s = lambda v: (v > 0) - (v < 0)  # noqa: E731
    return (4, 2, 5, 0, 8, 1, 6, 3, 7)[(s(dy) + 1) * 3 + s(dx) + 1]


def expand_plan_host(plan, T: int, inside) -> Dict[str, np.ndarray]:
    """Sequential mirror of ``jn_traj_expand`` semantics for ONE episode: walk the plan's
    segments step by step, consume the pre-drawn replacement moves in order, keep the last T
    records.  ``inside(y, x)`` is the label predicate.  Written as the obvious sequential
    program (not as the kernel's closed forms) so it checks the kernel's algebra."""
    rec = []  # (y, x, action_taken, best_action)
    y, x = plan.start
    rec.append([y, x, 0, 0])
    draws = list(plan.draws)
    di = 0
    for to, tgt, first in zip(plan.seg_to, plan.seg_tgt, plan.seg_first):
        if first:
            best = _heading(y, x, tgt[0], tgt[1])
            if best == 8:
                best = draws[di]; di += 1
            rec[-1][3] = best
        while (y, x) != (to[0], to[1]):
            act = _heading(y, x, to[0], to[1])
            y, x = y + _DELTA[act][0], x + _DELTA[act][1]
            best = _heading(y, x, tgt[0], tgt[1])
            if best == 8:
                best = draws[di]; di += 1
            rec.append([y, x, act, best])
    assert di == len(draws), (di, len(draws))
    ep_len = len(rec)
    rec = rec[max(ep_len - T, 0):]
    out = {
        "positions": np.zeros((T, 2), np.int64), "current_actions": np.zeros(T, np.int64),
        "next_actions": np.zeros(T, np.int64), "labels": np.zeros(T, np.int64), "masks": np.zeros(T, np.float32),
    }
    for i, (py, px, a, b) in enumerate(rec):
        out["positions"][i] = (py, px); out["current_actions"][i] = a; out["next_actions"][i] = b
        out["labels"][i] = int(inside(py, px)); out["masks"][i] = 1.0
    out["ep_len"] = ep_len
    return out


def simple_case(fx, name):
    """Inputs of one simple-env golden scenario."""
    c = scenario(fx, name)
    P, T, kmin, kmax, binomial, seed, py, px = (int(v) for v in c["cfg"])
    pos = None if py < 0 else (py, px)
    return c, dict(P=P, T=T, kmin=kmin, kmax=kmax, binomial=bool(binomial), seed=seed, position=pos)


def seed_python_random(seed: int):
    random.seed(seed * 7 + 1)  # convention of make_golden.py
