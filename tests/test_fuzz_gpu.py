"""Randomised parity: the CUDA path against the oracle over random geometries, dtypes and options
(the GPU-side counterpart of `tests/golden/make_golden.py --fuzz`, which pins the oracle on the reference)."""
import random

import numpy as np
import pytest
import torch

from helpers import focus_restatement, load_golden, random_boxes, synth_u8
from oracle.gaze_oracle import GazeOracle, returns_oracle
from oracle.traj_oracle import generate_trajectories_oracle

pytestmark = pytest.mark.gpu


def test_fuzz_general_env_against_oracle():
    from jolineedle_b200.env.general_env import NeedleGeneralEnv
    from jolineedle_b200.reinforce import rollout_tail

    rng = np.random.default_rng(1234)
    table = torch.from_numpy(load_golden("norm.npz")["u8_over_255"])
    for it in range(40):
        P = int(rng.choice([8, 16, 32, 64]))
        gh, gw = int(rng.integers(1, 12)), int(rng.integers(1, 12))
        b, n, T = int(rng.integers(1, 9)), int(rng.integers(1, 5)), int(rng.integers(1, 14))
        stop, u8_mode, focus, history = (bool(rng.integers(0, 2)) for _ in range(4))
        h, w = gh * P, gw * P
        u8 = torch.from_numpy(synth_u8(b, 3, h, w, salt=it))
        images = table[u8.long()]
        boxes = random_boxes(rng, b, n, h, w, 3 * P + 2)
        if n > 1:
            boxes[0, n - 1] = 0
        boxes[-1, 0] = (w - 3, h - 3, w + 20, h + 20)  # sticks out of the image
        orc = GazeOracle(images, boxes, P, T, 1, stop)
        env = NeedleGeneralEnv((u8 if u8_mode else images).cuda(), torch.from_numpy(boxes), P, T, 1, stop,
                               normalize=u8_mode, focus=focus, history=history)
        assert np.array_equal(env.bbox_masks.cpu().numpy(), orc.bbox_masks), it

        def expect(p):
            return focus_restatement(p) if focus else p

        torch.manual_seed(it); p_o, i_o = orc.reset()
        torch.manual_seed(it); p_e, i_e = env.reset()
        assert torch.equal(p_e.cpu(), expect(p_o)) and np.array_equal(i_e["positions"].cpu().numpy(), i_o["positions"])
        rew, term = [], []
        for t in range(T):
            a = rng.integers(0, 9 if stop else 8, size=b).astype(np.int64)
            o, e = orc.step(a), env.step(torch.from_numpy(a))
            assert torch.equal(e[0].cpu(), expect(o[0])), (it, t)
            assert np.array_equal(e[1].cpu().numpy(), o[1]) and np.array_equal(e[2].cpu().numpy(), o[2]), (it, t)
            assert np.array_equal(e[3].cpu().numpy(), o[3]) and np.array_equal(e[4]["positions"].cpu().numpy(), o[4]["positions"])
            rew.append(e[1]); term.append(e[2])
        assert np.array_equal(env.visited_patches.cpu().numpy(), orc.visited)
        assert np.array_equal(env.prop_patches_found.cpu().numpy(), orc.prop_patches_found())
        tail = rollout_tail(torch.stack(rew), torch.stack(term))
        masks = torch.cat([torch.ones((b, 1), dtype=torch.bool), ~torch.stack(term, 1).cpu()], dim=1)
        want, lm = returns_oracle(torch.stack(rew, 1).cpu(), masks)
        assert torch.equal(tail["returns"].cpu(), want) and torch.equal(tail["logit_masks"].cpu(), lm)
        if history:
            assert tuple(env.patch_history().shape[:2]) == (b, T + 1)
        env.check_status()


def test_fuzz_supervised_batches_against_oracle():
    from jolineedle_b200.env.simple_env import generate_trajectories
    from jolineedle_b200.utils import BBox, Position

    rng = np.random.default_rng(99)
    table = torch.from_numpy(load_golden("norm.npz")["u8_over_255"])
    for it in range(30):
        P = int(rng.choice([16, 32, 64]))
        b, T = int(rng.integers(1, 10)), int(rng.integers(1, 12))
        binomial, u8_mode = bool(rng.integers(0, 2)), bool(rng.integers(0, 2))
        kmin = int(rng.integers(0, 3)); kmax = kmin + int(rng.integers(0, 3))
        planner = ["native", "python"][it % 2]
        imgs_u8, boxes_raw = [], []
        for i in range(b):
            gh, gw = int(rng.integers(1, 9)), int(rng.integers(1, 9))
            h, w = gh * P, gw * P
            imgs_u8.append(torch.from_numpy(synth_u8(1, 3, h, w, salt=it * 16 + i)[0]))
            raw = []
            for _ in range(int(rng.integers(0, 4))):
                bw, bh = (int(v) for v in rng.integers(2, 2 * P, size=2))
                x1, y1 = int(rng.integers(-4, w - 1)), int(rng.integers(-4, h - 1))
                raw.append((x1, y1, x1 + bw, y1 + bh))
            boxes_raw.append(raw)
        imgs_f32 = [table[t.long()] for t in imgs_u8]
        seeds = [int(s) for s in rng.integers(0, 2**40, size=b)]
        position = None
        if it % 4 == 0:
            position = Position(0, 0)
        random.seed(it)
        want = generate_trajectories_oracle(imgs_f32, [[((y1, x1), (y2, x2)) for (x1, y1, x2, y2) in r] for r in boxes_raw],
                                            list(range(b)), P, T, kmin, kmax, binomial, position=position, seeds=seeds)
        random.seed(it)
        got = generate_trajectories(
            {"image": [(u if u8_mode else f).cuda() for u, f in zip(imgs_u8, imgs_f32)],
             "bboxes": [[BBox(Position(y1, x1), Position(y2, x2)) for (x1, y1, x2, y2) in r] for r in boxes_raw],
             "class_id": list(range(b))},
            P, T, kmin, kmax, binomial_keypoints=binomial, position=position, seeds=seeds, normalize=u8_mode,
            planner=planner)
        assert set(got) == set(want)
        for k in want:
            assert got[k].dtype == want[k].dtype and torch.equal(got[k].cpu(), want[k]), (it, planner, k)


def test_fuzz_supervised_float_boxes_against_oracle():
    """Boxes that are not whole pixels (jn_patch_bitmaps_f64 / jn_local_boxes_f64 + the python planner): random
    scales, boxes hugging patch edges and the 5 % threshold, mixed with whole-pixel boxes in the same batch.
    (On exactly these 30 seeded batches -- 118 float boxes -- the oracle was run against the unmodified
    reference in the build container: 0 mismatches.)"""
    from jolineedle_b200.env.simple_env import generate_trajectories
    from jolineedle_b200.utils import BBox, Position

    rng = np.random.default_rng(2024)
    table = torch.from_numpy(load_golden("norm.npz")["u8_over_255"])
    for it in range(30):
        P = int(rng.choice([16, 32, 64]))
        b, T = int(rng.integers(1, 7)), int(rng.integers(2, 12))
        binomial = bool(rng.integers(0, 2))
        imgs_u8, boxes_raw = [], []
        for i in range(b):
            gh, gw = int(rng.integers(2, 8)), int(rng.integers(2, 8))
            h, w = gh * P, gw * P
            imgs_u8.append(torch.from_numpy(synth_u8(1, 3, h, w, salt=it * 8 + i)[0]))
            raw = []
            scale = float(rng.choice([1.0, 1.37, 0.83, 2.0 / 3.0, 1.0001]))
            for _ in range(int(rng.integers(0, 4))):
                bw, bh = (float(v) for v in rng.integers(2, 2 * P, size=2))
                x1, y1 = float(rng.integers(0, w - 1)) * scale, float(rng.integers(0, h - 1)) * scale
                kind = int(rng.integers(0, 4))
                if kind == 0:  # exactly on patch edges
                    x1, y1 = float(P * rng.integers(0, gw)), float(P * rng.integers(0, gh))
                elif kind == 1:  # a sliver around 5 % of the patch area
                    bw, bh = 0.05 * P + float(rng.choice([-0.01, 0.0, 0.01])), float(P)
                raw.append((x1, y1, min(x1 + bw * scale, w - 0.5), min(y1 + bh * scale, h - 0.5)))
            boxes_raw.append([r for r in raw if r[2] > r[0] and r[3] > r[1]])
        imgs_f32 = [table[t.long()] for t in imgs_u8]
        seeds = [int(s) for s in rng.integers(0, 2**40, size=b)]
        random.seed(it)
        want = generate_trajectories_oracle(imgs_f32, [[((y1, x1), (y2, x2)) for (x1, y1, x2, y2) in r] for r in boxes_raw],
                                            list(range(b)), P, T, 0, 2, binomial, seeds=seeds)
        random.seed(it)
        got = generate_trajectories(
            {"image": [u.cuda() for u in imgs_u8],
             "bboxes": [[BBox(Position(y1, x1), Position(y2, x2)) for (x1, y1, x2, y2) in r] for r in boxes_raw],
             "class_id": list(range(b))},
            P, T, 0, 2, binomial_keypoints=binomial, seeds=seeds, normalize=True, check=True)
        assert set(got) == set(want)
        for k in want:
            assert got[k].dtype == want[k].dtype and torch.equal(got[k].cpu(), want[k]), (it, k)
