"""Episodes with more key-point segments than one shared-memory table of the expansion kernel holds
(kMaxSeg = 128): a box covering hundreds of patches of a large grid.  The reference keeps the last
``max_ep_len`` records of such walks (simple_env.py:573-584); so must K3."""
import random

import numpy as np
import pytest
import torch

from helpers import synth_u8, to_f32
from oracle.traj_oracle import generate_trajectories_oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("planner", ["native", "python"])
@pytest.mark.parametrize("T", [8, 50, 700])
def test_more_than_128_segments_match_oracle(planner, T):
    from jolineedle_b200.env.simple_env import generate_trajectories
    from jolineedle_b200.utils import BBox, Position

    P, gh, gw = 8, 36, 40
    h, w = gh * P, gw * P
    # image 0: one box over 20 x 30 patches (600 box patches -> hundreds of segments); image 1: two large boxes;
    # image 2: a small one (short episode in the same batch)
    raw = [[(40, 24, 40 + 30 * P - 1, 24 + 20 * P - 1)],
           [(0, 0, 15 * P, 12 * P), (20 * P + 3, 18 * P + 1, 39 * P, 35 * P)],
           [(100, 100, 120, 130)]]
    images = [to_f32(synth_u8(1, 3, h, w, salt=i)[0]) for i in range(3)]
    seeds = [11, 12, 13]
    random.seed(5)
    want = generate_trajectories_oracle(images, [[((y1, x1), (y2, x2)) for (x1, y1, x2, y2) in r] for r in raw],
                                        [0, 1, 2], P, T, 0, 3, True, seeds=seeds)
    random.seed(5)
    stats = {}
    got = generate_trajectories(
        {"image": [im.cuda() for im in images],
         "bboxes": [[BBox(Position(y1, x1), Position(y2, x2)) for (x1, y1, x2, y2) in r] for r in raw],
         "class_id": [0, 1, 2]}, P, T, 0, 3, binomial_keypoints=True, seeds=seeds, planner=planner, stats=stats,
        check=True)
    assert int(stats["ep_len"].max()) > 128  # untruncated length: well past one table of segments
    assert int(stats["status"].item()) == 0
    for k in want:
        assert torch.equal(got[k].cpu(), want[k]), (k, T, planner)
