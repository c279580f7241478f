"""Host-resident images: detection patches that repeat a trajectory glimpse are copied inside HBM instead
of being read over PCIe again (jn_tile_lookup + JN_GATHER_SKIP_NEGATIVE).  Results must not change."""
import random

import numpy as np
import pytest
import torch

from helpers import synth_u8, to_f32

pytestmark = pytest.mark.gpu


def make_batch(b, P, gh, gw, seed):
    from jolineedle_b200.utils import BBox, Position

    rng = np.random.default_rng(seed)
    u8 = [torch.from_numpy(synth_u8(1, 3, gh * P, gw * P, salt=i)[0]) for i in range(b)]
    boxes = []
    for _ in range(b):
        raw = []
        for _ in range(int(rng.integers(0, 4))):
            x1, y1 = int(rng.integers(0, gw * P - 70)), int(rng.integers(0, gh * P - 70))
            raw.append(BBox(Position(y1, x1), Position(y1 + int(rng.integers(8, 70)), x1 + int(rng.integers(8, 70)))))
        boxes.append(raw)
    return u8, boxes


@pytest.mark.parametrize("src", ["f32", "u8"])
def test_detection_patches_reuse_trajectory_glimpses(src):
    from jolineedle_b200.env.simple_env import generate_trajectories

    b, P, gh, gw, T = 24, 64, 5, 6, 8
    u8, boxes = make_batch(b, P, gh, gw, 9)
    imgs = [t.float() / 255 for t in u8] if src == "f32" else u8
    normalize = src == "u8"
    seeds = list(range(100, 100 + b))
    dev_batch = {"image": [t.cuda() for t in imgs], "bboxes": boxes, "class_id": [0] * b}
    host_batch = {"image": [t.pin_memory() for t in imgs], "bboxes": boxes, "class_id": [0] * b}
    random.seed(1)
    want = generate_trajectories(dev_batch, P, T, 0, 3, True, seeds=seeds, normalize=normalize)
    random.seed(1)
    stats = {}
    got = generate_trajectories(host_batch, P, T, 0, 3, True, seeds=seeds, normalize=normalize, device="cuda",
                                stats=stats)
    for k in want:
        assert torch.equal(want[k], got[k]), k
    n_det = want["patches_yolox"].shape[0]
    from_host = int(stats["host_det_tiles"])
    assert 0 < from_host < n_det, (from_host, n_det)  # at least the empty patches come from the host, box patches mostly not
    assert int(stats["status"].item()) == 0


def test_skip_negative_leaves_tiles_untouched():
    from jolineedle_b200.gather import ImageSet

    P = 32
    imgs = to_f32(synth_u8(3, 3, 2 * P, 3 * P, salt=1))
    s = ImageSet(imgs.cuda(), P)
    pos = torch.tensor([[0, 1], [1, 2], [1, 0], [0, 0]], dtype=torch.int64).cuda()
    src = torch.tensor([2, -1, 0, -1], dtype=torch.int32).cuda()
    for engine in ("tensor", "bulk", "ldg"):
        out = torch.full((4, 3, P, P), 7.0, device="cuda")
        s.gather(pos, src_index=src, out=out, engine=engine, skip_negative=True)
        assert torch.equal(out[0].cpu(), imgs[2][:, 0:P, P:2 * P]) and torch.equal(out[2].cpu(), imgs[0][:, P:2 * P, 0:P])
        assert bool((out[1] == 7).all()) and bool((out[3] == 7).all()), engine
        s.gather(pos, src_index=src, out=out, engine=engine)  # default: negative = zero fill
        assert float(out[1].abs().sum()) == 0.0 and float(out[3].abs().sum()) == 0.0
