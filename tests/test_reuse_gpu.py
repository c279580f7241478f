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
    # trajectory slots that revisit a patch of their episode are copied inside HBM, not re-read from the host
    recorded = int(want["masks"].sum())
    pos, m = want["positions"].cpu().numpy(), want["masks"].cpu().numpy()
    distinct = sum(len({tuple(pos[i, t]) for t in range(T) if m[i, t] > 0}) for i in range(b))
    assert int(stats["host_traj_tiles"]) == distinct and distinct < recorded, (distinct, recorded)
    assert int(stats["status"].item()) == 0


def test_tile_dedupe_marks_first_occurrences_and_repeats():
    from jolineedle_b200 import _cabi

    T = 5
    #            episode 0: A B A (pad) (pad)      episode 1: C C D C D
    pos = torch.tensor([[1, 1], [1, 2], [1, 1], [0, 0], [0, 0], [2, 0], [2, 0], [0, 3], [2, 0], [0, 3]], dtype=torch.int64)
    src = torch.tensor([0, 0, 0, -1, -1, 1, 1, 1, 1, 1], dtype=torch.int32)
    first = torch.empty(10, dtype=torch.int32, device="cuda")
    repeat = torch.empty(10, dtype=torch.int32, device="cuda")
    p, s_ = pos.cuda(), src.cuda()
    _cabi.check(_cabi.lib().jn_tile_dedupe(p.data_ptr(), s_.data_ptr(), 10, T, first.data_ptr(), repeat.data_ptr(),
                                           _cabi.stream_ptr(p.device)))
    assert first.tolist() == [0, 0, -2, -1, -1, 1, -2, 1, -2, -2]
    assert repeat.tolist() == [-2, -2, 0, -2, -2, -2, 5, -2, 5, 7]


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
        # -1 = zero fill, JN_SRC_SKIP (-2) = always left untouched, also without the flag
        out.fill_(7.0)
        s.gather(pos, src_index=torch.tensor([2, -2, -1, -2], dtype=torch.int32).cuda(), out=out, engine=engine)
        assert bool((out[1] == 7).all()) and bool((out[3] == 7).all()) and float(out[2].abs().sum()) == 0.0, engine
    u8 = ImageSet(torch.from_numpy(synth_u8(3, 3, 2 * P, 3 * P, salt=1)).cuda(), P)
    for engine in ("tensor", "bulk", "ldg"):  # the same through the converting kernel
        for focus in (False, True):
            out = torch.full(u8.out_shape(4, focus), 7.0, device="cuda")
            u8.gather(pos, src_index=torch.tensor([2, -2, -1, -2], dtype=torch.int32).cuda(), out=out, engine=engine,
                      normalize=True, focus=focus)
            assert bool((out[1] == 7).all()) and bool((out[3] == 7).all()) and float(out[2].abs().sum()) == 0.0
            assert float(out[0].sum()) > 0
