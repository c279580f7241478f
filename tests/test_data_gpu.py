"""The pinned-uint8 hand-off (jolineedle_b200/data.py) end to end: a collated, page-locked uint8 batch read in
place by the env must give the crops / rewards / flags the reference's float32 hand-off gives
(dataset.py:240,307-347 -> reinforce.py:313-324)."""
import numpy as np
import pytest
import torch

from jolineedle_b200.data import pinned_u8_collate, pinned_u8_list_collate
from jolineedle_b200.utils import BBox, Position
from oracle.gaze_oracle import GazeOracle
from oracle.traj_oracle import generate_trajectories_oracle
from test_data_cpu import padded_collate_restated, samples

pytestmark = pytest.mark.gpu


def test_pinned_u8_batch_drives_the_rl_env_like_the_float_batch():
    from jolineedle_b200.env.general_env import NeedleGeneralEnv

    rng = np.random.default_rng(4)
    raw = samples(rng, 6, [(100, 130), (128, 96), (70, 160)])
    P, T = 32, 10
    batch = pinned_u8_collate([{"image": s["hwc"], "bboxes": s["bboxes"], "class_id": s["class_id"]} for s in raw], P)
    assert batch["image"].is_pinned() and batch["image"].dtype == torch.uint8
    ref = padded_collate_restated([{"image": torch.from_numpy(s["hwc"]).permute(2, 0, 1).float() / 255, **s} for s in raw], P)
    orc = GazeOracle(ref["image"], ref["bboxes"].numpy(), P, T, 1, True)
    env = NeedleGeneralEnv(batch["image"], batch["bboxes"], P, T, 1, True, device="cuda", normalize=True,
                           zero_copy=True, history=True)
    torch.manual_seed(2); p_o, _ = orc.reset()
    torch.manual_seed(2); p_e, _ = env.reset()
    assert torch.equal(p_e.cpu(), p_o)
    for t in range(T):
        a = rng.integers(0, 9, size=len(raw)).astype(np.int64)
        o, e = orc.step(a), env.step(torch.from_numpy(a))
        assert torch.equal(e[0].cpu(), o[0]) and np.array_equal(e[1].cpu().numpy(), o[1]), t
        assert np.array_equal(e[2].cpu().numpy(), o[2]) and np.array_equal(e[4]["positions"].cpu().numpy(), o[4]["positions"])
    env.check_status()


def test_pinned_u8_list_drives_the_supervised_path():
    import random

    from jolineedle_b200.env.simple_env import generate_trajectories

    rng = np.random.default_rng(6)
    P, T = 32, 8
    raw = samples(rng, 4, [(96, 128), (64, 160)])
    batch = pinned_u8_list_collate([{"image": s["hwc"], "bboxes": s["bboxes"], "class_id": s["class_id"]} for s in raw])
    assert all(im.is_pinned() for im in batch["image"])
    f32 = [torch.from_numpy(s["hwc"]).permute(2, 0, 1).float() / 255 for s in raw]
    boxes = [[((b.up_left.y, b.up_left.x), (b.bottom_right.y, b.bottom_right.x)) for b in s["bboxes"]] for s in raw]
    random.seed(1)
    want = generate_trajectories_oracle(f32, boxes, batch["class_id"], P, T, 0, 3, True, seeds=[1, 2, 3, 4])
    random.seed(1)
    got = generate_trajectories(batch, P, T, 0, 3, binomial_keypoints=True, seeds=[1, 2, 3, 4], normalize=True,
                                device="cuda", check=True)
    for k in want:
        assert torch.equal(got[k].cpu(), want[k]), k
