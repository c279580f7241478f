"""The oracle against the committed golden vectors (produced by the UNMODIFIED reference,
tests/golden/make_golden.py) and against the reference's own known-answer tests."""
import hashlib
import random

import numpy as np
import pytest
import torch

from helpers import load_golden, scenario, simple_case, seed_python_random, to_f32
from oracle.gaze_oracle import (GazeOracle, bbox_patch_mask_closed_form, bbox_patch_mask_raster, returns_oracle,
                                returns_oracle_numpy, split_boxes_per_patch)
from oracle.traj_oracle import TrajectoryOracle, collate_oracle

GENERAL = ["stop9", "nostop8", "randstart", "grid32", "grid40"]


def sha(t: torch.Tensor) -> str:
    return hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest()


@pytest.mark.parametrize("name", GENERAL)
def test_general_env_oracle_matches_reference(name):
    c = scenario(load_golden("general_env.npz"), name)
    env = GazeOracle(to_f32(c["u8"]), c["bboxes"], int(c["P"]), int(c["T"]), 1, bool(c["stop_enabled"]))
    assert np.array_equal(env.bbox_masks, c["bbox_masks"])
    if int(c["seed"]) >= 0:
        torch.manual_seed(int(c["seed"]))
        patches, infos = env.reset()
    else:
        patches, infos = env.reset(c["positions"][0])
    kept = {int(s): c["kept_patches"][i] for i, s in enumerate(c["kept_steps"])}
    assert np.array_equal(infos["positions"], c["positions"][0])
    assert sha(patches) == str(c["patch_sha"][0])
    assert np.array_equal(patches.numpy(), kept[0])
    for t in range(c["actions"].shape[0]):
        patches, r, te, tr, infos = env.step(c["actions"][t])
        assert np.array_equal(infos["positions"], c["positions"][t + 1])
        assert r.dtype == np.float32 and np.array_equal(r, c["rewards"][t])
        assert np.array_equal(te, c["terminated"][t]) and np.array_equal(tr, c["truncated"][t])
        assert np.array_equal(env.visited, c["visited"][t + 1])
        assert np.array_equal(env.prop_patches_found(), c["prop_patches"][t + 1])
        assert np.array_equal(env.prop_bboxes_found(), c["prop_bboxes"][t + 1])
        assert sha(patches) == str(c["patch_sha"][t + 1])
        if t + 1 in kept:
            assert np.array_equal(patches.numpy(), kept[t + 1])


def test_uint8_images_pass_through():
    c = scenario(load_golden("general_env.npz"), "u8env")
    env = GazeOracle(torch.from_numpy(c["u8"]), np.zeros((3, 1, 4), np.int64), 32, 4)
    p0, _ = env.reset(np.array([[0, 0], [1, 2], [1, 1]]))
    p1 = env.step(np.array([1, 0, 7]))[0]
    assert p0.dtype == torch.uint8
    assert np.array_equal(p0.numpy(), c["p0"]) and np.array_equal(p1.numpy(), c["p1"])


def test_reference_kat_test_env():
    """reference tests/test_env.py:10-31 -- positions after reset and RIGHT, DOWN, DOWN."""
    images = torch.zeros(1, 3, 1792, 2240)
    images[:, 0, 0:448, 448:896] = 255
    env = GazeOracle(images, np.array([[[310, 810, 400, 850], [700, 1500, 800, 1600]]]), 448, 8, 1)
    _, infos = env.reset(np.array([[1, 0]]))
    assert np.array_equal(infos["positions"], [[1, 0]])
    env.step(np.array([1]))
    env.step(np.array([3]))
    out = env.step(np.array([3]))
    assert np.array_equal(out[4]["positions"], [[3, 1]])


def test_reference_kat_test_map():
    """reference tests/test_map.py:9-34 -- the split-box golden vector."""
    env = GazeOracle(torch.zeros((1, 3, 1792, 2240)), np.array([[[410, 410, 500, 500], [1500, 1500, 1600, 1600]]]),
                     448, 20, 1)
    targets = env.detection_targets()
    expect = np.array([[0, 410, 410, 447, 447], [0, 448, 410, 500, 447], [0, 410, 448, 447, 500],
                       [0, 448, 448, 500, 500], [0, 1500, 1500, 1600, 1600]], dtype=np.int64)
    assert len(targets) == 1 and np.array_equal(targets[0], expect)
    assert np.array_equal(load_golden("detection.npz")["kat_map/targets0"], expect)


@pytest.mark.parametrize("name", ["det1", "det3"])
def test_detection_oracle_matches_reference(name):
    c = scenario(load_golden("detection.npz"), name)
    env = GazeOracle(to_f32(c["u8"]), c["bboxes"], int(c["P"]), 8, 1)
    local, present = split_boxes_per_patch(c["bboxes"], env.rows, env.cols, int(c["P"]))
    assert np.array_equal(local, c["local"]) and np.array_equal(present, c["present"])
    targets = env.detection_targets()
    assert [len(t) for t in targets] == c["targets_len"].tolist()
    assert np.array_equal(np.concatenate(targets), c["targets_cat"])
    torch.manual_seed(int(c["seed"]))
    patches, boxes = env.detection_batch(sample_neg=2)
    assert np.array_equal(patches.numpy(), c["batch_patches"]) and np.array_equal(boxes, c["batch_boxes"])


def test_mask_closed_form_equals_raster():
    rng = np.random.default_rng(5)
    for _ in range(50):
        gh, gw, P = int(rng.integers(1, 7)), int(rng.integers(1, 7)), int(rng.choice([4, 8, 16]))
        h, w = gh * P, gw * P
        boxes = rng.integers(-P, max(h, w) + P, size=(3, 4, 4)).astype(np.int64)
        boxes[..., 2:] = np.maximum(boxes[..., 2:], boxes[..., :2])  # kornia validates x2 >= x1, y2 >= y1
        assert np.array_equal(bbox_patch_mask_closed_form(boxes, h, w, P), bbox_patch_mask_raster(boxes, h, w, P))


SAMPLE_KEYS = ("patches", "current_actions", "next_actions", "positions", "masks", "labels", "local_bboxes",
               "patches_yolox", "bboxes_yolox")


def run_traj_oracle(c, cfg):
    seed_python_random(cfg["seed"])
    boxes = [((y1, x1), (y2, x2)) for (x1, y1, x2, y2) in c["raw_boxes"].tolist()]
    env = TrajectoryOracle(to_f32(c["u8"]), cfg["P"], boxes, seed=cfg["seed"])
    return env, env.generate_sample(cfg["T"], cfg["kmin"], cfg["kmax"], cfg["binomial"], cfg["position"])


def test_simple_env_oracle_matches_reference():
    fx = load_golden("simple_env.npz")
    for name in fx["names"]:
        c, cfg = simple_case(fx, str(name))
        env, s = run_traj_oracle(c, cfg)
        for k in SAMPLE_KEYS:
            assert s[k].dtype == torch.from_numpy(c[k]).dtype, (name, k)
            assert np.array_equal(s[k].numpy(), c[k]), (name, k)
        assert sorted(env.bbox_patches) == [tuple(r) for r in c["bbox_patches"].tolist()]


def test_simple_env_oracle_matches_reference_float_boxes():
    """Boxes that are not whole pixels (the dataset's minimum-size resize, dataset.py:258-270): the reference keeps
    python floats through the 5 % rule, the centre patch and the local boxes."""
    fx = load_golden("simple_env_float.npz")
    assert any((fx[f"{n}/raw_boxes"] != np.floor(fx[f"{n}/raw_boxes"])).any() for n in fx["names"])
    for name in fx["names"]:
        c, cfg = simple_case(fx, str(name))
        env, s = run_traj_oracle(c, cfg)
        for k in SAMPLE_KEYS:
            assert s[k].dtype == torch.from_numpy(c[k]).dtype, (name, k)
            assert np.array_equal(s[k].numpy(), c[k]), (name, k)
        assert sorted(env.bbox_patches) == [tuple(r) for r in c["bbox_patches"].tolist()]


def test_collate_oracle_matches_reference():
    fx = load_golden("simple_env.npz")
    samples = []
    for j, idx in enumerate(fx["collate/members"].tolist()):
        c, cfg = simple_case(fx, f"s{idx:02d}")
        _, s = run_traj_oracle(c, cfg)
        s["class_id"] = torch.tensor(j, dtype=torch.long)
        samples.append(s)
    batch = collate_oracle(samples)
    ref = scenario(fx, "collate")
    for k, v in batch.items():
        assert np.array_equal(v.numpy(), ref[k]), k


@pytest.mark.parametrize("name", ["r20", "r7", "r1"])
def test_returns_oracle_matches_reference(name):
    c = scenario(load_golden("returns.npz"), name)
    ret, lm = returns_oracle(torch.from_numpy(c["rewards"]), torch.from_numpy(c["masks"]))
    assert np.array_equal(lm.numpy(), c["logit_masks"])
    assert np.array_equal(ret.numpy(), c["returns"])  # bit-exact
    assert np.array_equal(returns_oracle_numpy(c["rewards"], c["logit_masks"]), c["returns"])


def test_normalisation_table():
    table = load_golden("norm.npz")["u8_over_255"]
    assert np.array_equal((torch.arange(256, dtype=torch.uint8).float() / 255).numpy(), table)


def test_translate_oracle_equals_torchvision_affine():
    """dataset.py:207-214 uses F.affine with an integer translation; the oracle restates it as a shift."""
    import torchvision.transforms.functional as TF

    from oracle.gaze_oracle import translate_oracle

    g = torch.Generator().manual_seed(1)
    images = torch.rand((5, 3, 40, 56), generator=g)
    shifts = [(0, 0), (7, -3), (-11, 5), (55, 39), (-56, 2)]
    got = translate_oracle(images, shifts)
    for i, (tx, ty) in enumerate(shifts):
        want = TF.affine(images[i], angle=0, translate=[tx, ty], scale=1.0, shear=0.0, fill=0.0)
        assert torch.equal(got[i], want), (tx, ty)
