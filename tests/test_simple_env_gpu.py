"""K3a/K0/K1 parity of the supervised pipeline: golden fixtures of the reference and seeded
comparisons against the oracle, through the product's public API."""
import random

import numpy as np
import pytest
import torch

from helpers import load_golden, scenario, seed_python_random, simple_case, synth_u8, to_f32
from oracle.traj_oracle import TrajectoryOracle, collate_oracle, generate_trajectories_oracle

pytestmark = pytest.mark.gpu

SAMPLE_KEYS = ("patches", "current_actions", "next_actions", "positions", "masks", "labels", "local_bboxes",
               "patches_yolox", "bboxes_yolox")


def bboxes_of(raw):
    from jolineedle_b200.utils import BBox, Position

    return [BBox(Position(y1, x1), Position(y2, x2)) for (x1, y1, x2, y2) in raw]


def test_generate_sample_matches_reference_fixtures():
    from jolineedle_b200.env.simple_env import NeedleSimpleEnv
    from jolineedle_b200.utils import Position

    fx = load_golden("simple_env.npz")
    for name in fx["names"]:
        c, cfg = simple_case(fx, str(name))
        for variant in ("f32", "u8"):
            seed_python_random(cfg["seed"])
            img = to_f32(c["u8"]).cuda() if variant == "f32" else torch.from_numpy(c["u8"]).cuda()
            env = NeedleSimpleEnv(img, cfg["P"], bboxes_of(c["raw_boxes"].tolist()), seed=cfg["seed"],
                                  normalize=(variant == "u8"))
            pos = None if cfg["position"] is None else Position(*cfg["position"])
            s = env.generate_sample(cfg["T"], cfg["kmin"], cfg["kmax"], binomial_keypoints=cfg["binomial"], position=pos)
            assert set(s) == set(SAMPLE_KEYS)
            for k in SAMPLE_KEYS:
                assert s[k].dtype == torch.from_numpy(c[k]).dtype, (name, k)
                assert np.array_equal(s[k].cpu().numpy(), c[k]), (name, variant, k)


def test_float_boxes_match_reference_fixtures():
    """Boxes that are not whole pixels (dataset.py:258-270 scales them): labels, local boxes and detection boxes
    come from the float64 coordinates (jn_patch_bitmaps_f64 / jn_local_boxes_f64), not from truncated ones --
    per-env samples and the batched entry point, against samples of the unmodified reference."""
    from jolineedle_b200.env.simple_env import NeedleSimpleEnv, generate_trajectories

    fx = load_golden("simple_env_float.npz")
    for name in fx["names"]:
        c, cfg = simple_case(fx, str(name))
        boxes = bboxes_of(c["raw_boxes"].tolist())
        seed_python_random(cfg["seed"])
        env = NeedleSimpleEnv(to_f32(c["u8"]).cuda(), cfg["P"], boxes, seed=cfg["seed"])
        assert sorted(env.bbox_patches) == [tuple(r) for r in c["bbox_patches"].tolist()]
        s = env.generate_sample(cfg["T"], cfg["kmin"], cfg["kmax"], binomial_keypoints=cfg["binomial"])
        for k in SAMPLE_KEYS:
            assert s[k].dtype == torch.from_numpy(c[k]).dtype, (name, k)
            assert np.array_equal(s[k].cpu().numpy(), c[k]), (name, k)
        seed_python_random(cfg["seed"])
        got = generate_trajectories({"image": [torch.from_numpy(c["u8"]).cuda()], "bboxes": [boxes], "class_id": [0]},
                                    cfg["P"], cfg["T"], cfg["kmin"], cfg["kmax"], binomial_keypoints=cfg["binomial"],
                                    seeds=[cfg["seed"]], normalize=True, check=True)
        for k in SAMPLE_KEYS:
            want = c[k] if k in ("patches_yolox", "bboxes_yolox") else c[k][None]
            assert np.array_equal(got[k].cpu().numpy(), want), (name, "batched", k)


def test_collate_and_batched_entry_match_reference_fixtures():
    from jolineedle_b200.env.simple_env import NeedleSimpleEnv, generate_trajectories

    fx = load_golden("simple_env.npz")
    ref = scenario(fx, "collate")
    members = ref.pop("members").tolist()
    # (a) per-env samples + collate_fn, exactly like the reference trainer
    samples = []
    for j, idx in enumerate(members):
        c, cfg = simple_case(fx, f"s{idx:02d}")
        seed_python_random(cfg["seed"])
        env = NeedleSimpleEnv(to_f32(c["u8"]).cuda(), cfg["P"], bboxes_of(c["raw_boxes"].tolist()), seed=cfg["seed"])
        from jolineedle_b200.utils import Position

        pos = None if cfg["position"] is None else Position(*cfg["position"])
        s = env.generate_sample(cfg["T"], cfg["kmin"], cfg["kmax"], binomial_keypoints=cfg["binomial"], position=pos)
        s["class_id"] = torch.tensor(j, dtype=torch.long, device="cuda")
        samples.append(s)
    batch = NeedleSimpleEnv.collate_fn(samples)
    assert set(batch) == set(ref)
    for k, v in batch.items():
        assert np.array_equal(v.cpu().numpy(), ref[k]), k


@pytest.mark.parametrize("planner", ["native", "python"])
@pytest.mark.parametrize("binomial", [False, True])
@pytest.mark.parametrize("P,gh,gw,b,T", [(448, 5, 5, 4, 8), (64, 5, 6, 16, 8), (32, 9, 7, 8, 20)])
def test_batched_trajectories_match_oracle(planner, binomial, P, gh, gw, b, T):
    """cfg 1 (2240x2240, P=448, T=8, B=4) and smaller-patch batches, seeded, against the oracle's
    serial loop (supervised.py:116-136).  Images of one batch differ in size in the last case."""
    from jolineedle_b200.env.simple_env import generate_trajectories

    rng = np.random.default_rng(P + T + int(binomial))
    images, boxes = [], []
    for i in range(b):
        gh_i, gw_i = (gh, gw) if P != 32 else (gh - i % 3, gw + i % 2)
        h, w = gh_i * P, gw_i * P
        images.append(to_f32(synth_u8(1, 3, h, w, salt=i)[0]))
        raw = []
        for _ in range(int(rng.integers(0, 4))):
            bw, bh = (int(v) for v in rng.integers(4, P + P // 2, size=2))
            x1, y1 = int(rng.integers(0, w - 4)), int(rng.integers(0, h - 4))
            raw.append((x1, y1, min(x1 + bw, w - 1), min(y1 + bh, h - 1)))
        boxes.append(raw)
    seeds = [500 + i for i in range(b)]
    class_ids = list(range(b))
    random.seed(11)
    want = generate_trajectories_oracle(images, [[((y1, x1), (y2, x2)) for (x1, y1, x2, y2) in r] for r in boxes],
                                        class_ids, P, T, 0, 3, binomial, seeds=seeds)
    random.seed(11)
    got = generate_trajectories({"image": [im.cuda() for im in images], "bboxes": [bboxes_of(r) for r in boxes],
                                 "class_id": class_ids}, P, T, 0, 3, binomial_keypoints=binomial, seeds=seeds,
                                planner=planner)
    assert set(got) == set(want)
    for k in want:
        assert got[k].dtype == want[k].dtype and tuple(got[k].shape) == tuple(want[k].shape), k
        assert torch.equal(got[k].cpu(), want[k]), k


def test_large_batch_properties_cfg2_shape():
    """cfg 2 (B=256, 2240x2688, P=448, T=8, binomial key points 0-3) with uint8-resident images:
    every recorded slot holds the crop at its recorded position, padded slots are zero, actions
    are consistent with consecutive positions."""
    from jolineedle_b200.env.simple_env import generate_trajectories

    b, P, gh, gw, T = 256, 448, 5, 6, 8
    g = torch.Generator(device="cuda").manual_seed(8)
    base = torch.randint(0, 256, (16, 3, gh * P, gw * P), dtype=torch.uint8, device="cuda", generator=g)
    rng = np.random.default_rng(21)
    images = [base[i % 16] for i in range(b)]
    boxes = []
    for i in range(b):
        raw = []
        for _ in range(int(rng.integers(1, 4))):
            bw, bh = (int(v) for v in rng.integers(8, 448, size=2))
            x1, y1 = int(rng.integers(0, gw * P - bw)), int(rng.integers(0, gh * P - bh))
            raw.append((x1, y1, x1 + bw, y1 + bh))
        boxes.append(bboxes_of(raw))
    out = generate_trajectories({"image": images, "bboxes": boxes, "class_id": [0] * b}, P, T, 0, 3,
                                binomial_keypoints=True, seeds=list(range(b)), normalize=True)
    assert tuple(out["patches"].shape) == (b, T, 3, P, P) and out["patches"].dtype == torch.float32
    masks, pos = out["masks"].cpu(), out["positions"].cpu()
    delta = torch.tensor([(0, -1), (0, 1), (-1, 0), (1, 0), (-1, -1), (-1, 1), (1, -1), (1, 1), (0, 0)])
    table = torch.from_numpy(load_golden("norm.npz")["u8_over_255"]).cuda()  # exact CPU `x / 255` per byte value
    assert bool(((masks == 0) | (masks == 1)).all()) and bool((masks[:, 0] == 1).all())
    assert bool((masks[:, 1:] <= masks[:, :-1]).all())  # recorded slots form a prefix
    cur = out["current_actions"].cpu()
    for i in range(0, b, 7):
        n = int(masks[i].sum())
        for t in range(T):
            if t < n:
                y, x = pos[i, t].tolist()
                want = table[images[i][:, y * P:(y + 1) * P, x * P:(x + 1) * P].long()]
                assert torch.equal(out["patches"][i, t], want)
                if t > 0:
                    assert torch.equal(pos[i, t], pos[i, t - 1] + delta[cur[i, t]])
            else:
                assert float(out["patches"][i, t].abs().sum()) == 0.0
    assert int(out["next_actions"].max()) <= 7  # STOP never appears as a best action


def test_zero_copy_batch_from_pinned_host_images_matches_device_resident():
    """The end-to-end path of bench.py: pinned host images in, same tensors out as with resident images."""
    from jolineedle_b200.env.simple_env import generate_trajectories

    P, gh, gw, b, T = 64, 5, 6, 12, 8
    images = [to_f32(synth_u8(1, 3, gh * P, gw * P, salt=i)[0]) for i in range(b)]
    rng = np.random.default_rng(4)
    boxes = []
    for i in range(b):
        x1, y1 = int(rng.integers(0, gw * P - 80)), int(rng.integers(0, gh * P - 80))
        boxes.append(bboxes_of([(x1, y1, x1 + int(rng.integers(8, 80)), y1 + int(rng.integers(8, 80)))]))
    batch_dev = {"image": [im.cuda() for im in images], "bboxes": boxes, "class_id": [0] * b}
    batch_host = {"image": [im.pin_memory() for im in images], "bboxes": boxes, "class_id": [0] * b}
    random.seed(5)
    a = generate_trajectories(batch_dev, P, T, 0, 3, True, seeds=list(range(b)))
    random.seed(5)
    z = generate_trajectories(batch_host, P, T, 0, 3, True, seeds=list(range(b)), device="cuda")
    random.seed(5)
    u = generate_trajectories(batch_host, P, T, 0, 3, True, seeds=list(range(b)), device="cuda", zero_copy=False)
    for k in a:
        assert torch.equal(a[k], z[k]) and torch.equal(a[k], u[k]), k


def test_incremental_sample_api_matches_oracle():
    """The eval loop of the reference (supervised.py:280-360) drives the env step by step:
    reset -> init_sample -> add_to_sample -> step ...; deepcopy(env) must work too."""
    from copy import deepcopy

    from jolineedle_b200.env.common import Action
    from jolineedle_b200.env.simple_env import NeedleSimpleEnv
    from jolineedle_b200.utils import Position
    from oracle.traj_oracle import Pos, TrajectoryOracle

    fx = load_golden("simple_env.npz")
    for name in ("s03", "s10", "s21"):
        c, cfg = simple_case(fx, name)
        raw = c["raw_boxes"].tolist()
        img = to_f32(c["u8"])
        env = NeedleSimpleEnv(img.cuda(), cfg["P"], bboxes_of(raw), seed=cfg["seed"])
        orc = TrajectoryOracle(img, cfg["P"], [((y1, x1), (y2, x2)) for (x1, y1, x2, y2) in raw], seed=cfg["seed"])
        cpy = deepcopy(env)
        assert cpy.image is env.image and cpy.bbox_patches == env.bbox_patches and cpy.bbox_patches is not env.bbox_patches
        T = 5
        patch, infos = env.reset(Position(1, 1))
        o_patch, o_infos = orc.reset(Pos(1, 1))
        assert torch.equal(patch.cpu(), o_patch) and infos["inside_bbox"] == o_infos["inside_bbox"]
        sample = env.init_sample(T, "cuda")
        o_sample = orc._blank_sample(T)
        infos["best_action"] = Action.LEFT
        env.add_to_sample(sample, Action.LEFT, patch, infos, 0)
        orc._record(o_sample, 0, 0, o_patch, o_infos, 0)
        moves = [1, 3, 7, 0, 2, 5, 3]  # more steps than T: the buffers must double
        for i, m in enumerate(moves, start=1):
            patch, infos = env.step(Action(m))
            o_patch, o_infos = orc.step(m)
            assert infos["position"] == tuple(o_infos["position"]) and infos["number_patches_found"] == o_infos["number_patches_found"]
            assert torch.equal(infos["local_bboxes"], o_infos["local_bboxes"])
            infos["best_action"] = Action((m + 1) % 8)
            env.add_to_sample(sample, Action(m), patch, infos, i)
            orc._record(o_sample, i, m, o_patch, o_infos, (m + 1) % 8)
        for k in ("patches", "current_actions", "next_actions", "positions", "masks", "labels", "local_bboxes",
                  "patches_yolox", "bboxes_yolox"):
            assert sample[k].dtype == o_sample[k].dtype and torch.equal(sample[k].cpu(), o_sample[k]), (name, k)


@pytest.mark.parametrize("host_images", [False, True])
def test_batched_trajectories_in_focus_layout(host_images):
    """``focus=True``: the trajectory glimpses and the detection patches leave the gather in the YOLOX Focus
    space-to-depth layout; everything else is unchanged.  Also with pinned host images (tiles reused inside HBM
    are copied in that layout)."""
    from helpers import focus_restatement
    from jolineedle_b200.env.simple_env import generate_trajectories

    P, T, b = 32, 8, 6
    rng = np.random.default_rng(15)
    u8 = [torch.from_numpy(synth_u8(1, 3, 4 * P, 5 * P, salt=60 + i)[0]) for i in range(b)]
    boxes = []
    for i in range(b):
        raw = []
        for _ in range(int(rng.integers(0, 3))):
            bw, bh = (int(v) for v in rng.integers(4, 2 * P, size=2))
            x1, y1 = int(rng.integers(0, 5 * P - 4)), int(rng.integers(0, 4 * P - 4))
            raw.append((x1, y1, min(x1 + bw, 5 * P - 1), min(y1 + bh, 4 * P - 1)))
        boxes.append(bboxes_of(raw))
    images = [t.pin_memory() for t in u8] if host_images else [t.cuda() for t in u8]
    kw = dict(binomial_keypoints=True, seeds=list(range(40, 40 + b)), normalize=True, device="cuda", check=True)
    random.seed(2)
    plain = generate_trajectories({"image": images, "bboxes": boxes, "class_id": [0] * b}, P, T, 0, 3, **kw)
    random.seed(2)
    focus = generate_trajectories({"image": images, "bboxes": boxes, "class_id": [0] * b}, P, T, 0, 3, focus=True, **kw)
    assert set(plain) == set(focus)
    for k in plain:
        if k in ("patches", "patches_yolox"):
            assert tuple(focus[k].shape[-3:]) == (12, P // 2, P // 2)
            assert torch.equal(focus[k], focus_restatement(plain[k])), k
        else:
            assert torch.equal(focus[k], plain[k]), k


@pytest.mark.parametrize("where", ["cuda", "pinned"])
def test_stacked_image_batch_equals_the_list_of_images(where):
    """``batch["image"]`` as one [B, C, H, W] tensor (same-size images, e.g. a `pinned_u8_collate` batch): same
    result as the list the reference's DataLoader yields, on the device and read in place from pinned memory."""
    from jolineedle_b200.env.simple_env import generate_trajectories

    P, T, b = 32, 8, 5
    rng = np.random.default_rng(8)
    u8 = torch.from_numpy(synth_u8(b, 3, 4 * P, 6 * P, salt=70))
    boxes = []
    for i in range(b):
        raw = []
        for _ in range(int(rng.integers(0, 3))):
            bw, bh = (int(v) for v in rng.integers(4, 2 * P, size=2))
            x1, y1 = int(rng.integers(0, 6 * P - 4)), int(rng.integers(0, 4 * P - 4))
            raw.append((x1, y1, min(x1 + bw, 6 * P - 1), min(y1 + bh, 4 * P - 1)))
        boxes.append(bboxes_of(raw))
    stacked = u8.cuda() if where == "cuda" else u8.pin_memory()
    kw = dict(binomial_keypoints=True, seeds=list(range(b)), normalize=True, device="cuda", check=True)
    random.seed(4)
    a = generate_trajectories({"image": [stacked[i] for i in range(b)], "bboxes": boxes, "class_id": [1] * b}, P, T, 0, 3, **kw)
    random.seed(4)
    c = generate_trajectories({"image": stacked, "bboxes": boxes, "class_id": [1] * b}, P, T, 0, 3, **kw)
    assert set(a) == set(c)
    for k in a:
        assert torch.equal(a[k], c[k]), k
