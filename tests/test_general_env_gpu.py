"""K0/K2/K1 parity of the batched RL env: golden fixtures of the reference, the reference's own
KATs, and seeded comparisons against the oracle at BASELINE cfg-1 geometry."""
import numpy as np
import pytest
import torch

from helpers import load_golden, random_boxes, scenario, synth_u8, to_f32
from oracle.gaze_oracle import GazeOracle

pytestmark = pytest.mark.gpu

GENERAL = ["stop9", "nostop8", "randstart", "grid32", "grid40"]


def make_env(*a, **k):
    from jolineedle_b200.env.general_env import NeedleGeneralEnv

    return NeedleGeneralEnv(*a, **k)


@pytest.mark.parametrize("name", GENERAL)
@pytest.mark.parametrize("variant", ["f32", "u8_normalize"])
def test_env_matches_reference_fixtures(name, variant):
    c = scenario(load_golden("general_env.npz"), name)
    P, T, stop = int(c["P"]), int(c["T"]), bool(c["stop_enabled"])
    if variant == "f32":
        env = make_env(to_f32(c["u8"]).cuda(), torch.from_numpy(c["bboxes"]), P, T, 1, stop)
    else:  # uint8-resident images, normalise on gather: must equal ToTensor-then-crop bit for bit
        env = make_env(torch.from_numpy(c["u8"]).cuda(), torch.from_numpy(c["bboxes"]), P, T, 1, stop, normalize=True)
    assert np.array_equal(env.bbox_masks.cpu().numpy(), c["bbox_masks"])
    if int(c["seed"]) >= 0:
        torch.manual_seed(int(c["seed"]))
        patches, infos = env.reset()
    else:
        patches, infos = env.reset(torch.from_numpy(c["positions"][0]))
    kept = {int(s): c["kept_patches"][i] for i, s in enumerate(c["kept_steps"])}
    assert np.array_equal(infos["positions"].cpu().numpy(), c["positions"][0])
    assert patches.dtype == torch.float32 and np.array_equal(patches.cpu().numpy(), kept[0])
    for t in range(c["actions"].shape[0]):
        patches, r, te, tr, infos = env.step(torch.from_numpy(c["actions"][t]))
        assert infos["positions"].dtype == torch.int64 and r.dtype == torch.float32
        assert te.dtype == torch.bool and tr.dtype == torch.bool
        assert np.array_equal(infos["positions"].cpu().numpy(), c["positions"][t + 1]), t
        assert np.array_equal(r.cpu().numpy(), c["rewards"][t]), t  # bit-exact fp32
        assert np.array_equal(te.cpu().numpy(), c["terminated"][t]), t
        assert np.array_equal(tr.cpu().numpy(), c["truncated"][t]), t
        assert np.array_equal(env.visited_patches.cpu().numpy(), c["visited"][t + 1]), t
        assert np.array_equal(env.prop_patches_found.cpu().numpy(), c["prop_patches"][t + 1]), t
        assert np.array_equal(env.prop_bboxes_found.cpu().numpy(), c["prop_bboxes"][t + 1]), t
        assert np.array_equal(env.terminated.cpu().numpy(), c["terminated"][t]), t
        if t + 1 in kept:
            assert np.array_equal(patches.cpu().numpy(), kept[t + 1]), t
    env.check_status()


def test_uint8_images_pass_through():
    c = scenario(load_golden("general_env.npz"), "u8env")
    env = make_env(torch.from_numpy(c["u8"]).cuda(), torch.zeros((3, 1, 4), dtype=torch.long), 32, 4, 1)
    p0, _ = env.reset(torch.tensor([[0, 0], [1, 2], [1, 1]]))
    p1 = env.step(torch.tensor([1, 0, 7]))[0]
    assert p0.dtype == torch.uint8 and tuple(p0.shape) == (3, 1, 3, 32, 32)
    assert np.array_equal(p0.cpu().numpy(), c["p0"]) and np.array_equal(p1.cpu().numpy(), c["p1"])


def test_reference_kat_test_env():
    """reference tests/test_env.py:10-31, re-expressed against the new env."""
    from jolineedle_b200.env.common import Action

    images = torch.zeros(1, 3, 1792, 2240)
    images[:, 0, 0:448, 448:896] = 255
    env = make_env(images=images.cuda(), bboxes=torch.tensor([[[310, 810, 400, 850], [700, 1500, 800, 1600]]]),
                   patch_size=448, max_ep_len=8, n_glimps_levels=1)
    patches, infos = env.reset(torch.tensor([[1, 0]]))
    assert torch.equal(infos["positions"].cpu(), torch.tensor([[1, 0]]))
    assert tuple(patches.shape) == (1, 1, 3, 448, 448)
    p1 = env.step(torch.tensor([Action.RIGHT.value]))[0]
    env.step(torch.tensor([Action.DOWN.value]))
    patches, reward, terminated, truncated, infos = env.step(torch.tensor([Action.DOWN.value]))
    assert torch.equal(infos["positions"].cpu(), torch.tensor([[3, 1]]))
    # beyond the reference's assertions: the marked patch (0,1) was never entered; (1,1) is all zeros
    assert float(p1.abs().sum()) == 0.0


def test_reference_kat_test_map_and_detection_fixtures():
    """reference tests/test_map.py:9-34 + the detection fixtures."""
    env = make_env(images=torch.zeros((1, 3, 1792, 2240), device="cuda"),
                   bboxes=torch.tensor([[[410, 410, 500, 500], [1500, 1500, 1600, 1600]]]), patch_size=448,
                   max_ep_len=20, n_glimps_levels=1)
    targets = env.get_detection_targets()
    expect = torch.tensor([[0, 410, 410, 447, 447], [0, 448, 410, 500, 447], [0, 410, 448, 447, 500],
                           [0, 448, 448, 500, 500], [0, 1500, 1500, 1600, 1600]], dtype=torch.int64)
    assert len(targets) == 1 and torch.equal(targets[0].cpu(), expect)
    fx = load_golden("detection.npz")
    for name in ("det1", "det3"):
        c = scenario(fx, name)
        env = make_env(to_f32(c["u8"]).cuda(), torch.from_numpy(c["bboxes"]), int(c["P"]), 8, 1)
        local, present = env.parse_bboxes()
        assert np.array_equal(local.cpu().numpy(), c["local"]) and np.array_equal(present.cpu().numpy(), c["present"])
        targets = env.get_detection_targets()
        assert [len(t) for t in targets] == c["targets_len"].tolist()
        assert np.array_equal(torch.cat(targets).cpu().numpy(), c["targets_cat"])
        torch.manual_seed(int(c["seed"]))
        patches, boxes = env.get_detection_batch(sample_neg=2)
        assert np.array_equal(patches.cpu().numpy(), c["batch_patches"])
        assert boxes.dtype == torch.int64 and np.array_equal(boxes.cpu().numpy(), c["batch_boxes"])
        env.check_status()


@pytest.mark.parametrize("stop", [False, True])
@pytest.mark.parametrize("P,gh,gw,b,T", [(448, 5, 5, 4, 8), (448, 5, 6, 8, 20), (256, 8, 8, 6, 32)])
def test_env_matches_oracle_on_seeded_rollouts(stop, P, gh, gw, b, T):
    """BASELINE cfg 1 (2240x2240, P=448, T=8, B=4), cfg-3 geometry at small B, and a P=256 case."""
    rng = np.random.default_rng(P + T + int(stop))
    h, w = gh * P, gw * P
    u8 = synth_u8(b, 3, h, w, salt=T)
    boxes = random_boxes(rng, b, 3, h, w, 447)
    boxes[0, 2] = 0  # zero-padded row
    images = to_f32(u8)
    orc = GazeOracle(images, boxes, P, T, 1, stop)
    for variant in ("f32", "u8", "history"):
        if variant == "u8":
            env = make_env(torch.from_numpy(u8).cuda(), torch.from_numpy(boxes), P, T, 1, stop, normalize=True)
        else:
            env = make_env(images.cuda(), torch.from_numpy(boxes), P, T, 1, stop, history=(variant == "history"))
        assert np.array_equal(env.bbox_masks.cpu().numpy(), orc.bbox_masks)
        torch.manual_seed(99)
        p_o, i_o = orc.reset()
        torch.manual_seed(99)
        p_e, i_e = env.reset()
        assert np.array_equal(i_e["positions"].cpu().numpy(), i_o["positions"])
        assert torch.equal(p_e.cpu(), p_o)
        arng = np.random.default_rng(7)
        for t in range(T):
            a = arng.integers(0, 9 if stop else 8, size=b).astype(np.int64)
            o = orc.step(a)
            e = env.step(torch.from_numpy(a).cuda())
            assert torch.equal(e[0].cpu(), o[0]), (variant, t)
            assert np.array_equal(e[1].cpu().numpy(), o[1]) and np.array_equal(e[2].cpu().numpy(), o[2])
            assert np.array_equal(e[3].cpu().numpy(), o[3])
            assert np.array_equal(e[4]["positions"].cpu().numpy(), o[4]["positions"])
        if variant == "history":
            assert tuple(env.patch_history().shape) == (b, T + 1, 3, P, P)
            assert torch.equal(env.patch_history()[:, T].cpu(), o[0][:, 0])
        env.check_status()


def test_invalid_action_and_box_are_flagged():
    env = make_env(torch.zeros(2, 3, 32, 32, device="cuda"), torch.tensor([[[0, 0, 5, 5]], [[0, 0, 5, 5]]]), 16, 4, 1)
    env.reset(torch.tensor([[0, 0], [1, 1]]))
    env.step(torch.tensor([3, 11]))
    with pytest.raises(ValueError):
        env.check_status()
    with pytest.raises(AssertionError):
        make_env(torch.zeros(1, 3, 30, 32, device="cuda"), torch.zeros(1, 1, 4, dtype=torch.long), 16, 4, 1)
    with pytest.raises(AssertionError):
        make_env(torch.zeros(2, 3, 32, 32, device="cuda"), torch.zeros(1, 1, 4, dtype=torch.long), 16, 4, 1)


def test_large_batch_properties_cfg3_shape():
    """cfg-3 geometry (B=1024 episodes, 5x6 grid, T=20, STOP) with uint8-resident images: crops
    equal direct device-side slicing, rewards obey their closed form, flags are consistent."""
    b, P, gh, gw, T = 1024, 448, 5, 6, 20
    n_img = 24  # 24 distinct 2240x2688 images shared by the 1024 episodes via an explicit batch copy is too big;
    # use a smaller distinct set replicated: episodes i uses image i % n_img (replicated tensor = 18 GB at B=1024)
    g = torch.Generator(device="cuda").manual_seed(5)
    base = torch.randint(0, 256, (n_img, 3, gh * P, gw * P), dtype=torch.uint8, device="cuda", generator=g)
    images = base.repeat(b // n_img + 1, 1, 1, 1)[:b].contiguous()
    rng = np.random.default_rng(3)
    boxes = torch.from_numpy(random_boxes(rng, b, 2, gh * P, gw * P, 447))
    env = make_env(images, boxes, P, T, 1, True, normalize=True)
    torch.manual_seed(1)
    patches, infos = env.reset()
    masks = env.bbox_masks
    total = masks.sum(dim=(1, 2))
    ar = torch.arange(b, device="cuda")
    table = torch.from_numpy(load_golden("norm.npz")["u8_over_255"]).cuda()  # exact CPU `x / 255` per byte value
    visited_ref = torch.zeros_like(masks)
    visited_ref[ar, infos["positions"][:, 0], infos["positions"][:, 1]] = True
    stopped = torch.zeros(b, dtype=torch.bool, device="cuda")
    for t in range(T):
        a = torch.randint(0, 9, (b,), device="cuda", generator=g)
        pos_prev = infos["positions"]
        patches, r, te, tr, infos = env.step(a)
        pos = infos["positions"]
        delta = torch.tensor([(0, -1), (0, 1), (-1, 0), (1, 0), (-1, -1), (-1, 1), (1, -1), (1, 1), (0, 0)], device="cuda")
        want_pos = pos_prev + delta[a]
        want_pos[:, 0].clamp_(0, gh - 1); want_pos[:, 1].clamp_(0, gw - 1)
        assert torch.equal(pos, want_pos)
        stopped |= a == 8
        fresh = masks[ar, pos[:, 0], pos[:, 1]] & ~visited_ref[ar, pos[:, 0], pos[:, 1]]
        found = (visited_ref & masks).sum(dim=(1, 2))
        stop_eval = torch.where(found == total, found, found - total) * stopped
        want_r = (fresh.float() + torch.tensor(-1 / T, dtype=torch.float32, device="cuda")) + stop_eval.float()
        assert torch.equal(r, want_r)
        visited_ref[ar, pos[:, 0], pos[:, 1]] = True
        assert torch.equal(te, stopped) and bool((tr == (t + 1 >= T)).all())
        if t in (0, T - 1):
            k = torch.randint(0, b, (16,), generator=torch.Generator().manual_seed(t)).tolist()
            for i in k:
                y, x = pos[i].tolist()
                want = table[images[i, :, y * P:(y + 1) * P, x * P:(x + 1) * P].long()]
                assert torch.equal(patches[i, 0], want)
    assert torch.equal(env.visited_patches, visited_ref)
    env.check_status()


def test_cfg4_geometry_8192_patch256_seq32():
    """BASELINE cfg 4 geometry: 8192x8192 images, P=256 -> 32x32 grid (1024-bit bitmaps: one word per
    lane in the warp-per-episode step kernel), T=32, STOP enabled; against the oracle, uint8-resident."""
    b, P, g_, T = 2, 256, 32, 32
    gen = torch.Generator().manual_seed(44)
    u8 = torch.randint(0, 256, (b, 3, g_ * P, g_ * P), dtype=torch.uint8, generator=gen)
    images = u8.float() / 255
    rng = np.random.default_rng(8)
    boxes = random_boxes(rng, b, 4, g_ * P, g_ * P, 700)
    orc = GazeOracle(images, boxes, P, T, 1, True)
    env = make_env(u8.cuda(), torch.from_numpy(boxes), P, T, 1, True, normalize=True)
    assert np.array_equal(env.bbox_masks.cpu().numpy(), orc.bbox_masks)
    torch.manual_seed(3); p_o, i_o = orc.reset()
    torch.manual_seed(3); p_e, i_e = env.reset()
    assert torch.equal(p_e.cpu(), p_o) and np.array_equal(i_e["positions"].cpu().numpy(), i_o["positions"])
    for t in range(T):
        # walk towards / over the first box so that rewards and the found-all branch are exercised
        a = rng.integers(0, 9, size=b).astype(np.int64) if t % 3 else np.array([7, 4], dtype=np.int64)
        o, e = orc.step(a), env.step(torch.from_numpy(a))
        assert torch.equal(e[0].cpu(), o[0]), t
        assert np.array_equal(e[1].cpu().numpy(), o[1]) and np.array_equal(e[2].cpu().numpy(), o[2])
        assert np.array_equal(e[3].cpu().numpy(), o[3]) and np.array_equal(e[4]["positions"].cpu().numpy(), o[4]["positions"])
    assert np.array_equal(env.visited_patches.cpu().numpy(), orc.visited)
    assert np.array_equal(env.prop_patches_found.cpu().numpy(), orc.prop_patches_found())
    env.check_status()


@pytest.mark.parametrize("gh,gw", [(40, 36), (33, 33), (6, 11), (1, 70), (64, 32), (50, 50), (46, 45)])
def test_grids_of_more_than_32_bitmap_words(gh, gw):
    """Grids beyond 1024 patches (40x36 = 45 bitmap words: the step kernel keeps a second block of 32 words in
    registers), just past it (33x33 = 35 words), three words, a single row of 70 patches, exactly two blocks
    (64x32 = 64 words), and past them (50x50 = 79 words, 46x45 = 65: the lanes loop over the rest) -- every output
    against the oracle."""
    b, P, T = 37, 8, 12  # 37 episodes: two warps of the lane-per-episode kernel, the second one partial
    rng = np.random.default_rng(gh * 100 + gw)
    u8 = synth_u8(b, 3, gh * P, gw * P, salt=gh)
    images = to_f32(u8)
    boxes = random_boxes(rng, b, 3, gh * P, gw * P, 9 * P)
    for stop in (True, False):
        orc = GazeOracle(images, boxes, P, T, 1, stop)
        env = make_env(torch.from_numpy(u8).cuda(), torch.from_numpy(boxes), P, T, 1, stop, normalize=True)
        assert np.array_equal(env.bbox_masks.cpu().numpy(), orc.bbox_masks)
        torch.manual_seed(gw); p_o, i_o = orc.reset()
        torch.manual_seed(gw); p_e, i_e = env.reset()
        assert torch.equal(p_e.cpu(), p_o) and np.array_equal(i_e["positions"].cpu().numpy(), i_o["positions"])
        for t in range(T):
            a = rng.integers(0, 9 if stop else 8, size=b).astype(np.int64)
            o, e = orc.step(a), env.step(torch.from_numpy(a))
            assert torch.equal(e[0].cpu(), o[0]), t
            for k in (1, 2, 3):
                assert np.array_equal(e[k].cpu().numpy(), o[k]), (t, k)
            assert np.array_equal(e[4]["positions"].cpu().numpy(), o[4]["positions"])
        assert np.array_equal(env.visited_patches.cpu().numpy(), orc.visited)
        assert np.array_equal(env.prop_patches_found.cpu().numpy(), orc.prop_patches_found())
        assert np.array_equal(env.terminated.cpu().numpy(), orc.terminated())
        env.check_status()


def test_stepping_past_max_ep_len_and_rebound_positions():
    """The reference env keeps stepping after truncation (truncated stays True) and lets callers move it by hand
    (apply_movements rebinds env.positions): the per-episode rings roll over, results stay those of the oracle,
    and tensors handed out earlier are never overwritten."""
    b, P, gh, gw, T = 5, 16, 4, 5, 3
    rng = np.random.default_rng(77)
    u8 = synth_u8(b, 3, gh * P, gw * P, salt=2)
    boxes = random_boxes(rng, b, 2, gh * P, gw * P, 2 * P)
    orc = GazeOracle(to_f32(u8), boxes, P, T, 1, True)
    env = make_env(to_f32(u8).cuda(), torch.from_numpy(boxes), P, T, 1, True)
    start = np.stack([rng.integers(0, gh, b), rng.integers(0, gw, b)], 1).astype(np.int64)
    orc.reset(start); env.reset(torch.from_numpy(start))
    kept = []
    for t in range(3 * T + 2):
        a = rng.integers(0, 9, size=b).astype(np.int64)
        if t == 4:  # move by hand between two steps, like a caller driving the pieces itself
            env.apply_movements(torch.from_numpy(a).cuda())
            moved = orc.positions + np.array([(0, -1), (0, 1), (-1, 0), (1, 0), (-1, -1), (-1, 1), (1, -1), (1, 1), (0, 0)])[a]
            moved[:, 0] = np.clip(moved[:, 0], 0, gh - 1); moved[:, 1] = np.clip(moved[:, 1], 0, gw - 1)
            orc.positions = moved
            orc.has_stopped |= a == 8
            a = rng.integers(0, 9, size=b).astype(np.int64)
        o, e = orc.step(a), env.step(torch.from_numpy(a))
        assert torch.equal(e[0].cpu(), o[0]), t
        for k in (1, 2, 3):
            assert np.array_equal(e[k].cpu().numpy(), o[k]), (t, k)
        assert np.array_equal(e[4]["positions"].cpu().numpy(), o[4]["positions"])
        kept.append((e[1], e[4]["positions"], o[1].copy(), o[4]["positions"].copy()))
    for r, p, r_o, p_o in kept:  # every step's tensors still hold that step's values
        assert np.array_equal(r.cpu().numpy(), r_o) and np.array_equal(p.cpu().numpy(), p_o)
    with pytest.raises(RuntimeError):
        env.rollout_buffers()  # the rings only hold the last steps of an episode that ran past max_ep_len
