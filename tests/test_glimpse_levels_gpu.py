"""Multi-level glimpses (n_glimps_levels > 1, general_env.py:84-115) against a fixture of the unmodified
reference: every level -- built by jn_resize_aa_reflect, the reference's CPU torchvision pad + antialiased resize
restated tap by tap -- and every crop taken from it, BIT for BIT."""
import numpy as np
import pytest
import torch

from helpers import load_golden, scenario, to_f32

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["lv2", "lv3"])
@pytest.mark.parametrize("history", [False, True])
def test_glimpse_pyramid_matches_reference(name, history):
    from jolineedle_b200.env.general_env import NeedleGeneralEnv

    c = scenario(load_golden("glimpse_levels.npz"), name)
    levels = int(c["levels"])
    env = NeedleGeneralEnv(to_f32(c["u8"]).cuda(), torch.from_numpy(c["boxes"]), 16, 6, levels, True, history=history)
    assert tuple(env.images.shape) == tuple(c["images"].shape)
    assert torch.equal(env.images[:, 0].cpu(), torch.from_numpy(c["images"][:, 0]))  # level 0 untouched
    assert np.array_equal(env.images.cpu().numpy(), c["images"])  # every level, bit for bit
    patches, _ = env.reset(torch.from_numpy(c["start"]))
    got = [patches]
    for t, a in enumerate(c["actions"]):
        out = env.step(torch.from_numpy(a))
        got.append(out[0])
        assert np.array_equal(out[1].cpu().numpy(), c["rewards"][t])  # rewards do not depend on the levels
    for t, p in enumerate(got):
        want = c["patches"][t]
        assert tuple(p.shape) == tuple(want.shape) == (3, levels, 3, 16, 16)
        assert np.array_equal(p.cpu().numpy(), want), t
    if history:
        hist = env.patch_history()
        assert tuple(hist.shape) == (3, (len(c["actions"]) + 1) * levels, 3, 16, 16)
        assert torch.equal(hist[:, -levels:], got[-1])
    env.check_status()


def test_uint8_glimpse_pyramid_matches_reference():
    """uint8 images (the dtype the reference's docstring names, general_env.py:28): torchvision resizes them
    through float32 + torch.round, level after level; plain uint8 crops and normalised ones."""
    from jolineedle_b200.env.general_env import NeedleGeneralEnv

    c = scenario(load_golden("glimpse_levels.npz"), "lv3u8")
    levels = int(c["levels"])
    table = torch.from_numpy(load_golden("norm.npz")["u8_over_255"])
    for normalize in (False, True):
        env = NeedleGeneralEnv(torch.from_numpy(c["u8"]).cuda(), torch.from_numpy(c["boxes"]), 16, 6, levels, True,
                               normalize=normalize)
        assert env.images.dtype == torch.uint8 and np.array_equal(env.images.cpu().numpy(), c["images"])
        got = [env.reset(torch.from_numpy(c["start"]))[0]]
        for t, a in enumerate(c["actions"]):
            out = env.step(torch.from_numpy(a))
            got.append(out[0])
            assert np.array_equal(out[1].cpu().numpy(), c["rewards"][t])
        for t, p in enumerate(got):
            want = torch.from_numpy(c["patches"][t])
            assert torch.equal(p.cpu(), table[want.long()] if normalize else want), (t, normalize)
        env.check_status()


@pytest.mark.parametrize("as_bytes", [False, True])
@pytest.mark.parametrize("P,gh,gw,levels", [(32, 4, 5, 3), (448, 5, 6, 2), (64, 7, 3, 4)])
def test_pyramid_levels_equal_torchvision_cpu(P, gh, gw, levels, as_bytes):
    """Other shapes (incl. the LARD geometry), straight against the reference's own calls on the CPU."""
    import torchvision.transforms.functional as TF

    from jolineedle_b200.pyramid import build_levels

    g = torch.Generator().manual_seed(P + levels)
    images = torch.randint(0, 256, (2, 3, gh * P, gw * P), dtype=torch.uint8, generator=g)
    if not as_bytes:
        images = images.float() / 255
    want, cur = [images], images
    for _ in range(levels - 1):  # general_env.py:95-111
        cur = TF.resize(TF.pad(cur, padding=[P] * 4, padding_mode="reflect"), size=[gh * P, gw * P], antialias=True)
        want.append(cur)
    got = build_levels(images.cuda(), P, levels)
    assert torch.equal(got.cpu(), torch.stack(want, dim=1))
