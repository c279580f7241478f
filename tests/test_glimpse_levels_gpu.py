"""Multi-level glimpses (n_glimps_levels > 1, general_env.py:84-115) against a fixture of the unmodified
reference.  Level 0 is a bit-exact crop; higher levels go through torchvision's antialiased resize, which on the
GPU agrees with the reference's CPU result to float rounding: tolerance 2e-6 absolute on values in [0, 1]
(the north star's 1e-6 relative for floats, widened by one ulp of 1.0 for the resampling sum)."""
import numpy as np
import pytest
import torch

from helpers import load_golden, scenario, to_f32

pytestmark = pytest.mark.gpu

ATOL = 2e-6


@pytest.mark.parametrize("name", ["lv2", "lv3"])
@pytest.mark.parametrize("history", [False, True])
def test_glimpse_pyramid_matches_reference(name, history):
    from jolineedle_b200.env.general_env import NeedleGeneralEnv

    c = scenario(load_golden("glimpse_levels.npz"), name)
    levels = int(c["levels"])
    env = NeedleGeneralEnv(to_f32(c["u8"]).cuda(), torch.from_numpy(c["boxes"]), 16, 6, levels, True, history=history)
    assert tuple(env.images.shape) == tuple(c["images"].shape)
    assert torch.equal(env.images[:, 0].cpu(), torch.from_numpy(c["images"][:, 0]))  # level 0 untouched
    assert np.abs(env.images.cpu().numpy() - c["images"]).max() <= ATOL
    patches, _ = env.reset(torch.from_numpy(c["start"]))
    got = [patches]
    for t, a in enumerate(c["actions"]):
        out = env.step(torch.from_numpy(a))
        got.append(out[0])
        assert np.array_equal(out[1].cpu().numpy(), c["rewards"][t])  # rewards do not depend on the levels
    for t, p in enumerate(got):
        want = c["patches"][t]
        assert tuple(p.shape) == tuple(want.shape) == (3, levels, 3, 16, 16)
        assert torch.equal(p[:, 0].cpu(), torch.from_numpy(want[:, 0])), t  # level 0: bit-exact
        assert np.abs(p.cpu().numpy() - want).max() <= ATOL, t
    if history:
        hist = env.patch_history()
        assert tuple(hist.shape) == (3, (len(c["actions"]) + 1) * levels, 3, 16, 16)
        assert torch.equal(hist[:, -levels:], got[-1])
    env.check_status()
