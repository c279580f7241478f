"""Batched env on pinned HOST images (zero_copy=True): crops, rewards, flags and positions must equal the
device-resident env's, and a patch crosses PCIe once per episode (jn_visit_sources)."""
import numpy as np
import pytest
import torch

from helpers import random_boxes

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("normalize,focus", [(True, False), (True, True), (False, False), (False, True)])
def test_zero_copy_env_equals_resident_env(normalize, focus):
    from jolineedle_b200.env.general_env import NeedleGeneralEnv

    b, P, gh, gw, T = 6, 64, 3, 4, 12
    g = torch.Generator().manual_seed(5)
    u8 = torch.randint(0, 256, (b, 3, gh * P, gw * P), dtype=torch.uint8, generator=g)
    images = u8 if normalize else u8.float() / 255
    rng = np.random.default_rng(2)
    boxes = torch.from_numpy(random_boxes(rng, b, 2, gh * P, gw * P, P))
    ref = NeedleGeneralEnv(images.cuda(), boxes, P, T, 1, True, normalize=normalize, focus=focus, history=True)
    env = NeedleGeneralEnv(images.pin_memory(), boxes, P, T, 1, True, normalize=normalize, focus=focus, history=True,
                           device="cuda", zero_copy=True)
    start = torch.from_numpy(np.stack([rng.integers(0, gh, b), rng.integers(0, gw, b)], 1).astype(np.int64))
    p_r, _ = ref.reset(start)
    p_e, _ = env.reset(start)
    assert torch.equal(p_e, p_r)
    seen = [{tuple(start[i].tolist())} for i in range(b)]
    for t in range(T):
        a = torch.from_numpy(rng.integers(0, 9, size=b).astype(np.int64))
        r, e = ref.step(a), env.step(a)
        for k in range(4):
            assert torch.equal(e[k], r[k]), (t, k)
        assert torch.equal(e[4]["positions"], r[4]["positions"])
        for i, yx in enumerate(e[4]["positions"].tolist()):
            seen[i].add(tuple(yx))
    assert torch.equal(env.patch_history(), ref.patch_history())
    assert int(env.host_tiles) == sum(len(s) for s in seen) < b * (T + 1)
    env.check_status()
    # a second episode on the same env starts with an empty first-visit table
    p_e, _ = env.reset(start)
    assert torch.equal(p_e, p_r)


def test_zero_copy_needs_pinned_images():
    from jolineedle_b200.env.general_env import NeedleGeneralEnv

    images = torch.zeros(2, 3, 64, 64)
    with pytest.raises(ValueError):
        NeedleGeneralEnv(images, torch.zeros(2, 1, 4, dtype=torch.int64), 32, 4, 1, device="cuda", zero_copy=True)
