"""Translate-as-gather-offset: crops of the *translated* image, without ever
materialising it, must equal `F.affine(translate=[tx, ty], fill=0)` followed by the crop."""
import numpy as np
import pytest
import torch

from helpers import focus_restatement, load_golden, random_boxes
from oracle.gaze_oracle import GazeOracle, translate_oracle

pytestmark = pytest.mark.gpu


def crops_of(images, pos, src, P):
    return torch.stack([images[int(k)][:, y * P:(y + 1) * P, x * P:(x + 1) * P]
                        for (y, x), k in zip(pos.tolist(), src.tolist())])


@pytest.mark.parametrize("engine", ["auto", "tensor", "tensor-aligned", "ldg"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.uint8])
@pytest.mark.parametrize("P,gh,gw", [(64, 4, 5), (256, 3, 3), (448, 2, 3)])
def test_gather_with_translation(engine, dtype, P, gh, gw):
    """auto / tensor: arbitrary x offsets ride the TMA engine through 16-byte aligned superset loads
    (xform kernel); tensor-aligned: x offsets the caller vouches for, pure-DMA copy kernel; ldg: plain loads."""
    from jolineedle_b200.gather import ImageSet

    aligned = engine == "tensor-aligned"
    if aligned and P > 256:
        pytest.skip("the pure-DMA translated copy needs tiles of one TMA box (256 elements)")
    engine = "tensor" if aligned else engine
    b = 5
    g = torch.Generator().manual_seed(P)
    u8 = torch.randint(0, 256, (b, 3, gh * P, gw * P), dtype=torch.uint8, generator=g)
    images = u8 if dtype == torch.uint8 else u8.float() / 255
    # (tx, ty): none, small, negative, larger than a patch, and far enough to empty whole tiles
    shifts_xy = np.array([(0, 0), (13, -7), (-P // 3, P // 5), (P + 9, -(P + 3)), (-(gw * P - 5), gh * P - 2)])
    if aligned:  # x offsets that are multiples of 16 bytes; y is free
        shifts_xy[:, 0] = (shifts_xy[:, 0] // 16) * 16
    shifted = translate_oracle(images, shifts_xy)
    s = ImageSet(images.cuda(), P)
    n = 24
    rng = np.random.default_rng(3)
    pos = torch.from_numpy(np.stack([rng.integers(0, gh, n), rng.integers(0, gw, n)], 1).astype(np.int64))
    src = torch.from_numpy(rng.integers(0, b, n).astype(np.int32))
    d_shifts = torch.from_numpy(shifts_xy[:, ::-1].copy().astype(np.int32)).cuda()  # kernels take (ty, tx)
    table = torch.from_numpy(load_golden("norm.npz")["u8_over_255"])
    u8_in = dtype == torch.uint8
    for normalize, focus in ((False, False), (u8_in, False), (u8_in, True), (False, True)):
        if focus and not normalize and u8_in:
            continue  # uint8 -> uint8 Focus is nobody's layout
        if engine == "tensor" and not aligned and u8_in and not normalize:
            # uint8 -> uint8 copies have no register pass that could realign them: only vouched-for offsets
            with pytest.raises(ValueError):
                s.gather(pos.cuda(), src_index=src.cuda(), shifts=d_shifts, engine="tensor")
            continue
        want = crops_of(shifted, pos, src, P)
        if normalize:
            want = table[want.long()]
        if focus:
            want = focus_restatement(want)
        got = s.gather(pos.cuda(), src_index=src.cuda(), shifts=d_shifts, shifts_aligned=aligned, normalize=normalize,
                       focus=focus, engine=engine)
        assert torch.equal(got.cpu(), want), (engine, dtype, P, normalize, focus)


def test_bulk_engine_refuses_translation_and_lists_fall_back():
    from jolineedle_b200.gather import ImageSet

    P = 64
    imgs = [torch.rand(3, 2 * P, 3 * P), torch.rand(3, 3 * P, 2 * P)]
    s = ImageSet([t.cuda() for t in imgs], P)
    pos = torch.tensor([[1, 2], [2, 0]], dtype=torch.int64).cuda()
    shifts = torch.tensor([[5, -9], [-70, 3]], dtype=torch.int32).cuda()  # (ty, tx)
    with pytest.raises(ValueError):
        s.gather(pos, shifts=shifts, engine="bulk")
    with pytest.raises(ValueError):  # lists of images have no single tensor map
        s.gather(pos, shifts=shifts, engine="tensor")
    got = s.gather(pos, shifts=shifts)  # auto -> plain loads for a list of images
    for i, im in enumerate(imgs):
        ty, tx = shifts[i].tolist()
        sh = translate_oracle(im[None], [(tx, ty)])[0]
        y, x = pos[i].tolist()
        assert torch.equal(got[i].cpu(), sh[:, y * P:(y + 1) * P, x * P:(x + 1) * P])


@pytest.mark.parametrize("P,gh,gw,T", [(256, 4, 4, 10), (448, 3, 3, 6)])
def test_env_with_translate_equals_env_on_shifted_images(P, gh, gw, T):
    """cfg-4 style augment-translate: NeedleGeneralEnv(images, boxes + shift, translate=shift) behaves
    exactly like the reference env fed the F.affine-translated images and the shifted boxes."""
    from jolineedle_b200.env.general_env import NeedleGeneralEnv

    b = 4
    g = torch.Generator().manual_seed(7)
    u8 = torch.randint(0, 256, (b, 3, gh * P, gw * P), dtype=torch.uint8, generator=g)
    rng = np.random.default_rng(5)
    boxes = random_boxes(rng, b, 2, gh * P, gw * P, P)
    shifts_xy = np.stack([rng.integers(-P // 2, P // 2, b), rng.integers(-P // 2, P // 2, b)], 1)
    moved_boxes = boxes + np.concatenate([shifts_xy, shifts_xy], axis=1)[:, None, :]  # dataset.py:215-225
    images = u8.float() / 255
    orc = GazeOracle(translate_oracle(images, shifts_xy), moved_boxes, P, T, 1, True)
    env = NeedleGeneralEnv(u8.cuda(), torch.from_numpy(moved_boxes), P, T, 1, True, normalize=True,
                           translate=torch.from_numpy(shifts_xy))
    assert np.array_equal(env.bbox_masks.cpu().numpy(), orc.bbox_masks)
    start = np.stack([rng.integers(0, gh, b), rng.integers(0, gw, b)], 1).astype(np.int64)
    p_o, _ = orc.reset(start)
    p_e, _ = env.reset(torch.from_numpy(start))
    assert torch.equal(p_e.cpu(), p_o)
    for t in range(T):
        a = rng.integers(0, 9, size=b).astype(np.int64)
        o, e = orc.step(a), env.step(torch.from_numpy(a))
        assert torch.equal(e[0].cpu(), o[0]) and np.array_equal(e[1].cpu().numpy(), o[1])
        assert np.array_equal(e[2].cpu().numpy(), o[2]) and np.array_equal(e[4]["positions"].cpu().numpy(), o[4]["positions"])
    env.check_status()
