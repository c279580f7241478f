"""pad_to_patch: images whose sizes are not multiples of the patch behave exactly like their zero-padded
copies (complete_to_patch_size / padded_collate_fn, dataset.py:307-347,379-406), without the copy."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import focus_restatement, load_golden, random_boxes

pytestmark = pytest.mark.gpu


def padded(images, P):
    h, w = images.shape[-2:]
    return F.pad(images, (0, -w % P, 0, -h % P), mode="constant", value=0)


@pytest.mark.parametrize("engine", ["auto", "tensor", "ldg"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.uint8])
@pytest.mark.parametrize("P,h,w", [(64, 200, 272), (256, 600, 528), (448, 1000, 1136), (32, 70, 75)])
def test_padded_set_equals_gather_from_padded_copy(engine, dtype, P, h, w):
    from jolineedle_b200.gather import ImageSet

    elem = 1 if dtype == torch.uint8 else 4
    if engine == "tensor" and (w * elem) % 16:
        pytest.skip("rows that are not 16-byte multiples have no tensor map (plain loads serve them)")
    b = 3
    g = torch.Generator().manual_seed(P + h)
    u8 = torch.randint(0, 256, (b, 3, h, w), dtype=torch.uint8, generator=g)
    images = u8 if dtype == torch.uint8 else u8.float() / 255
    full = padded(images, P)
    gh, gw = full.shape[-2] // P, full.shape[-1] // P
    pos = torch.tensor([[y, x] for y in range(gh) for x in range(gw)], dtype=torch.int64)
    src = torch.arange(len(pos), dtype=torch.int32) % b
    s = ImageSet(images.cuda(), P, pad_to_patch=True)
    table = torch.from_numpy(load_golden("norm.npz")["u8_over_255"])
    u8_in = dtype == torch.uint8
    for normalize, focus in ((False, False), (u8_in, False), (u8_in, True), (False, True)):
        if focus and not normalize and u8_in:
            continue
        if engine == "tensor" and u8_in and not normalize:
            continue  # uint8 -> uint8 has no converting pass: plain loads
        want = torch.stack([full[int(k)][:, y * P:(y + 1) * P, x * P:(x + 1) * P] for (y, x), k in zip(pos.tolist(), src.tolist())])
        if normalize:
            want = table[want.long()]
        if focus:
            want = focus_restatement(want)
        got = s.gather(pos.cuda(), src_index=src.cuda(), normalize=normalize, focus=focus, engine=engine)
        assert torch.equal(got.cpu(), want), (engine, dtype, P, normalize, focus)
    # a position outside the padded grid is still reported
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    s.gather(torch.tensor([[gh, 0]], dtype=torch.int64).cuda(), normalize=u8_in, status=status, engine="auto")
    assert int(status.item()) == 1
    with pytest.raises(AssertionError):
        ImageSet(images.cuda(), P)  # without the flag the reference's precondition holds


def test_env_on_unpadded_images_equals_env_on_padded_copies():
    from jolineedle_b200.env.general_env import NeedleGeneralEnv

    b, P, h, w, T = 5, 64, 230, 304, 10  # LARD-like: neither side is a multiple of the patch
    g = torch.Generator().manual_seed(1)
    u8 = torch.randint(0, 256, (b, 3, h, w), dtype=torch.uint8, generator=g)
    rng = np.random.default_rng(4)
    boxes = torch.from_numpy(random_boxes(rng, b, 2, h, w, P))
    ref = NeedleGeneralEnv(padded(u8, P).cuda(), boxes, P, T, 1, True, normalize=True)
    env = NeedleGeneralEnv(u8.cuda(), boxes, P, T, 1, True, normalize=True, pad_to_patch=True)
    assert (env.n_vertical_patches, env.n_horizontal_patches) == (ref.n_vertical_patches, ref.n_horizontal_patches)
    assert torch.equal(env.bbox_masks, ref.bbox_masks)
    start = torch.from_numpy(np.stack([rng.integers(0, 4, b), rng.integers(0, 5, b)], 1).astype(np.int64))
    assert torch.equal(env.reset(start)[0], ref.reset(start)[0])
    for t in range(T):
        a = torch.from_numpy(rng.integers(0, 9, size=b).astype(np.int64))
        e, r = env.step(a), ref.step(a)
        for k in range(4):
            assert torch.equal(e[k], r[k]), (t, k)
        assert torch.equal(e[4]["positions"], r[4]["positions"])
    env.check_status()
    with pytest.raises(AssertionError):
        NeedleGeneralEnv(u8.cuda(), boxes, P, T, 1, True)


@pytest.mark.parametrize("engine", ["auto", "ldg"])
def test_padded_and_translated(engine):
    """Both zero fills at once: the image is translated (F.affine semantics) inside its own frame first, then
    padded to the patch grid -- what the reference's dataset does (transform, then padded_collate_fn)."""
    from jolineedle_b200.gather import ImageSet
    from oracle.gaze_oracle import translate_oracle

    P, h, w, b = 64, 150, 208, 4
    g = torch.Generator().manual_seed(8)
    u8 = torch.randint(0, 256, (b, 3, h, w), dtype=torch.uint8, generator=g)
    shifts_xy = np.array([(0, 0), (13, -7), (-40, 21), (70, 90)])
    want_images = padded(translate_oracle(u8, shifts_xy), P)
    gh, gw = want_images.shape[-2] // P, want_images.shape[-1] // P
    pos = torch.tensor([[y, x] for y in range(gh) for x in range(gw)] * b, dtype=torch.int64)
    src = torch.arange(b, dtype=torch.int32).repeat_interleave(gh * gw)
    s = ImageSet(u8.cuda(), P, pad_to_patch=True)
    d_shifts = torch.from_numpy(shifts_xy[:, ::-1].copy().astype(np.int32)).cuda()  # (ty, tx)
    table = torch.from_numpy(load_golden("norm.npz")["u8_over_255"])
    want = torch.stack([want_images[int(k)][:, y * P:(y + 1) * P, x * P:(x + 1) * P] for (y, x), k in zip(pos.tolist(), src.tolist())])
    got = s.gather(pos.cuda(), src_index=src.cuda(), shifts=d_shifts, normalize=True, engine=engine)
    assert torch.equal(got.cpu(), table[want.long()])
    got = s.gather(pos.cuda(), src_index=src.cuda(), shifts=d_shifts, normalize=True, focus=True, engine=engine)
    assert torch.equal(got.cpu(), focus_restatement(table[want.long()]))
