"""K1 parity: every gather engine against plain torch slicing (bit-exact), through the C ABI."""
import numpy as np
import pytest
import torch

from helpers import focus_restatement, load_golden, synth_u8

pytestmark = pytest.mark.gpu


def ref_gather(images, positions, src, P, normalize, focus):
    """images: list of CPU [C,H,W] tensors; returns the expected CPU output."""
    out = []
    for i, (y, x) in enumerate(positions.tolist()):
        k = i if src is None else int(src[i])
        if k < 0:
            tile = torch.zeros_like(images[0][:, :P, :P])
        else:
            tile = images[k][:, y * P:(y + 1) * P, x * P:(x + 1) * P]
        if normalize:
            tile = tile.float() / 255
        if focus:
            tile = focus_restatement(tile)
        out.append(tile)
    return torch.stack(out)


def make_images(b, h, w, dtype, salt=0):
    u8 = torch.from_numpy(synth_u8(b, 3, h, w, salt))
    return u8 if dtype == torch.uint8 else u8.float() / 255


MODES = [(False, False), (True, False), (False, True), (True, True)]


@pytest.mark.parametrize("engine", ["tensor", "bulk", "ldg", "auto"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.uint8])
@pytest.mark.parametrize("P,gh,gw,b", [(16, 5, 6, 7), (32, 3, 4, 5), (64, 2, 3, 3), (448, 2, 3, 2), (256, 2, 2, 2)])
def test_gather_engines_bit_exact(engine, dtype, P, gh, gw, b):
    from jolineedle_b200.gather import ImageSet

    imgs = make_images(b, gh * P, gw * P, dtype, salt=P)
    dev = imgs.cuda()
    s = ImageSet(dev, P)
    rng = np.random.default_rng(P + b)
    n = b
    pos = torch.from_numpy(np.stack([rng.integers(0, gh, n), rng.integers(0, gw, n)], 1).astype(np.int64))
    for normalize, focus in MODES:
        if normalize and dtype != torch.uint8:
            continue
        if focus and not normalize and dtype == torch.uint8 and engine in ("tensor", "bulk"):
            continue  # u8 -> u8 Focus only exists on the LDG engine
        got = s.gather(pos.cuda(), normalize=normalize, focus=focus, engine=engine)
        want = ref_gather(list(imgs), pos, None, P, normalize, focus)
        torch.cuda.synchronize()
        assert got.dtype == want.dtype and tuple(got.shape) == tuple(want.shape)
        assert torch.equal(got.cpu(), want), (engine, dtype, P, normalize, focus)


@pytest.mark.parametrize("engine", ["tensor", "bulk", "ldg"])
def test_src_index_zero_fill_and_strided_output(engine):
    from jolineedle_b200.gather import ImageSet

    P, gh, gw, b, T = 32, 3, 4, 4, 5
    imgs = make_images(b, gh * P, gw * P, torch.float32, salt=3)
    s = ImageSet(imgs.cuda(), P)
    rng = np.random.default_rng(0)
    n = b * T
    pos = torch.from_numpy(np.stack([rng.integers(0, gh, n), rng.integers(0, gw, n)], 1).astype(np.int64))
    src = torch.from_numpy(np.repeat(np.arange(b), T).astype(np.int32))
    src[3] = -1; src[n - 1] = -1; src[7] = -1
    hist = torch.full((b, T, 3, P, P), 7.0, device="cuda")
    s.gather(pos.cuda(), src_index=src.cuda(), out=hist.view(n, 3, P, P), engine=engine)
    want = ref_gather(list(imgs), pos, src, P, False, False).view(b, T, 3, P, P)
    assert torch.equal(hist.cpu(), want)
    # a time slot of a history buffer: item stride = T tiles
    slot = hist[:, 2]
    pos_b = pos[:b].cuda()
    s.gather(pos_b, out=slot, engine=engine)
    assert torch.equal(slot.cpu(), ref_gather(list(imgs), pos[:b], None, P, False, False))
    assert torch.equal(hist[:, 1].cpu(), want[:, 1])  # neighbours untouched


@pytest.mark.parametrize("engine", ["bulk", "ldg", "auto"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.uint8])
def test_list_of_images_of_different_sizes(engine, dtype):
    from jolineedle_b200.gather import ImageSet

    P = 32
    sizes = [(3, 4), (2, 2), (4, 3), (1, 5)]
    imgs = [make_images(1, gh * P, gw * P, dtype, salt=i)[0] for i, (gh, gw) in enumerate(sizes)]
    s = ImageSet([t.cuda() for t in imgs], P)
    items = [(0, 2, 3), (1, 1, 1), (2, 3, 0), (3, 0, 4), (2, 0, 2), (-1, 0, 0), (0, 0, 0)]
    src = torch.tensor([i[0] for i in items], dtype=torch.int32)
    pos = torch.tensor([[i[1], i[2]] for i in items], dtype=torch.int64)
    normalize = dtype == torch.uint8
    got = s.gather(pos.cuda(), src_index=src.cuda(), normalize=normalize, engine=engine)
    assert torch.equal(got.cpu(), ref_gather(imgs, pos, src, P, normalize, False))


def test_unaligned_shapes_fall_back_to_ldg():
    from jolineedle_b200.gather import ImageSet

    P, gh, gw = 10, 3, 5  # 10-byte uint8 rows: no TMA engine can address them
    imgs = make_images(2, gh * P, gw * P, torch.uint8, salt=9)
    s = ImageSet(imgs.cuda(), P)
    assert not s.engine_available("bulk") and not s.engine_available("tensor")
    pos = torch.tensor([[2, 4], [0, 1]], dtype=torch.int64)
    for normalize, focus in MODES:
        got = s.gather(pos.cuda(), normalize=normalize, focus=focus)
        assert torch.equal(got.cpu(), ref_gather(list(imgs), pos, None, P, normalize, focus))
    with pytest.raises(ValueError):
        s.gather(pos.cuda(), engine="bulk")


def test_normalisation_is_totensor_exact_for_every_byte():
    from jolineedle_b200.gather import ImageSet

    P = 16
    img = torch.arange(256, dtype=torch.uint8).view(1, 1, 16, 16).repeat(1, 3, 1, 1)
    s = ImageSet(img.cuda(), P)
    pos = torch.zeros((1, 2), dtype=torch.int64, device="cuda")
    table = torch.from_numpy(load_golden("norm.npz")["u8_over_255"])
    for engine in ("tensor", "bulk", "ldg"):
        got = s.gather(pos, normalize=True, engine=engine).cpu()
        assert torch.equal(got[0, 0].flatten(), table), engine


def test_out_of_grid_position_is_flagged_not_read():
    from jolineedle_b200.gather import ImageSet

    P = 16
    imgs = make_images(2, 2 * P, 2 * P, torch.float32)
    s = ImageSet(imgs.cuda(), P)
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    pos = torch.tensor([[0, 1], [2, 0]], dtype=torch.int64, device="cuda")
    for engine in ("tensor", "bulk", "ldg"):
        status.zero_()
        out = torch.full((2, 3, P, P), -1.0, device="cuda")
        s.gather(pos, out=out, engine=engine, status=status)
        assert int(status.item()) & 1
        assert torch.equal(out[0].cpu(), imgs[0][:, :P, P:2 * P])
        assert bool((out[1] == -1).all())  # skipped tile untouched


def test_bad_image_index_on_a_list_of_images_is_flagged_not_dereferenced():
    """src_index beyond the set: reported through the status word, the tile skipped -- on multi-slab sets the
    per-image record of such an index must not be loaded at all (it lies outside the table)."""
    from jolineedle_b200.gather import ImageSet

    P = 16
    imgs = [make_images(1, 2 * P, 3 * P, torch.float32, salt=i)[0] for i in range(3)]
    s = ImageSet([t.cuda() for t in imgs], P)
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    pos = torch.tensor([[1, 2], [0, 0], [1, 1]], dtype=torch.int64, device="cuda")
    src = torch.tensor([2, 1 << 20, 0], dtype=torch.int32, device="cuda")
    for engine in ("bulk", "ldg", "auto"):
        status.zero_()
        out = torch.full((3, 3, P, P), -1.0, device="cuda")
        s.gather(pos, src_index=src, out=out, engine=engine, status=status)
        torch.cuda.synchronize()
        assert int(status.item()) & 1, engine
        assert torch.equal(out[0].cpu(), imgs[2][:, P:2 * P, 2 * P:3 * P])
        assert bool((out[1] == -1).all()) and torch.equal(out[2].cpu(), imgs[0][:, P:2 * P, P:2 * P])


def test_pinned_images_must_be_contiguous():
    from jolineedle_b200.gather import ImageSet

    host = torch.zeros((2, 3, 32, 64), dtype=torch.uint8).pin_memory()
    with pytest.raises(ValueError):
        ImageSet(host[:, :, :, ::2], 16, device="cuda")


def test_size_mismatch_raises_like_the_reference():
    from jolineedle_b200.gather import ImageSet

    with pytest.raises(AssertionError):
        ImageSet(torch.zeros(1, 3, 100, 96, device="cuda"), 16)


@pytest.mark.parametrize("dtype,normalize", [(torch.float32, False), (torch.uint8, True), (torch.uint8, False)])
def test_full_size_round_trip_lard_shape(dtype, normalize):
    """cfg-2/3 geometry (2240x2688, P=448): gathering all 30 patches of every image and pasting
    them back must reproduce the images (size-independent property, checked on the device)."""
    from jolineedle_b200.gather import ImageSet

    b, P, gh, gw = 12, 448, 5, 6
    g = torch.Generator(device="cuda").manual_seed(1234)
    u8 = torch.randint(0, 256, (b, 3, gh * P, gw * P), dtype=torch.uint8, device="cuda", generator=g)
    imgs = u8 if dtype == torch.uint8 else u8.float() / 255
    s = ImageSet(imgs, P)
    ys, xs, bs = torch.meshgrid(torch.arange(gh), torch.arange(gw), torch.arange(b), indexing="ij")
    pos = torch.stack([ys.flatten(), xs.flatten()], 1).cuda()
    src = bs.flatten().to(torch.int32).cuda()
    # NB: `x.float() / 255` on a CUDA tensor multiplies by a reciprocal; the reference's ToTensor divides on
    # the CPU.  The 256-entry table (tests/golden/norm.npz, produced by torch CPU) is the exact expectation.
    table = torch.from_numpy(load_golden("norm.npz")["u8_over_255"]).cuda()
    want = table[imgs.long()] if normalize else imgs
    for engine in ("tensor", "bulk"):
        tiles = s.gather(pos, src_index=src, normalize=normalize, engine=engine)
        back = tiles.view(gh, gw, b, 3, P, P).permute(2, 3, 0, 4, 1, 5).reshape(b, 3, gh * P, gw * P)
        assert torch.equal(back, want), engine
        # Focus layout == the Focus slicing of the plain crop
        if dtype == torch.float32 or normalize:
            f = s.gather(pos[:64], src_index=src[:64], normalize=normalize, focus=True, engine=engine)
            assert torch.equal(f, focus_restatement(tiles[:64])), engine


@pytest.mark.parametrize("engine", ["ldg", "bulk", "tensor"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.uint8])
def test_zero_copy_gather_from_pinned_host_memory(engine, dtype):
    """Pinned host images are gathered in place (no upload): only the glimpsed tiles cross PCIe."""
    from jolineedle_b200.gather import ImageSet

    P, gh, gw, b = 64, 3, 4, 5
    imgs = make_images(b, gh * P, gw * P, dtype, salt=11)
    normalize = dtype == torch.uint8
    rng = np.random.default_rng(2)
    pos = torch.from_numpy(np.stack([rng.integers(0, gh, b), rng.integers(0, gw, b)], 1).astype(np.int64))
    want = ref_gather(list(imgs), pos, None, P, normalize, False)
    # one pinned batch tensor
    s = ImageSet(imgs.pin_memory(), P, device="cuda")
    assert s.host_mapped
    assert torch.equal(s.gather(pos.cuda(), normalize=normalize, engine=engine).cpu(), want)
    # a list of separately pinned images (the DataLoader case)
    if engine != "tensor":
        s2 = ImageSet([im.clone().pin_memory() for im in imgs], P, device="cuda")
        assert torch.equal(s2.gather(pos.cuda(), normalize=normalize, engine=engine).cpu(), want)
    # pageable host memory is refused
    from jolineedle_b200 import _cabi

    with pytest.raises(_cabi.NativeLibraryError):
        ImageSet(imgs.clone(), P, device="cuda")


def test_converting_gather_back_to_back_and_on_two_streams():
    """The converting kernel claims its chunks from a global counter that it leaves at zero for the next
    launch; launches rotate through a pool of counters so that two streams never share one.  Many launches
    of different sizes back to back, then interleaved on two streams, must all be exact."""
    from jolineedle_b200.gather import ImageSet

    P, gh, gw, b = 64, 4, 5, 6
    imgs = make_images(b, gh * P, gw * P, torch.uint8, salt=3)
    s = ImageSet(imgs.cuda(), P)
    rng = np.random.default_rng(11)
    cases = []
    for n in (1, 3, 17, 64, 257, 1024, 5, 2048, 2):
        pos = torch.from_numpy(np.stack([rng.integers(0, gh, n), rng.integers(0, gw, n)], 1).astype(np.int64))
        src = torch.from_numpy(rng.integers(-1, b, n).astype(np.int32))
        focus = bool(n % 2)
        cases.append((pos.cuda(), src.cuda(), focus, ref_gather(list(imgs), pos, src, P, True, focus)))
    outs = [s.gather(p, src_index=k, normalize=True, focus=f) for (p, k, f, _) in cases for _ in range(3)]
    torch.cuda.synchronize()
    for i, o in enumerate(outs):
        assert torch.equal(o.cpu(), cases[i // 3][3]), i
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    outs = []
    for rep in range(4):
        for i, (p, k, f, _) in enumerate(cases):
            with torch.cuda.stream(streams[(i + rep) % 2]):
                outs.append((i, s.gather(p, src_index=k, normalize=True, focus=f)))
    torch.cuda.synchronize()
    for i, o in outs:
        assert torch.equal(o.cpu(), cases[i][3]), i
