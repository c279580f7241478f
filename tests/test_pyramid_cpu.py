"""The host half of the glimpse pyramid: the antialiased-bilinear filter taps (``pyramid.aa_weights``) and the
order of operations the CUDA kernels follow (rows then columns, ``t = s0*w0; t = fma(s_j, w_j, t)``), checked on
the CPU against torch's own kernel -- the arithmetic the reference runs (general_env.py:95-111)."""
import numpy as np
import pytest
import torch

from jolineedle_b200.pyramid import aa_weights


def reflect(i, n):
    i = np.where(i < 0, -i, i)
    return np.where(i >= n, 2 * (n - 1) - i, i)


def fma_chain(values, weights):
    """values [..., n], weights [n]: the chain in float32 with fused multiply-adds (exact product in float64)."""
    t = (values[..., 0] * weights[0]).astype(np.float32)
    for j in range(1, len(weights)):
        t = (values[..., j].astype(np.float64) * np.float64(weights[j]) + t.astype(np.float64)).astype(np.float32)
    return t


def level_restated(x, pad):
    b, c, h, w = x.shape
    fx, cx, wx = aa_weights(w + 2 * pad, w)
    fy, cy, wy = aa_weights(h + 2 * pad, h)
    tmp = np.empty_like(x)
    for i in range(w):
        cols = reflect(fx[i] + np.arange(cx[i]) - pad, w)
        tmp[..., i] = fma_chain(x[..., cols], wx[i, :cx[i]])
    out = np.empty_like(x)
    for i in range(h):
        rows = reflect(fy[i] + np.arange(cy[i]) - pad, h)
        out[..., i, :] = fma_chain(np.moveaxis(tmp[..., rows, :], -2, -1), wy[i, :cy[i]])
    return out


@pytest.mark.skipif(torch.backends.cpu.get_cpu_capability() not in ("AVX2", "AVX512"),
                    reason="ATen's DEFAULT build does not contract multiply-adds")
@pytest.mark.parametrize("P,gh,gw", [(16, 5, 6), (32, 3, 4), (8, 9, 7), (56, 5, 6)])
def test_restated_level_equals_torchvision_bit_for_bit(P, gh, gw):
    import torchvision.transforms.functional as TF

    g = torch.Generator().manual_seed(P)
    h, w = gh * P, gw * P
    images = torch.randint(0, 256, (2, 3, h, w), dtype=torch.uint8, generator=g).float() / 255
    want = TF.resize(TF.pad(images, padding=[P] * 4, padding_mode="reflect"), size=[h, w], antialias=True)
    assert np.array_equal(level_restated(images.numpy(), P), want.numpy())


def test_filter_taps_are_normalised_and_inside_the_axis():
    for size, pad in ((96, 16), (2240, 448), (2688, 448), (8192, 256), (80, 16)):
        first, count, weights = aa_weights(size + 2 * pad, size)
        assert weights.dtype == np.float32 and weights.shape[0] == size and weights.shape[1] == 5
        assert (first >= 0).all() and (first + count <= size + 2 * pad).all() and (count >= 2).all()
        sums = np.array([weights[i, :count[i]].astype(np.float64).sum() for i in range(size)])
        assert np.allclose(sums, 1.0, atol=1e-6)
        assert all((weights[i, count[i]:] == 0).all() for i in range(size))


@pytest.mark.skipif(torch.backends.cpu.get_cpu_capability() not in ("AVX2", "AVX512"),
                    reason="ATen's DEFAULT build does not contract multiply-adds")
def test_uint8_levels_are_the_float_resize_rounded_half_to_even():
    """torchvision's route for uint8 tensors (cast to float32, resize, torch.round, cast back) -- what the uint8
    mode of jn_resize_aa_reflect follows -- level after level."""
    import torchvision.transforms.functional as TF

    P, h, w = 16, 80, 96
    g = torch.Generator().manual_seed(3)
    cur = torch.randint(0, 256, (2, 3, h, w), dtype=torch.uint8, generator=g)
    mine = cur.numpy()
    for _ in range(3):
        cur = TF.resize(TF.pad(cur, padding=[P] * 4, padding_mode="reflect"), size=[h, w], antialias=True)
        mine = np.rint(level_restated(mine.astype(np.float32), P)).astype(np.uint8)
        assert cur.dtype == torch.uint8 and np.array_equal(mine, cur.numpy())
