"""Glimpse pyramid (``NeedleGeneralEnv.init_glimps_images``, general_env.py:84-115) without torchvision.

Level k+1 = level k reflect-padded by one patch on every side and resized back to ``H x W`` with torch's
antialiased bilinear filter.  The reference computes that on the CPU; ATen's kernel (``_upsample_bilinear2d_aa``)
is separable -- rows first, then columns -- with per-output-pixel weights computed in float32 and each output a
chain ``t = s0 * w0; t = fma(s_j, w_j, t)`` (its AVX2 / AVX-512 builds contract the multiply-add; that is what
every x86-64 server dispatches to).  :func:`aa_weights` restates the weight computation operation by operation,
``jn_resize_aa_reflect`` (csrc/jn_pyramid.cuh) runs the two passes with ``fmaf`` in the same order and folds the
reflect padding into its index arithmetic, so the levels equal the reference's bit for bit
(tests/test_glimpse_levels_gpu.py against a fixture of the unmodified reference).  uint8 images take the route
torchvision gives them: cast to float32, the same resize, ``torch.round`` (half to even), cast back.
"""
from typing import Tuple

import numpy as np
import torch

from . import _cabi


def aa_weights(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """``(first, count, weights[out_size, k])`` of the antialiased bilinear filter for one axis, as
    ``HelperInterpBase::_compute_indices_min_size_weights_aa`` computes them for float32 tensors
    (align_corners=False, no explicit scale): float32 values, the same mixed float / double intermediate
    steps, the same truncating casts."""
    f32, f64 = np.float32, np.float64
    scale = f32(f32(in_size) / f32(out_size))  # area_pixel_compute_scale<float>
    support = f32(f64(1.0) * f64(scale)) if scale >= 1.0 else f32(1.0)  # (interp_size * 0.5) * scale, interp_size = 2
    invscale = f32(f64(1.0) / f64(scale)) if scale >= 1.0 else f32(1.0)
    k = int(np.ceil(support)) * 2 + 1
    first = np.zeros(out_size, dtype=np.int32)
    count = np.zeros(out_size, dtype=np.int32)
    weights = np.zeros((out_size, k), dtype=np.float32)
    for i in range(out_size):
        center = f32(f64(scale) * (f64(i) + 0.5))
        lo = int(f64(f32(center - support)) + 0.5)  # static_cast<int64_t>: truncation toward zero
        hi = int(f64(f32(center + support)) + 0.5)
        xmin = max(lo, 0)
        n = min(hi, in_size) - xmin
        total = f32(0.0)
        for j in range(n):
            x = f32((f64(f32(f32(j + xmin) - center)) + 0.5) * f64(invscale))
            x = f32(abs(x))
            w = f32(f32(1.0) - x) if x < 1.0 else f32(0.0)
            weights[i, j] = w
            total = f32(total + w)
        if total != 0:
            weights[i, :n] = weights[i, :n] / total  # float32 division per weight
        first[i], count[i] = xmin, n
    return first, count, weights


def build_levels(images: torch.Tensor, patch_size: int, n_levels: int) -> torch.Tensor:
    """``[B, n_levels, C, H, W]`` stack of progressively zoomed-out copies of float32 or uint8 CUDA
    ``images [B, C, H, W]`` (level 0 = the input).  uint8 levels are what torchvision makes of uint8 tensors: the
    float32 resize of the bytes, rounded half to even, level after level."""
    _cabi.require_cuda(images, "images")
    code = _cabi.dtype_code(images.dtype)
    elem = images.element_size()
    b, c, h, w = images.shape
    if patch_size >= h or patch_size >= w:
        raise ValueError("reflect padding needs patch_size < image size")
    dev = images.device
    out = torch.empty((b, n_levels, c, h, w), dtype=images.dtype, device=dev)
    out[:, 0] = images
    if n_levels == 1:
        return out
    tables = []
    for in_size, out_size in ((w + 2 * patch_size, w), (h + 2 * patch_size, h)):
        first, count, weights = aa_weights(in_size, out_size)
        tables.append((torch.from_numpy(first).to(dev), torch.from_numpy(count).to(dev),
                       torch.from_numpy(weights).to(dev), weights.shape[1]))
    (fx, cx, wx, kx), (fy, cy, wy, ky) = tables
    tmp = torch.empty((b, c, h, w), dtype=torch.float32, device=dev)  # rows already resized, unpadded height
    lib = _cabi.lib()
    with _cabi.on_device(dev):
        for level in range(1, n_levels):
            src, dst = out[:, level - 1], out[:, level]
            _cabi.check(lib.jn_resize_aa_reflect(
                src.data_ptr(), src.stride(0) * elem, tmp.data_ptr(), dst.data_ptr(), dst.stride(0) * elem, code, b, c, h, w,
                patch_size, fx.data_ptr(), cx.data_ptr(), wx.data_ptr(), kx, fy.data_ptr(), cy.data_ptr(),
                wy.data_ptr(), ky, _cabi.stream_ptr(dev)))
    return out
