"""Batched supervised trajectories: host plans -> device tensors (reference:
``SupervisedTrainer.generate_trajectories``, ``src/supervised.py:95-136``).

Two planners produce the same :class:`PackedPlans`:

  ``native``  ``jn_plan_run`` (csrc/jn_planner.cpp): the reference's random streams (numpy
              PCG64 / python MT19937 / CPython set order) restated in C++; ~1 us per episode.
  ``python``  ``NeedleSimpleEnv.plan_sample``: calls the real numpy / random / set objects;
              used for inputs the native planner does not cover (float boxes, seeds above
              2^64) and as its cross-check in the tests.

The device half (:func:`expand_packed`) is shared: K0 5 %-area bitmaps, K3 ``jn_traj_expand``,
``jn_local_boxes`` and the K1 gathers, all on torch's current stream.
"""
import ctypes
import random
from itertools import chain
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from .. import _cabi
from ..gather import ImageSet
from ..utils import Position

_chain = chain.from_iterable


class PackedPlans:
    """Flat host arrays describing ``n`` planned episodes (see ``jn_traj_expand``)."""

    __slots__ = ("n", "start", "seg_begin", "seg_to", "seg_tgt", "seg_flags", "draw_begin", "draws", "det_begin",
                 "det_yx", "rows", "cols", "n_boxes", "boxes", "n_max", "boxes_f64")

    def __init__(self):
        self.boxes_f64 = None  # [n, n_max, 4] float64 when some box coordinate is not a whole pixel


# ---------------------------------------------------------------------------------------------------
# python planner -> packed
# ---------------------------------------------------------------------------------------------------
def boxes_array(bboxes: Sequence[Sequence], n_max: Optional[int] = None, want_float: bool = False):
    """``[n, max(n_max, 1), 4]`` int64 x1,y1,x2,y2 + per-image counts; ``exact`` tells whether every
    coordinate was an integer (the native planner works on integer pixels).  ``want_float``: a fifth value,
    the same array in float64 when the boxes are not exact (None otherwise) -- the reference keeps such boxes
    as python floats (the dataset's minimum-size resize scales them, dataset.py:258-270), so labels and local
    boxes must come from the un-truncated coordinates."""
    n = len(bboxes)
    counts = np.array([len(b) for b in bboxes], dtype=np.int32)
    n_max = int(counts.max()) if n_max is None and n else (n_max or 0)
    arr = np.zeros((n, max(n_max, 1), 4), dtype=np.int64)
    total = int(counts.sum())
    if total == 0:
        return (arr, counts, n_max, True, None) if want_float else (arr, counts, n_max, True)
    # one conversion for the whole batch: BBox = ((y1, x1), (y2, x2)) -> flat list of 4 * total numbers
    # (flattening in python first is ~10x faster than letting numpy walk the nested tuples)
    # straight into a float64 buffer, no intermediate list (pixel coordinates are exact in float64)
    flat = np.fromiter(_chain(_chain(_chain(bboxes))), dtype=np.float64, count=4 * total)
    exact = bool(np.all(flat == np.floor(flat)))
    flat = flat.reshape(total, 4)[:, [1, 0, 3, 2]]  # -> x1, y1, x2, y2
    image = np.repeat(np.arange(n), counts)
    slot = np.arange(total) - np.repeat(np.cumsum(counts) - counts, counts)
    arr[image, slot] = flat.astype(np.int64)  # (truncation like `.int()`; only used when exact)
    if not want_float:
        return arr, counts, n_max, exact
    arr_f = None
    if not exact:
        arr_f = np.zeros(arr.shape, dtype=np.float64)
        arr_f[image, slot] = flat.astype(np.float64)
    return arr, counts, n_max, exact, arr_f


def pack_python_plans(envs, plans) -> PackedPlans:
    n = len(envs)
    p = PackedPlans()
    p.n = n
    p.seg_begin = np.zeros(n + 1, dtype=np.int32)
    p.draw_begin = np.zeros(n + 1, dtype=np.int32)
    p.det_begin = np.zeros(n + 1, dtype=np.int32)
    for i, pl in enumerate(plans):
        p.seg_begin[i + 1] = p.seg_begin[i] + len(pl.seg_to)
        p.draw_begin[i + 1] = p.draw_begin[i] + len(pl.draws)
        p.det_begin[i + 1] = p.det_begin[i] + len(pl.det_positions)
    n_seg, n_det = int(p.seg_begin[-1]), int(p.det_begin[-1])
    p.start = np.array([pl.start for pl in plans], dtype=np.int32).reshape(n, 2)
    p.seg_to = np.array([c for pl in plans for c in pl.seg_to], dtype=np.int32).reshape(n_seg, 2)
    p.seg_tgt = np.array([c for pl in plans for c in pl.seg_tgt], dtype=np.int32).reshape(n_seg, 2)
    p.seg_flags = np.array([f for pl in plans for f in pl.seg_first], dtype=np.uint8)
    p.draws = np.array([d for pl in plans for d in pl.draws], dtype=np.uint8)
    p.det_yx = np.array([c for pl in plans for c in pl.det_positions], dtype=np.int32).reshape(n_det, 2)
    p.rows = np.array([e.patch_height for e in envs], dtype=np.int32)
    p.cols = np.array([e.patch_width for e in envs], dtype=np.int32)
    p.boxes, p.n_boxes, p.n_max, _, p.boxes_f64 = boxes_array([e.raw_bboxes for e in envs], want_float=True)
    return p


# ---------------------------------------------------------------------------------------------------
# native planner -> packed
# ---------------------------------------------------------------------------------------------------
class _NativePlan:
    def __init__(self):
        self.handle = ctypes.c_void_p()
        _cabi.check(_cabi.lib().jn_plan_create(ctypes.byref(self.handle)))

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            try:
                _cabi.lib().jn_plan_destroy(h)
            except Exception:
                pass


_native_plan: Optional[_NativePlan] = None
_det_capacity: Dict[tuple, int] = {}  # high-water mark of detection tiles per batch (see expand_packed)


_native_verdict: Optional[bool] = None  # None = not checked yet in this process


def native_planner_verified() -> bool:
    """The native planner restates CPython's ``set`` / tuple hash / ``random`` and numpy's PCG64 streams; a new
    interpreter or numpy release may change any of them.  On first use, a dozen seeded episodes are planned both
    ways (the python planner calls the real objects); on any difference the native planner is switched off for
    this process with a warning -- seeded trajectories then keep following the reference, just more slowly."""
    global _native_verdict
    if _native_verdict is not None:
        return _native_verdict
    import warnings

    from .simple_env import NeedleSimpleEnv
    from ..utils import BBox

    class _Shape:
        def __init__(self, h, w):
            self.shape, self.is_cuda = (3, h, w), True

    import contextlib
    import io

    saved = random.getstate()
    ok = True
    try:
      with contextlib.redirect_stdout(io.StringIO()):  # the python planner prints the reference's warnings
          rng = np.random.default_rng(20240229)
          for binomial in (False, True):
              P = 16
              grids = [(5, 6), (8, 9), (3, 3), (12, 7), (5, 6), (1, 4)]
              bboxes = []
              for gh, gw in grids:
                  raw = []
                  for _ in range(int(rng.integers(0, 4))):
                      bw, bh = (int(v) for v in rng.integers(2, 3 * P, size=2))
                      x1, y1 = int(rng.integers(-4, gw * P)), int(rng.integers(-4, gh * P))
                      raw.append(BBox(Position(y1, x1), Position(y1 + bh, x1 + bw)))
                  bboxes.append(raw)
              seeds = [int(v) for v in rng.integers(0, 2**62, size=len(grids))]
              rows = np.array([g[0] for g in grids], dtype=np.int32)
              cols = np.array([g[1] for g in grids], dtype=np.int32)
              boxes, n_boxes, n_max, _ = boxes_array(bboxes)
              random.seed(99)
              a = plan_native(boxes, n_boxes, n_max, rows, cols, P, seeds, 0, 3, binomial, None)
              state_native = random.getstate()
              random.seed(99)
              envs = [NeedleSimpleEnv(_Shape(gh * P, gw * P), P, bboxes[i], seeds[i]) for i, (gh, gw) in enumerate(grids)]
              b = pack_python_plans(envs, [e.plan_sample(0, 3, binomial, None) for e in envs])
              same = state_native == random.getstate() and all(
                  np.array_equal(getattr(a, k), getattr(b, k))
                  for k in ("start", "seg_begin", "seg_to", "seg_tgt", "seg_flags", "draw_begin", "draws", "det_begin",
                            "det_yx"))
              ok = ok and same
    except Exception as exc:  # a planner that cannot even run is not trusted either
        warnings.warn(f"jolineedle_b200: native planner self-check raised {exc!r}")
        ok = False
    finally:
        random.setstate(saved)
    if not ok:
        warnings.warn("jolineedle_b200: the native planner no longer reproduces python's random / set / numpy "
                      "streams on this interpreter (CPython or numpy changed?); falling back to the python planner")
    _native_verdict = ok
    return ok


def native_supported(rows: np.ndarray, cols: np.ndarray, exact_boxes: bool, seeds) -> bool:
    if not exact_boxes or (len(rows) and (int(rows.max()) > 4096 or int(cols.max()) > 4096)):
        return False
    if seeds is not None and len(seeds):
        try:  # one vectorised look at the common case: a list of non-negative python / numpy ints
            arr = np.asarray(seeds)
        except (OverflowError, ValueError):
            arr = None
        if arr is not None and arr.dtype.kind in "iu" and arr.ndim == 1:
            return bool(arr.min() >= 0)
        for s in seeds:  # mixed lists (None = unseeded episode), huge python ints
            if s is not None and not (isinstance(s, (int, np.integer)) and 0 <= int(s) < (1 << 64)):
                return False
    return True


class _PlanTicket:
    """A native plan in flight (``jn_plan_start``): the input arrays it reads, kept alive until ``finish``."""

    __slots__ = ("arrays", "mt", "version", "gauss", "n", "n_max")


def plan_native_start(boxes: np.ndarray, n_boxes: np.ndarray, n_max: int, rows: np.ndarray, cols: np.ndarray,
                      patch_size: int, seeds: Optional[Sequence[Optional[int]]], min_keypoints: int,
                      max_keypoints: int, binomial_keypoints: bool, position: Optional[Position]) -> _PlanTicket:
    """Start ``jn_plan_run`` on a thread of the library (no GIL involved); :func:`plan_native_finish` joins it.
    Python's global ``random`` state is handed to the C++ side and written back at the end, so
    interleaving with other users of ``random`` behaves as if ``random.choice`` had been called from
    python -- provided nobody touches ``random`` between start and finish."""
    global _native_plan
    if _native_plan is None:
        _native_plan = _NativePlan()
    lib, h = _cabi.lib(), _native_plan.handle
    n = len(rows)
    seed_arr = np.zeros(n, dtype=np.uint64)
    has_seed = np.zeros(n, dtype=np.uint8)
    if seeds is not None:
        if None not in seeds:
            seed_arr, has_seed = np.array(seeds, dtype=np.uint64), np.ones(n, dtype=np.uint8)
        else:
            for i, s in enumerate(seeds):
                if s is not None:
                    seed_arr[i], has_seed[i] = int(s), 1
    start = None
    if position is not None:
        start = np.tile(np.array([int(position[0]), int(position[1])], dtype=np.int32), (n, 1))
    version, mt_words, gauss = random.getstate()
    mt = np.array(mt_words, dtype=np.uint32)
    rc = lib.jn_plan_start(h, n, boxes.ctypes.data, n_boxes.ctypes.data, boxes.shape[1] if n_max > 0 else 0,
                           rows.ctypes.data, cols.ctypes.data, patch_size, seed_arr.ctypes.data, has_seed.ctypes.data,
                           min_keypoints, max_keypoints, 1 if binomial_keypoints else 0,
                           None if start is None else start.ctypes.data, mt.ctypes.data)
    if rc != _cabi.JN_OK:
        raise _cabi.NativeLibraryError(f"native planner could not start (status {rc}): a plan is already running")
    t = _PlanTicket()
    t.arrays = (boxes, n_boxes, rows, cols, seed_arr, has_seed, start)
    t.mt, t.version, t.gauss, t.n, t.n_max = mt, version, gauss, n, n_max
    return t


def plan_native_finish(t: _PlanTicket) -> PackedPlans:
    lib, h = _cabi.lib(), _native_plan.handle
    rc = lib.jn_plan_wait(h)
    if rc != _cabi.JN_OK:
        msg = lib.jn_plan_error(h).decode("utf-8", "replace")
        if rc == _cabi.JN_ERR_INVALID:
            raise AssertionError(msg)  # e.g. start position outside the grid (simple_env.py:73-74)
        raise _cabi.NativeLibraryError(f"native planner failed (status {rc}): {msg}")
    random.setstate((t.version, tuple(t.mt.tolist()), t.gauss))
    n = t.n
    boxes, n_boxes, rows, cols = t.arrays[:4]
    n_seg, n_draw, n_det = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    lib.jn_plan_sizes(h, ctypes.byref(n_seg), ctypes.byref(n_draw), ctypes.byref(n_det))
    p = PackedPlans()
    p.n = n
    p.start = np.empty((n, 2), dtype=np.int32)
    p.seg_begin = np.empty(n + 1, dtype=np.int32)
    p.seg_to = np.empty((n_seg.value, 2), dtype=np.int32)
    p.seg_tgt = np.empty((n_seg.value, 2), dtype=np.int32)
    p.draw_begin = np.empty(n + 1, dtype=np.int32)
    p.det_begin = np.empty(n + 1, dtype=np.int32)
    p.det_yx = np.empty((n_det.value, 2), dtype=np.int32)
    p.seg_flags = np.empty(n_seg.value, dtype=np.uint8)
    p.draws = np.empty(n_draw.value, dtype=np.uint8)
    lib.jn_plan_export(h, p.start.ctypes.data, p.seg_begin.ctypes.data, p.seg_to.ctypes.data, p.seg_tgt.ctypes.data,
                       p.draw_begin.ctypes.data, p.det_begin.ctypes.data, p.det_yx.ctypes.data,
                       p.seg_flags.ctypes.data, p.draws.ctypes.data)
    p.rows, p.cols, p.n_boxes, p.boxes, p.n_max = rows, cols, n_boxes, boxes, t.n_max
    return p


def plan_native(boxes: np.ndarray, n_boxes: np.ndarray, n_max: int, rows: np.ndarray, cols: np.ndarray,
                patch_size: int, seeds: Optional[Sequence[Optional[int]]], min_keypoints: int, max_keypoints: int,
                binomial_keypoints: bool, position: Optional[Position]) -> PackedPlans:
    """Run ``jn_plan_run`` on the host and wait for it."""
    return plan_native_finish(plan_native_start(boxes, n_boxes, n_max, rows, cols, patch_size, seeds, min_keypoints,
                                                max_keypoints, binomial_keypoints, position))


# ---------------------------------------------------------------------------------------------------
# device half
# ---------------------------------------------------------------------------------------------------
def _upload(array: np.ndarray, device) -> torch.Tensor:
    """Host array -> device through torch's pinned-memory cache: a truly asynchronous copy (a
    pageable source would make the copy wait for the stream, serialising host planning of batch
    i+1 behind the gathers of batch i)."""
    staged = torch.empty(array.shape, dtype=torch.from_numpy(array[:0]).dtype, pin_memory=True)
    staged.numpy()[...] = array
    return staged.to(device, non_blocking=True)


def _upload_blocks(arrays: Sequence[np.ndarray], device) -> List[torch.Tensor]:
    """Several host arrays -> device in ONE pinned staging buffer and ONE asynchronous copy (a batch used to
    pay four pinned allocations and four copies); returns one typed, 16-byte aligned view per array."""
    offsets, total = [], 0
    for a in arrays:
        offsets.append(total)
        total += -(-a.nbytes // 16) * 16
    staged = torch.empty(max(total, 16), dtype=torch.uint8, pin_memory=True)
    host = staged.numpy()
    for a, off in zip(arrays, offsets):
        if a.nbytes:
            host[off:off + a.nbytes] = np.ascontiguousarray(a).reshape(-1).view(np.uint8)
    dev = staged.to(device, non_blocking=True)
    out = []
    for a, off in zip(arrays, offsets):
        dtype = torch.from_numpy(a[:0].reshape(-1)).dtype
        out.append(dev[off:off + a.nbytes].view(dtype).view(a.shape))
    return out



def expand_packed(image_set: ImageSet, p: PackedPlans, max_ep_len: int, normalize: bool = False,
                  engine: str = "auto", reuse_glimpses: bool = True, focus: bool = False) -> Dict[str, torch.Tensor]:
    """Packed plans -> the collated sample dict of the reference (keys ``patches``,
    ``current_actions``, ``next_actions``, ``positions``, ``masks``, ``labels``, ``local_bboxes``,
    ``patches_yolox``, ``bboxes_yolox``; plus ``_ep_len`` / ``_status`` for diagnostics)."""
    lib = _cabi.lib()
    n, T, n_max = p.n, int(max_ep_len), p.n_max
    dev, P = image_set.device, image_set.patch_size
    if image_set.out_dtype(normalize) != torch.float32:
        raise ValueError("supervised samples are float32: pass float32 images, or uint8 images with normalize=True")
    n_seg, n_draw, n_det = len(p.seg_flags), len(p.draws), len(p.det_yx)
    det_src = np.repeat(np.arange(n, dtype=np.int32), np.diff(p.det_begin))
    # three uploads: int32 block, uint8 block, int64 block
    i32 = np.concatenate([p.start.ravel(), p.seg_begin, p.seg_to.ravel(), p.seg_tgt.ravel(), p.draw_begin, p.rows,
                          p.cols, p.n_boxes, det_src])
    u8 = np.concatenate([p.seg_flags, p.draws, np.zeros(1, dtype=np.uint8)])
    i64 = np.concatenate([p.boxes.ravel(), p.det_yx.astype(np.int64).ravel()])
    # boxes that are not whole pixels: labels and local boxes from the float64 coordinates (python planner only)
    float_boxes = p.boxes_f64 is not None
    blocks = _upload_blocks([i32, u8, i64] + ([p.boxes_f64] if float_boxes else []), dev)
    d_i32, d_u8, d_i64 = blocks[:3]
    d_boxes_f64 = blocks[3] if float_boxes else None
    patch_bitmaps = lib.jn_patch_bitmaps_f64 if float_boxes else lib.jn_patch_bitmaps
    local_boxes = lib.jn_local_boxes_f64 if float_boxes else lib.jn_local_boxes

    o = 0

    def take(count):
        nonlocal o
        view = d_i32[o:o + count]
        o += count
        return view

    d_start, d_seg_begin = take(2 * n), take(n + 1)
    d_seg_to, d_seg_tgt = take(2 * n_seg), take(2 * n_seg)
    d_draw_begin, d_rows, d_cols, d_nboxes, d_det_src = take(n + 1), take(n), take(n), take(n), take(n_det)
    d_flags, d_draws = d_u8[:n_seg], d_u8[n_seg:n_seg + n_draw + 1]
    d_boxes = d_boxes_f64 if float_boxes else d_i64[:p.boxes.size].view(p.boxes.shape)
    d_det_pos = d_i64[p.boxes.size:].view(n_det, 2)

    words = int(((p.rows.astype(np.int64) * p.cols + 31) // 32).max())
    stream = _cabi.stream_ptr(dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    area = torch.empty((n, words), dtype=torch.int32, device=dev)
    out = {
        "patches": torch.empty((n, T) + image_set.out_shape(1, focus)[1:], dtype=torch.float32, device=dev),
        "current_actions": torch.empty((n, T), dtype=torch.long, device=dev),
        "next_actions": torch.empty((n, T), dtype=torch.long, device=dev),
        "positions": torch.empty((n, T, 2), dtype=torch.long, device=dev),
        "masks": torch.empty((n, T), dtype=torch.float32, device=dev),
        "labels": torch.empty((n, T), dtype=torch.long, device=dev),
        "local_bboxes": torch.empty((n, T, n_max, 6), dtype=torch.float32, device=dev),
    }
    gather_src = torch.empty((n, T), dtype=torch.int32, device=dev)
    ep_len = torch.empty((n,), dtype=torch.int32, device=dev)
    with _cabi.on_device(dev):
        # K0: 5 %-area bitmaps (labels = inside_bbox, simple_env.py:225,478)
        _cabi.check(patch_bitmaps(d_boxes.data_ptr(), d_nboxes.data_ptr(), n, p.boxes.shape[1], P, 0, 0,
                                         d_rows.data_ptr(), d_cols.data_ptr(), _cabi.RULE_AREA5, area.data_ptr(),
                                         words, stream))
        # K3: plan -> per-step records
        _cabi.check(lib.jn_traj_expand(
            d_start.data_ptr(), d_seg_begin.data_ptr(), d_seg_to.data_ptr(), d_seg_tgt.data_ptr(),
            d_flags.data_ptr(), d_draw_begin.data_ptr(), d_draws.data_ptr(), area.data_ptr(), words,
            d_cols.data_ptr(), n, T, out["positions"].data_ptr(), out["current_actions"].data_ptr(),
            out["next_actions"].data_ptr(), out["labels"].data_ptr(), out["masks"].data_ptr(),
            gather_src.data_ptr(), ep_len.data_ptr(), status.data_ptr(), stream))
        # per-step local boxes (simple_env.py:479)
        if n_max > 0:
            _cabi.check(local_boxes(d_boxes.data_ptr(), d_nboxes.data_ptr(), n_max, P,
                                    out["positions"].data_ptr(), gather_src.data_ptr(), n * T,
                                    out["local_bboxes"].data_ptr(), stream))
    # K1: the glimpses themselves, straight into [B, T, C, P, P]; padded slots are zero-filled
    traj_tiles = out["patches"].view((n * T,) + out["patches"].shape[2:])
    reuse_set = None
    if image_set.host_mapped and reuse_glimpses:
        # Host-resident images: a walk with detours revisits patches (~10 % of the slots); those tiles are
        # copied inside HBM from their first occurrence instead of crossing PCIe again.  The trajectory
        # buffer doubles as a set of n*T one-patch images for that (and for the detection patches below).
        reuse_set = ImageSet(traj_tiles, traj_tiles.shape[-1])  # one-patch images (P, or P/2 in the Focus layout)
        first_src = torch.empty((n * T,), dtype=torch.int32, device=dev)
        repeat_src = torch.empty((n * T,), dtype=torch.int32, device=dev)
        with _cabi.on_device(dev):
            _cabi.check(lib.jn_tile_dedupe(out["positions"].data_ptr(), gather_src.data_ptr(), n * T, T,
                                           first_src.data_ptr(), repeat_src.data_ptr(), stream))
        image_set.gather(out["positions"].view(n * T, 2), src_index=first_src, out=traj_tiles, normalize=normalize,
                         focus=focus, engine=engine, status=status, tag="trajectory")
        reuse_set.gather(None, src_index=repeat_src, out=traj_tiles, engine=engine, status=status,
                         tag="trajectory-reuse")
        out["_host_traj_tiles"] = (first_src >= 0).sum()  # trajectory tiles that did cross PCIe
    else:
        image_set.gather(out["positions"].view(n * T, 2), src_index=gather_src.view(n * T), out=traj_tiles,
                         normalize=normalize, focus=focus, engine=engine, status=status, tag="trajectory")
    # detection patches: every box patch + one random empty patch per image (simple_env.py:397-441)
    # (the number of detection patches varies from batch to batch, and a buffer of a new size is a fresh
    # multi-GB cudaMalloc -- tens of milliseconds -- for torch's caching allocator.  So the capacity only
    # ever grows, with 1/8 headroom, per (device, tile shape, batch size): after the first batches every request has the
    # same size and is served from the cache; the result is a view of the buffer's head)
    key = (dev, image_set.channels, P, n, focus)
    need = max(n_det, 1)
    det_cap = _det_capacity.get(key, 0)
    if need > det_cap:
        det_cap = _det_capacity[key] = -(-(need + need // 8) // 64) * 64
    det_buf = torch.empty(image_set.out_shape(det_cap, focus), dtype=torch.float32, device=dev)
    if reuse_set is not None and n_det > 0:
        # Host-resident images: most detection patches were just gathered as trajectory glimpses, so take
        # those from the [n*T, C, P, P] buffer in HBM instead of pulling them over PCIe a second time: two
        # passes over the same output, each leaving the other's items untouched (skip_negative) -- gather of
        # the host tiles (normalising when the images are uint8), plain copy of the reused ones.
        pos2 = torch.empty((n_det, 2), dtype=torch.long, device=dev)
        src2 = torch.empty((n_det,), dtype=torch.int32, device=dev)
        with _cabi.on_device(dev):
            _cabi.check(lib.jn_tile_lookup(out["positions"].data_ptr(), gather_src.data_ptr(), T,
                                           d_det_pos.data_ptr(), d_det_src.data_ptr(), n_det, image_set.n_images,
                                           pos2.data_ptr(), src2.data_ptr(), stream))
        minus_one, n_img = torch.full_like(src2, -1), image_set.n_images
        image_set.gather(pos2, src_index=torch.where(src2 < n_img, src2, minus_one), out=det_buf[:n_det],
                         normalize=normalize, focus=focus, engine=engine, status=status, tag="detection",
                         skip_negative=True)
        reuse_set.gather(pos2, src_index=torch.where(src2 >= n_img, src2 - n_img, minus_one), out=det_buf[:n_det],
                         engine=engine, status=status, tag="detection-reuse", skip_negative=True)
        out["patches_yolox"] = det_buf[:n_det]
        out["_host_det_tiles"] = (src2 < image_set.n_images).sum()  # detection tiles that did cross PCIe
    else:
        out["patches_yolox"] = image_set.gather(d_det_pos, src_index=d_det_src, out=det_buf[:n_det],
                                                normalize=normalize, focus=focus, engine=engine, status=status,
                                                tag="detection")
    det_boxes = torch.empty((n_det, n_max, 6), dtype=torch.float32, device=dev)
    if n_max > 0 and n_det > 0:
        with _cabi.on_device(dev):
            _cabi.check(local_boxes(d_boxes.data_ptr(), d_nboxes.data_ptr(), n_max, P, d_det_pos.data_ptr(),
                                    d_det_src.data_ptr(), n_det, det_boxes.data_ptr(), stream))
    out["bboxes_yolox"] = det_boxes
    out["_ep_len"] = ep_len
    out["_status"] = status
    return out


# ---------------------------------------------------------------------------------------------------
# batched entry point
# ---------------------------------------------------------------------------------------------------
def plan_batch(bboxes: Sequence[Sequence], heights: Sequence[int], widths: Sequence[int], patch_size: int,
               min_keypoints: int, max_keypoints: int, binomial_keypoints: bool = False,
               position: Optional[Position] = None, seeds: Optional[Sequence[Optional[int]]] = None,
               planner: str = "auto", deferred: bool = False):
    """Host half for a batch of images given only their sizes and boxes (no pixel is touched).
    ``deferred``: return a zero-argument callable that yields the plans -- with the native planner the work
    then runs on a library thread while the caller goes on (``generate_trajectories`` builds the image set
    meanwhile)."""
    for h, w in zip(heights, widths):
        # same precondition as get_patch (simple_env.py:68-69)
        assert h % patch_size == 0 and w % patch_size == 0, f"image {h}x{w} is not a multiple of {patch_size}"
    rows = np.array([h // patch_size for h in heights], dtype=np.int32)
    cols = np.array([w // patch_size for w in widths], dtype=np.int32)
    boxes, n_boxes, n_max, exact = boxes_array(bboxes)
    if planner not in ("auto", "native", "python"):
        raise ValueError(f"unknown planner {planner!r}")
    use_native = planner != "python" and native_supported(rows, cols, exact, seeds)
    if planner == "native" and not use_native:
        raise ValueError("the native planner needs integer boxes, grids up to 4096x4096 and seeds below 2**64")
    if use_native and not native_planner_verified():
        if planner == "native":
            raise _cabi.NativeLibraryError("the native planner failed its self-check against the python planner")
        use_native = False
    if use_native:
        ticket = plan_native_start(boxes, n_boxes, n_max, rows, cols, patch_size, seeds, min_keypoints, max_keypoints,
                                   binomial_keypoints, position)
        return (lambda: plan_native_finish(ticket)) if deferred else plan_native_finish(ticket)
    from .simple_env import NeedleSimpleEnv

    class _Shape:  # plan_sample only needs the image's shape
        def __init__(self, h, w):
            self.shape = (3, h, w)
            self.is_cuda = True

    envs = [NeedleSimpleEnv(_Shape(heights[i], widths[i]), patch_size, bboxes[i], None if seeds is None else seeds[i])
            for i in range(len(bboxes))]
    plans = [e.plan_sample(min_keypoints, max_keypoints, binomial_keypoints, position) for e in envs]
    packed = pack_python_plans(envs, plans)
    return (lambda: packed) if deferred else packed


def generate_trajectories(
    batch: Dict,
    patch_size: int,
    max_seq_len: int,
    min_keypoints: int,
    max_keypoints: int,
    binomial_keypoints: bool = False,
    position: Optional[Position] = None,
    seeds: Optional[Sequence[Optional[int]]] = None,
    normalize: bool = False,
    device=None,
    engine: str = "auto",
    planner: str = "auto",
    zero_copy: bool = True,
    stats: Optional[dict] = None,
    check: bool = False,
    focus: bool = False,
) -> Dict[str, torch.Tensor]:
    """Batched supervised trajectories (``SupervisedTrainer.generate_trajectories``,
    supervised.py:95-136): ``batch`` holds lists ``image`` ([C,H,W] tensors), ``bboxes`` (lists
    of ``BBox``) and ``class_id``.  Returns the collated dict of the reference on the GPU.
    ``image`` may also be ONE stacked ``[B, C, H, W]`` tensor when the images share a size.
    ``seeds`` (one per image) makes the plans reproducible; the reference builds unseeded envs.
    CPU images are uploaded to ``device`` first (there is no CPU path).  ``check`` synchronises and raises if
    a kernel flagged its input (a position outside the grid, a plan the expansion could not follow); without it
    the flags travel in ``stats["status"]`` and nothing waits for the device.  ``focus``: ``patches`` and
    ``patches_yolox`` come out in the YOLOX Focus space-to-depth layout ``[4C, P/2, P/2]`` (what the detector's
    stem computes first), written that way by the gather itself."""
    stacked = batch["image"] if isinstance(batch["image"], torch.Tensor) else None
    if stacked is not None:
        # one [B, C, H, W] tensor (images of one size, e.g. what `pinned_u8_collate` makes of a LARD batch): a
        # single slab -- no per-image bookkeeping on the host, tensor tiles instead of per-row copies
        if stacked.dim() != 4:
            raise ValueError("a stacked image batch must be [B, C, H, W]")
        if device is not None and not (stacked.is_cuda or (zero_copy and stacked.is_pinned())):
            stacked = stacked.to(device, non_blocking=True)
        images = stacked
        heights, widths = [stacked.shape[2]] * stacked.shape[0], [stacked.shape[3]] * stacked.shape[0]
    else:
        images: List[torch.Tensor] = list(batch["image"])
        if device is not None:
            # Pinned host images are NOT uploaded: a supervised episode looks at ~10 of an image's 30
            # patches once, so the gather reads just those tiles over PCIe (zero-copy).  Pageable host
            # images have to be staged through a full upload.
            images = [im if (im.is_cuda or (zero_copy and im.is_pinned())) else im.to(device, non_blocking=True)
                      for im in images]
        heights, widths = [im.shape[1] for im in images], [im.shape[2] for im in images]
    plans = plan_batch(batch["bboxes"], heights, widths, patch_size,
                       min_keypoints, max_keypoints, binomial_keypoints, position, seeds, planner, deferred=True)
    try:
        image_set = ImageSet(images, patch_size, device=device)  # while the native planner runs on its own thread
    finally:
        packed = plans()  # always joined: the planner reads arrays that die with this frame
    out = expand_packed(image_set, packed, max_seq_len, normalize, engine, focus=focus)
    class_id = np.array([int(c) for c in batch["class_id"]], dtype=np.int64)
    out["class_id"] = _upload(class_id, image_set.device)
    diagnostics = {k: out.pop(k) for k in [k for k in out if k.startswith("_")]}
    if check:
        flags = int(diagnostics["_status"].item())
        if flags & _cabi.STATUS_BAD_PLAN:
            raise RuntimeError("trajectory expansion rejected a plan (fewer replacement moves than the walk needs)")
        if flags & _cabi.STATUS_BAD_POSITION:
            raise IndexError("a planned position lies outside its image's patch grid")
    if stats is not None:  # device scalars / tensors for benchmarks: untruncated lengths, tiles read from the host
        stats.update({k[1:]: v for k, v in diagnostics.items()})
    return out
