"""Image sets and the glimpse gather (K1) -- thin Python handles over the C ABI.

An :class:`ImageSet` borrows one ``[B, C, H, W]`` CUDA tensor, or a list of ``[C, H, W]`` /
``[b, C, H, W]`` CUDA tensors of possibly different sizes, and serves ``[C, P, P]`` tiles out
of them.  No pixel is copied at construction; the tensors are kept alive by the handle.
"""
import ctypes
from operator import attrgetter
from typing import Dict, List, Optional, Sequence, Union

import numpy as np
import torch

from . import _cabi

class LaunchTimer:
    """Per-launch timing of the gathers (bench.py's roofline leg): CUDA events recorded on the launch stream
    right around each gather call, taken from a pool created up front so that the timed region does not pay
    for event construction.  ``records`` holds ``(tag, n_items, start_event, end_event)``."""

    def __init__(self, capacity: int = 0):
        self.pool = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                     for _ in range(capacity)]
        self.records: list = []

    def begin(self, device=None):
        k = len(self.records)
        pair = self.pool[k] if k < len(self.pool) else (torch.cuda.Event(enable_timing=True),
                                                        torch.cuda.Event(enable_timing=True))
        # (Event.record() without a stream looks the current stream up through several python layers, ~13 us;
        # the cached Stream object of the raw handle costs a dict lookup)
        stream = None if device is None else _stream_object(device, _cabi.stream_ptr(device))
        pair[0].record(stream) if stream is not None else pair[0].record()
        return pair, stream

    def end(self, token, tag: str, n_items: int):
        pair, stream = token
        pair[1].record(stream) if stream is not None else pair[1].record()
        self.records.append((tag, n_items, pair[0], pair[1]))


# When a LaunchTimer is installed here, every gather is bracketed by a pair of its events.
TIMING: Optional[LaunchTimer] = None


_DTYPE_OF = attrgetter("dtype")
_STREAMS: Dict[tuple, "torch.cuda.Stream"] = {}


def _stream_object(device, raw: int):
    """torch Stream object of the current raw stream (building one costs ~25 us: cached per handle)."""
    key = (device.index, raw)
    stream = _STREAMS.get(key)
    if stream is None:
        stream = _STREAMS[key] = torch.cuda.current_stream(device)
    return stream


class ImageSet:
    def __init__(self, images: Union[torch.Tensor, Sequence[torch.Tensor]], patch_size: int, device=None,
                 pad_to_patch: bool = False):
        """``device`` is only needed for *pinned host* images: those are not uploaded -- the
        gather kernels read the tiles they need straight out of the page-locked host memory
        (same pointer under unified addressing), so only glimpsed pixels ever cross PCIe.

        ``pad_to_patch``: image sizes need not be multiples of ``patch_size``; the set behaves as if
        every image had been zero-padded at the bottom / right (``complete_to_patch_size`` /
        ``padded_collate_fn``, dataset.py:307-347,379-406) without materialising the padding."""
        slabs: List[torch.Tensor] = [images] if isinstance(images, torch.Tensor) else list(images)
        if not slabs:
            raise ValueError("empty image set")
        # This runs once per batch with hundreds of images: one comprehension per property (a python-level
        # loop with seven attribute reads and five appends per tensor cost 0.8 ms per 256 images).
        first = slabs[0]
        dtype, channels = first.dtype, first.shape[-3]
        shapes = [t.shape for t in slabs]
        if not set(map(len, shapes)) <= {3, 4}:
            bad = next(sh for sh in shapes if len(sh) not in (3, 4))
            raise ValueError(f"images must be [C,H,W] or [B,C,H,W], got shape {tuple(bad)}")
        where = [t.get_device() for t in slabs]  # CUDA device index, -1 for host tensors
        self.host_mapped = min(where) < 0
        if self.host_mapped:
            for t, d in zip(slabs, where):
                if d < 0 and (device is None or not t.is_pinned()):
                    _cabi.require_cuda(t, "images")
        devices = {d for d in where if d >= 0}
        if len(devices) > 1 or set(map(_DTYPE_OF, slabs)) != {dtype} or {sh[-3] for sh in shapes} != {channels}:
            raise ValueError("all images of a set must share device, dtype and channel count")
        cuda_device = torch.device("cuda", next(iter(devices))) if devices else None
        for t, d in zip(slabs, where):
            if d < 0 and not t.is_contiguous():
                # .contiguous() of a pinned tensor is an ordinary pageable copy: its pointer means nothing to the GPU
                raise ValueError("host-mapped (pinned) images must be contiguous; pin the contiguous copy instead")
        norm = [t if t.is_contiguous() else t.contiguous() for t in slabs]
        counts = [sh[0] if len(sh) == 4 else 1 for sh in shapes]
        heights, widths = [sh[-2] for sh in shapes], [sh[-1] for sh in shapes]
        ptrs = [t.data_ptr() for t in norm]
        self.device = torch.device(device) if (self.host_mapped or cuda_device is None) else cuda_device
        if self.device.type != "cuda":
            raise _cabi.NativeLibraryError(f"image sets live on a CUDA device, got {self.device}")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        if cuda_device is not None and cuda_device != self.device:
            raise ValueError(f"images live on {cuda_device} but the set was asked for {self.device}")
        self.dtype, self.channels = dtype, channels
        self.patch_size = int(patch_size)
        self.pad_to_patch = bool(pad_to_patch)
        self._slabs = norm  # keeps the memory alive
        self.counts, self.heights, self.widths = counts, heights, widths
        self.n_images = sum(counts)
        n = len(norm)
        # per-image table of multi-slab sets: torch-owned scratch (stream-ordered), not cudaMalloc
        # (written by the library into pinned host scratch, uploaded here with torch so that the pinned
        # block is not recycled before the copy has run; no cudaMalloc / cudaFree / sync per batch)
        self._table = self._table_host = None
        if n > 1:
            self._table_host = torch.empty(32 * self.n_images, dtype=torch.uint8, pin_memory=True)
            self._table = torch.empty(32 * self.n_images, dtype=torch.uint8, device=self.device)
        handle = ctypes.c_void_p()
        a_ptrs = np.array(ptrs, dtype=np.uint64)
        a_dims = np.array([counts, heights, widths], dtype=np.int32)  # rows are contiguous int32 arrays
        with _cabi.on_device(self.device):
            create = _cabi.lib().jn_images_create_padded if pad_to_patch else _cabi.lib().jn_images_create
            rc = create(
                ctypes.byref(handle), n, a_ptrs.ctypes.data, a_dims[0].ctypes.data, a_dims[1].ctypes.data,
                a_dims[2].ctypes.data, channels,
                _cabi.dtype_code(dtype), self.patch_size, _cabi.ptr(self._table_host), _cabi.ptr(self._table),
                _cabi.stream_ptr(self.device),
            )
            self._table_ready, self._table_streams = None, set()
            if rc == _cabi.JN_OK and self._table is not None:
                self._table.copy_(self._table_host, non_blocking=True)
                # gathers on OTHER streams must not run before this copy: they wait for the event once each
                raw = _cabi.stream_ptr(self.device)
                self._table_ready = torch.cuda.Event()
                self._table_ready.record(_stream_object(self.device, raw))
                self._table_streams.add(raw)
        # same precondition as the reference envs: sizes must be multiples of the patch size
        _cabi.check(rc, invalid_exc=AssertionError)
        self._handle = handle

    def __del__(self):
        h, self._handle = getattr(self, "_handle", None), None
        if h:
            try:
                _cabi.lib().jn_images_destroy(h)
            except Exception:  # interpreter shutdown
                pass

    def _order_after_table(self):
        """Multi-slab sets: the per-image table was uploaded on the stream that was current at construction; the
        first gather on any other stream waits for that upload."""
        if self._table_ready is not None:
            sp = _cabi.stream_ptr(self.device)
            if sp not in self._table_streams:
                torch.cuda.current_stream(self.device).wait_event(self._table_ready)
                self._table_streams.add(sp)

    def engine_available(self, engine: str) -> bool:
        return bool(_cabi.lib().jn_images_tma_ok(self._handle, _cabi.ENGINES[engine]))

    def out_shape(self, n_items: int, focus: bool):
        p, c = self.patch_size, self.channels
        return (n_items, 4 * c, p // 2, p // 2) if focus else (n_items, c, p, p)

    def out_dtype(self, normalize: bool) -> torch.dtype:
        return torch.float32 if normalize else self.dtype

    def bind(self, normalize: bool = False, focus: bool = False, engine: str = "auto",
             status: Optional[torch.Tensor] = None, tag: str = "", shifts: Optional[torch.Tensor] = None,
             shifts_aligned: bool = False):
        """``launch(positions, out)`` for a fixed gather configuration, with the argument checks of
        :meth:`gather` done once: the batched env calls it every step with its own, known-good tensors
        (``positions`` a contiguous int64 ``[n, 2]`` on this device, ``out`` with contiguous items of the
        right shape and dtype, item i of image i)."""
        if shifts is not None and (shifts.dtype != torch.int32 or tuple(shifts.shape) != (self.n_images, 2)
                                   or shifts.device != self.device or not shifts.is_contiguous()):
            raise ValueError(f"shifts must be a contiguous int32 [{self.n_images}, 2] tensor of (ty, tx) on {self.device}")
        flags = (_cabi.GATHER_NORMALIZE if normalize else 0) | (_cabi.GATHER_FOCUS if focus else 0)
        if shifts is not None and shifts_aligned:
            flags |= _cabi.GATHER_SHIFT_ALIGNED
        lib, handle, code = _cabi.lib(), self._handle, _cabi.ENGINES[engine]
        p_status, p_shifts, device = _cabi.ptr(status), _cabi.ptr(shifts), self.device
        keep = (status, shifts, self)  # the raw pointers above stay valid as long as the closure lives

        def launch(positions: torch.Tensor, out: torch.Tensor):
            self._order_after_table()
            n = positions.shape[0]
            stride = (out.stride(0) if n > 1 else out[0].numel()) * out.element_size()
            timing = TIMING
            with _cabi.on_device(device):
                pair = timing.begin(device) if timing is not None else None
                rc = lib.jn_gather(handle, positions.data_ptr(), None, p_shifts, n, out.data_ptr(), stride, flags, code,
                                   p_status, _cabi.stream_ptr(device))
                if pair is not None:
                    timing.end(pair, tag, n)
            if rc:
                _cabi.check(rc)
            return out

        launch.keep = keep
        return launch

    def gather(
        self,
        positions: Optional[torch.Tensor],
        src_index: Optional[torch.Tensor] = None,
        out: Optional[torch.Tensor] = None,
        normalize: bool = False,
        focus: bool = False,
        engine: str = "auto",
        status: Optional[torch.Tensor] = None,
        tag: str = "",
        shifts: Optional[torch.Tensor] = None,
        shifts_aligned: bool = False,
        skip_negative: bool = False,
    ) -> torch.Tensor:
        """Tile ``positions[i] = (y, x)`` of image ``src_index[i]`` (default ``i``; negative =
        zeros) -> ``out[i]``.  ``out`` may be any tensor whose ``out[i]`` is contiguous (e.g. a
        time slot ``history[:, t]`` of a ``[B, T, C, P, P]`` buffer)."""
        if positions is None:  # patch (0, 0) of every item's image: sets of one-patch images
            if src_index is None:
                raise ValueError("positions=None needs src_index (one one-patch image per item)")
            n = src_index.numel()
        else:
            _cabi.require_cuda(positions, "positions")
            if positions.dtype != torch.int64 or positions.dim() != 2 or positions.shape[1] != 2:
                raise ValueError("positions must be a LongTensor of shape [n, 2]")
            positions = positions.contiguous()
            n = positions.shape[0]
        if src_index is not None:
            if src_index.dtype != torch.int32 or src_index.numel() != n:
                raise ValueError("src_index must be an int32 tensor with one entry per position")
            src_index = src_index.contiguous()
        if shifts is not None:
            # per-image integer translation (ty, tx) with zero fill, see jn_gather in the header
            if shifts.dtype != torch.int32 or tuple(shifts.shape) != (self.n_images, 2) or shifts.device != self.device:
                raise ValueError(f"shifts must be an int32 [{self.n_images}, 2] tensor of (ty, tx) on {self.device}")
            shifts = shifts.contiguous()
        shape, dtype = self.out_shape(n, focus), self.out_dtype(normalize)
        if out is None:
            out = torch.empty(shape, dtype=dtype, device=self.device)
        else:
            if tuple(out.shape) != shape or out.dtype != dtype or out.device != self.device:
                raise ValueError(f"out must be {shape} {dtype} on {self.device}, got {tuple(out.shape)} {out.dtype}")
            if n > 0 and not out[0].is_contiguous():
                raise ValueError("out[i] must be contiguous")
        flags = (_cabi.GATHER_NORMALIZE if normalize else 0) | (_cabi.GATHER_FOCUS if focus else 0)
        if shifts is not None and shifts_aligned:  # every x shift is a multiple of 16 bytes: TMA may serve it
            flags |= _cabi.GATHER_SHIFT_ALIGNED
        if skip_negative:  # negative src_index: leave out[i] as it is (default: zero-fill it)
            flags |= _cabi.GATHER_SKIP_NEGATIVE
        stride = out.stride(0) * out.element_size() if n > 1 else out[0].numel() * out.element_size() if n else 0
        self._order_after_table()
        timing = TIMING
        with _cabi.on_device(self.device):
            pair = timing.begin(self.device) if timing is not None else None
            rc = _cabi.lib().jn_gather(
                self._handle, _cabi.ptr(positions), _cabi.ptr(src_index), _cabi.ptr(shifts), n, out.data_ptr(), stride,
                flags,
                _cabi.ENGINES[engine], _cabi.ptr(status), _cabi.stream_ptr(self.device),
            )
            if pair is not None:
                timing.end(pair, tag, n)
        _cabi.check(rc)
        return out
