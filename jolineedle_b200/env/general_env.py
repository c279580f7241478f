"""Batched RL gaze environment on B200 -- drop-in for the reference's ``NeedleGeneralEnv``
(``src/env/general_env.py:14-573``): same constructor, ``reset`` / ``step`` signatures, return
values, attributes and quirks; the work is done by the CUDA kernels behind
``include/jolineedle_b200.h`` (K0 overlap bitmaps, K1 glimpse gather, K2 warp-per-episode step).

State layout in HBM (per episode): position int64[2]; visited / bbox bitmaps as uint32 words
(bit = y*cols + x); steps int64; has_stopped u8.  The ``[B, rows, cols]`` bool tensors the
reference exposes (``bbox_masks``, ``visited_patches``) are materialised on demand.

Extensions (keyword-only, default = reference behaviour):
  normalize  uint8 images -> float32 crops equal to ``ToTensor`` then crop (dataset.py:240)
  focus      crops come out in the YOLOX Focus layout ``[4C, P/2, P/2]``
  history    pre-allocate ``[B, max_ep_len + 1, C, P, P]`` and write step t's crops into slot
             t + 1 (``patch_history(t)`` then replaces the caller's per-step ``torch.concat``,
             reinforce.py:175-179)
  translate  ``[B, 2]`` integer ``(tx, ty)`` per image: crops come from the image shifted by that many
             pixels with zero fill -- the ``--augment-translate`` augmentation (dataset.py:157-226,
             ``F.affine(translate=[tx, ty], fill=0)``) folded into the gather coordinates instead of a
             shifted copy of the image made on the CPU.  The caller shifts the boxes (as the dataset does).
  (``n_glimps_levels > 1``: the glimpse pyramid of float32 images, bit-identical to the reference's CPU
  torchvision pad + antialiased resize -- ``jolineedle_b200/pyramid.py``)
  device     upload CPU inputs to this CUDA device (the reference keeps everything on
             ``images.device``; there is no CPU path here)
  zero_copy  pinned HOST images are not uploaded: every step reads the glimpsed tiles in place over PCIe,
             and with ``history`` a patch an episode has already seen is served from the crop history in HBM
  pad_to_patch  image sizes that are not multiples of the patch: the env behaves as if the images had been
             zero-padded at the bottom / right (``padded_collate_fn``, dataset.py:307-347) without
             materialising the padding (the TMA unit zero-fills what lies outside the image)
"""
import ctypes
from typing import List, Optional, Tuple

import torch
from torch import Tensor

from .. import _cabi
from .. import gather as _gather_mod
from ..gather import ImageSet
from .common import Action, ACTION_DELTAS, DELTA_TABLE  # noqa: F401  (re-exported like the reference module)


_GYM = None  # gymnasium module, False once an import has failed (a failed import walks sys.path every time)


def _spaces(batch_size: int, channels: int, patch: int, rows: int, cols: int):
    """``observation_space`` / ``action_space`` as the reference declares them (general_env.py:61-72): gymnasium
    objects when that package is installed, plain descriptions with the same fields otherwise (no caller of the
    reference reads them; the env does not depend on gymnasium)."""
    global _GYM
    if _GYM is None:
        try:
            import gymnasium

            _GYM = gymnasium
        except Exception:
            _GYM = False
    shape = (batch_size, channels, patch, patch)
    if _GYM:
        return (_GYM.spaces.Box(low=0, high=1, shape=shape),
                _GYM.spaces.Tuple((_GYM.spaces.Discrete(rows), _GYM.spaces.Discrete(cols))))
    from types import SimpleNamespace

    return (SimpleNamespace(low=0, high=1, shape=shape),
            SimpleNamespace(spaces=(SimpleNamespace(n=rows), SimpleNamespace(n=cols))))


class NeedleGeneralEnv:
    def __init__(
        self,
        images: Tensor,
        bboxes: Tensor,
        patch_size: int,
        max_ep_len: int,
        n_glimps_levels: int,
        stop_enabled: bool = False,
        *,
        normalize: bool = False,
        focus: bool = False,
        history: bool = False,
        engine: str = "auto",
        device=None,
        translate: Optional[Tensor] = None,
        zero_copy: bool = False,
        pad_to_patch: bool = False,
    ):
        # same preconditions as general_env.py:37-39,50-51
        assert images.shape[0] == bboxes.shape[0]
        assert len(images.shape) == 4
        assert n_glimps_levels > 0
        if n_glimps_levels != 1 and translate is not None:
            raise NotImplementedError("the glimpse pyramid (n_glimps_levels > 1) does not combine with translate")
        # zero_copy: pinned HOST images are not uploaded -- every step's gather reads the glimpsed tiles in place
        # over PCIe, and with history=True a patch an episode has already seen is copied from the crop history in
        # HBM instead (jn_visit_sources), so each patch crosses PCIe at most once per episode
        self._zero_copy = bool(zero_copy and device is not None and not images.is_cuda and images.is_pinned())
        if zero_copy and not self._zero_copy and not images.is_cuda:
            raise ValueError("zero_copy needs pinned host images and device=")
        if self._zero_copy and n_glimps_levels != 1:
            raise NotImplementedError("zero_copy serves one glimpse level")
        if device is not None and not images.is_cuda and not self._zero_copy:
            images = images.to(device, non_blocking=True)
        if not self._zero_copy:
            _cabi.require_cuda(images, "images")
        self.patch_size = patch_size
        self.max_ep_len = max_ep_len
        self.n_glimps_levels = n_glimps_levels
        self.stop_enabled = stop_enabled
        self.batch_size, self.n_channels, self.height, self.width = images.shape
        # pad_to_patch: sizes that are not multiples of the patch behave as if the images had been zero-padded at
        # the bottom / right (padded_collate_fn, dataset.py:307-347) -- the gather zero-fills what lies outside
        if not pad_to_patch:
            assert self.height % self.patch_size == 0
            assert self.width % self.patch_size == 0
        elif n_glimps_levels != 1:
            raise NotImplementedError("pad_to_patch serves one glimpse level")
        self.n_vertical_patches = -(-self.height // self.patch_size)
        self.n_horizontal_patches = -(-self.width // self.patch_size)
        self.device = torch.device(device) if self._zero_copy else images.device
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self._normalize, self._focus, self._engine = normalize, focus, engine
        self.observation_space, self.action_space = _spaces(self.batch_size, self.n_channels, patch_size,
                                                            self.n_vertical_patches, self.n_horizontal_patches)
        self._shifts, self._shifts_aligned = None, False
        if translate is not None:
            t = torch.as_tensor(translate)
            assert tuple(t.shape) == (self.batch_size, 2), "translate must be [batch_size, 2] = (tx, ty) per image"
            # kernels take (ty, tx); x shifts that are all multiples of 16 bytes can ride the TMA path
            elem = 1 if images.dtype == torch.uint8 else 4
            self._shifts_aligned = bool(((t[:, 0].to(torch.int64) * elem) % 16 == 0).all()) if not t.is_cuda else False
            self._shifts = t.to(torch.int32).flip(1).contiguous().to(self.device)

        if n_glimps_levels == 1:
            self._set = ImageSet(images, patch_size, device=self.device if self._zero_copy else None,
                                 pad_to_patch=pad_to_patch)
            # [B, G=1, C, H, W] view of the caller's tensor (callers read env.images[0, 0]); the reflect-pad +
            # resize the reference computes and throws away at one level (general_env.py:95-111) is not done
            self.images = self._set._slabs[0].unsqueeze(1)
        else:
            self.images = self.init_glimps_images(images)  # [B, G, C, H, W]
            self._set = ImageSet(self.images.view((-1,) + tuple(self.images.shape[2:])), patch_size)
        self.bboxes = bboxes
        self._boxes_dev = bboxes.to(device=self.device, dtype=torch.int64).contiguous()
        self._words = (self.n_vertical_patches * self.n_horizontal_patches + 31) // 32
        self._status = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._cost = float(torch.tensor(-1 / self.max_ep_len, dtype=torch.float32))  # fl32(-1/T), general_env.py:340
        self._lib = _cabi.lib()

        # K0: patch x bbox containment (any-pixel rule), general_env.py:75,360-379
        self._bbox_words = torch.empty((self.batch_size, self._words), dtype=torch.int32, device=self.device)
        n_boxes = self._boxes_dev.shape[1]
        with _cabi.on_device(self.device):
            _cabi.check(self._lib.jn_patch_bitmaps(
                self._boxes_dev.data_ptr(), None, self.batch_size, n_boxes, patch_size, self.n_vertical_patches,
                self.n_horizontal_patches, None, None, _cabi.RULE_ANY_PIXEL, self._bbox_words.data_ptr(),
                self._words, self._stream()))

        self._history: Optional[Tensor] = None
        if history:
            # the glimpse axis doubles as the time axis of the history, as in the trainer's concat (reinforce.py:176)
            self._history = torch.empty(
                (self.batch_size, (max_ep_len + 1) * n_glimps_levels) + self._set.out_shape(1, focus)[1:],
                dtype=self._set.out_dtype(normalize), device=self.device)
        self._history_set = None
        if self._zero_copy and self._history is not None:
            slots = self._history.shape[1]
            self._history_set = ImageSet(self._history.view((self.batch_size * slots,) + tuple(self._history.shape[2:])),
                                         self._history.shape[-1])  # one-patch images (P, or P/2 in the Focus layout)
        self.host_tiles = torch.zeros((), dtype=torch.long, device=self.device)  # tiles read over PCIe so far
        if n_glimps_levels > 1:
            levels = n_glimps_levels
            ids = torch.arange(self.batch_size * levels, dtype=torch.int32, device=self.device)
            self._level_src = [ids[l::levels].contiguous() for l in range(levels)]  # image b*G + l of level l
        self._t = 0
        self._launch_gather = None  # bound K1 launch of the stand-alone one-level gather (see _gather)
        self._tile_shape = self._set.out_shape(self.batch_size, focus)
        self._tile_dtype = self._set.out_dtype(normalize)
        elem = 4 if self._tile_dtype == torch.float32 else 1
        tile_bytes = elem * self._tile_shape[1] * self._tile_shape[2] * self._tile_shape[3]
        # one native call per step (jn_env_step_gather): argument block filled once, per-step fields rewritten
        self._fused = n_glimps_levels == 1
        self._args = _cabi.EnvStepArgs()
        self._args_ref = ctypes.byref(self._args)
        a = self._args
        a.n, a.rows, a.cols = self.batch_size, self.n_vertical_patches, self.n_horizontal_patches
        a.max_ep_len, a.stop_enabled, a.cost = max_ep_len, 1 if stop_enabled else 0, self._cost
        a.bbox, a.status = self._bbox_words.data_ptr(), self._status.data_ptr()
        a.slots = self._history.shape[1] if self._history is not None else 1
        a.shifts = _cabi.ptr(self._shifts)
        a.flags = ((_cabi.GATHER_NORMALIZE if normalize else 0) | (_cabi.GATHER_FOCUS if focus else 0)
                   | (_cabi.GATHER_SHIFT_ALIGNED if (self._shifts is not None and self._shifts_aligned) else 0))
        a.engine = _cabi.ENGINES[engine]
        a.out_item_stride_bytes = tile_bytes * (a.slots if self._history is not None else 1)
        a.host_tiles = self.host_tiles.data_ptr() if self._history_set is not None else None
        self._set_handle = self._set._handle if self._fused else None
        self._hist_handle = self._history_set._handle if self._history_set is not None else None
        # [B, 1, C, P, P] views of the history slots, made on first use
        self._slot_views: List[Optional[Tensor]] = [None] * (a.slots if self._history is not None else 0)
        self.init_env_variables()

    @torch.no_grad()
    def init_glimps_images(self, images: Tensor) -> Tensor:
        """Stack of progressively zoomed-out copies of the images (general_env.py:84-115): level 0 is the input,
        level k+1 = level k reflect-padded by one patch on every side and resized (antialiased bilinear) back to
        H x W.  Built once per env by ``jn_resize_aa_reflect`` (jolineedle_b200/pyramid.py): the reference's
        torchvision CPU arithmetic restated tap by tap, so the levels -- and the crops the K1 gather then takes
        from them -- equal the reference's bit for bit."""
        from ..pyramid import build_levels

        return build_levels(images, self.patch_size, self.n_glimps_levels)

    # ------------------------------------------------------------------------------------
    def _stream(self):
        return _cabi.stream_ptr(self.device)

    def _unpack(self, words: Tensor) -> Tensor:
        out = torch.empty((self.batch_size, self.n_vertical_patches, self.n_horizontal_patches), dtype=torch.bool,
                          device=self.device)
        with _cabi.on_device(self.device):
            _cabi.check(self._lib.jn_bitmap_unpack(words.data_ptr(), self.batch_size, self.n_vertical_patches,
                                                   self.n_horizontal_patches, out.data_ptr(), self._stream()))
        return out

    @property
    def bbox_masks(self) -> Tensor:
        """``[B, rows, cols]`` bool: patch holds at least one box pixel (general_env.py:360-379)."""
        return self._unpack(self._bbox_words)

    @property
    def visited_patches(self) -> Tensor:
        return self._unpack(self._visited_words)

    def init_env_variables(self, zero: bool = True):  # general_env.py:117-142
        """Fresh episode state.  Per-step results live in step-major rings allocated here -- row t of
        ``[max_ep_len, B]`` rewards / terminated / truncated and of ``[max_ep_len + 1, B, 2]`` positions is what
        step t returns -- so a step allocates nothing, and tensors handed out during one episode are never
        overwritten by the next (``reset`` takes new rings).  ``zero=False``: the reset kernel is about to
        overwrite the state, skip the memsets."""
        b, dev, words = self.batch_size, self.device, self._words
        make = torch.zeros if zero else torch.empty
        self._visited_words = make((b, words), dtype=torch.int32, device=dev)
        self.steps = make((b,), dtype=torch.long, device=dev)
        self.has_stopped = make((b,), dtype=torch.bool, device=dev)
        self._t = 0
        if self._history_set is not None:
            self._first_slot = torch.empty((b, self.n_vertical_patches * self.n_horizontal_patches),
                                           dtype=torch.int32, device=dev)
            if zero:
                self._first_slot.fill_(-1)
            self._host_src = torch.empty((b,), dtype=torch.int32, device=dev)
            self._hist_src = torch.empty((b,), dtype=torch.int32, device=dev)
        self._new_rings(zero)
        a = self._args
        a.visited, a.steps, a.has_stopped = self._visited_words.data_ptr(), self.steps.data_ptr(), self.has_stopped.data_ptr()
        if self._history_set is not None:
            a.first_slot, a.host_src, a.history_src = (self._first_slot.data_ptr(), self._host_src.data_ptr(),
                                                       self._hist_src.data_ptr())

    def _new_rings(self, zero: bool = False):
        b, dev, n = self.batch_size, self.device, self.max_ep_len
        self._pos_ring = (torch.zeros if zero else torch.empty)((n + 1, b, 2), dtype=torch.long, device=dev)
        self._rew_ring = torch.empty((n, b), dtype=torch.float32, device=dev)
        self._term_ring = torch.empty((n, b), dtype=torch.bool, device=dev)
        self._trunc_ring = torch.empty((n, b), dtype=torch.bool, device=dev)
        # one C++ call per ring instead of one python slicing per step and tensor
        self._pos_rows, self._rew_rows = self._pos_ring.unbind(0), self._rew_ring.unbind(0)
        self._term_rows, self._trunc_rows = self._term_ring.unbind(0), self._trunc_ring.unbind(0)
        self._pos_base, self._rew_base = self._pos_ring.data_ptr(), self._rew_ring.data_ptr()
        self._term_base, self._trunc_base = self._term_ring.data_ptr(), self._trunc_ring.data_ptr()
        self._ring_k = 0  # rows of the ring used so far (row k of the positions = state after k steps)
        self.positions = self._cur_row = self._pos_rows[0]

    def rollout_buffers(self) -> Tuple[Tensor, Tensor, Tensor]:
        """Step-major ``[t, B]`` rewards / terminated / truncated of the steps taken since ``reset`` (views of
        the rings; ``jolineedle_b200.reinforce.rollout_tail`` consumes them as they are -- no per-step stacking)."""
        k = self._ring_k
        if k != self._t:
            raise RuntimeError("the episode ran past max_ep_len: the rings only hold its last steps")
        return self._rew_ring[:k], self._term_ring[:k], self._trunc_ring[:k]

    def check_status(self):
        """Synchronise and raise if a kernel flagged invalid input (the reference raises at the
        call site; the kernels cannot, so they record it)."""
        flags = int(self._status.item())
        if flags & _cabi.STATUS_BAD_POSITION:
            raise IndexError("a position was outside the patch grid")
        if flags & _cabi.STATUS_BAD_ACTION:
            raise ValueError("an action code was outside [0, 8]")
        if flags & _cabi.STATUS_BAD_BOX:
            raise IndexError("a bounding box falls outside the patch grid")

    # ------------------------------------------------------------------------------------
    def _slot(self, t: int) -> Tensor:
        """``[B, 1, C, P, P]`` view of history slot ``t``."""
        v = self._slot_views[t]
        if v is None:
            v = self._slot_views[t] = self._history[:, t:t + 1]
        return v

    def _gather(self) -> Tensor:
        """Stand-alone gather at the current positions (the ``patches`` property, general_env.py:285-306; reset
        and step gather inside their own native call)."""
        if self.n_glimps_levels > 1:
            return self._gather_levels()
        if self._history is not None:
            out = self._history[:, self._t]
        else:
            out = torch.empty(self._tile_shape, dtype=self._tile_dtype, device=self.device)
        if self._history_set is not None:
            return self._gather_zero_copy(out).unsqueeze(1)
        if self._launch_gather is None:  # argument checks once per env, not once per call
            self._launch_gather = self._set.bind(normalize=self._normalize, focus=self._focus, engine=self._engine,
                                                 status=self._status, tag="step", shifts=self._shifts,
                                                 shifts_aligned=self._shifts_aligned)
        return self._launch_gather(self.positions, out).unsqueeze(1)  # [B, G=1, C, P, P]

    def _gather_zero_copy(self, out: Tensor) -> Tensor:
        """Slot ``t`` of the history from pinned host images, outside of a step: patches seen for the first time
        come over PCIe, revisited ones are copied from the slot that first held them."""
        b, dev = self.batch_size, self.device
        host_src = torch.empty((b,), dtype=torch.int32, device=dev)
        hist_src = torch.empty((b,), dtype=torch.int32, device=dev)
        with _cabi.on_device(dev):
            _cabi.check(self._lib.jn_visit_sources(
                self.positions.data_ptr(), self._first_slot.data_ptr(), b, self.n_vertical_patches,
                self.n_horizontal_patches, self._history.shape[1], self._t, host_src.data_ptr(), hist_src.data_ptr(),
                self._status.data_ptr(), self._stream()))
        self._set.gather(self.positions, src_index=host_src, out=out, normalize=self._normalize, focus=self._focus,
                         engine=self._engine, status=self._status, tag="step", shifts=self._shifts,
                         shifts_aligned=self._shifts_aligned)
        self._history_set.gather(None, src_index=hist_src, out=out, engine=self._engine, status=self._status,
                                 tag="step-reuse")
        self.host_tiles += (host_src >= 0).sum()
        return out

    def _gather_levels(self) -> Tensor:
        """``[B, G, C, P, P]``: the same patch out of every glimpse level (one gather per level; level l of
        episode b is image b*G + l of the stacked set)."""
        g = self.n_glimps_levels
        if self._history is not None:
            out = self._history[:, self._t * g:(self._t + 1) * g]
        else:
            out = torch.empty((self.batch_size, g) + self._set.out_shape(1, self._focus)[1:],
                              dtype=self._set.out_dtype(self._normalize), device=self.device)
        for level in range(g):
            self._set.gather(self.positions, src_index=self._level_src[level], out=out[:, level],
                             normalize=self._normalize, focus=self._focus, engine=self._engine, status=self._status,
                             tag="step")
        return out

    @property
    def patches(self) -> Tensor:  # general_env.py:285-306
        return self._gather()

    def patch_history(self, upto: Optional[int] = None) -> Tensor:
        """``[B, (t+1)*G, C, P, P]`` view of every crop produced so far (needs ``history=True``)."""
        if self._history is None:
            raise RuntimeError("construct the env with history=True to keep the crop history")
        return self._history[:, : ((self._t if upto is None else upto) + 1) * self.n_glimps_levels]

    def _call(self, fn, out: Optional[Tensor]):
        """One native call: state kernel + the gather of ``out`` behind it (events around it when bench.py asked
        for per-launch timings)."""
        a = self._args
        a.out = None if out is None else out.data_ptr()
        timing = _gather_mod.TIMING
        with _cabi.on_device(self.device):
            pair = timing.begin(self.device) if (timing is not None and out is not None) else None
            rc = fn(self._set_handle, self._hist_handle, self._args_ref, _cabi.stream_ptr(self.device))
            if pair is not None:
                timing.end(pair, "step", self.batch_size)
        if rc:
            _cabi.check(rc)

    def _crop_buffer(self, t: int) -> Tuple[Optional[Tensor], Tensor]:
        """(tensor the fused gather writes, ``[B, G, C, P, P]`` tensor handed to the caller) for slot ``t``."""
        if not self._fused:
            return None, None
        if self._history is not None:
            if t >= len(self._slot_views):
                raise RuntimeError(f"history=True keeps max_ep_len + 1 = {len(self._slot_views)} crops per episode; "
                                   f"step {t} does not fit (reset the env or build it without history)")
            view = self._slot(t)
            return view, view
        out = torch.empty(self._tile_shape, dtype=self._tile_dtype, device=self.device)
        return out, out.unsqueeze(1)

    def reset(self, positions: Optional[Tensor] = None) -> Tuple[Tensor, dict]:  # general_env.py:144-170
        self.init_env_variables(zero=False)
        row0 = self._pos_rows[0]
        if positions is not None:
            assert tuple(positions.shape) == (self.batch_size, 2)
            row0.copy_(positions, non_blocking=True)
        else:
            # host RNG in the reference's order: rows first, then columns, CPU default generator
            ys = torch.randint(low=0, high=self.n_vertical_patches, size=(self.batch_size,))
            xs = torch.randint(low=0, high=self.n_horizontal_patches, size=(self.batch_size,))
            staged = torch.empty((self.batch_size, 2), dtype=torch.long, pin_memory=True)
            torch.stack((ys, xs), dim=1, out=staged)  # pinned staging: the upload does not wait for the stream
            row0.copy_(staged, non_blocking=True)
        a = self._args
        a.pos_out, a.t = self._pos_base, 0
        if self._history_set is not None:
            self.host_tiles.zero_()
        out, patches = self._crop_buffer(0)
        self._call(self._lib.jn_env_reset_gather, out)
        if not self._fused:
            patches = self._gather()
        return patches, {"positions": self.positions}

    def step(self, actions: Tensor) -> Tuple[Tensor, Tensor, Tensor, Tensor, dict]:  # general_env.py:172-207
        b, dev = self.batch_size, self.device
        if actions.dtype != torch.long or actions.device != dev or not actions.is_contiguous():
            actions = actions.to(device=dev, dtype=torch.long).contiguous()
        assert actions.numel() == b
        k = self._ring_k
        if k >= self.max_ep_len:  # stepping past max_ep_len (truncated stays True): continue in fresh rings
            cur = self.positions
            self._new_rings()
            self._pos_rows[0].copy_(cur)
            self.positions = self._cur_row = self._pos_rows[0]
            k = 0
        pos = self.positions
        if pos is not self._cur_row:  # someone rebound env.positions (apply_movements): take it as it is
            pos = self.positions = pos.to(device=dev, dtype=torch.long).contiguous()
        a = self._args
        a.pos_in, a.actions = pos.data_ptr(), actions.data_ptr()
        a.pos_out = self._pos_base + (k + 1) * b * 16
        a.rewards = self._rew_base + k * b * 4
        a.terminated, a.truncated = self._term_base + k * b, self._trunc_base + k * b
        a.t = self._t + 1
        out, patches = self._crop_buffer(self._t + 1)
        self._call(self._lib.jn_env_step_gather, out)
        self._keep_actions = actions  # the kernels read it after this call returns
        self._t += 1
        self._ring_k = k + 1
        self.positions = self._cur_row = self._pos_rows[k + 1]
        if not self._fused:
            patches = self._gather()
        return patches, self._rew_rows[k], self._term_rows[k], self._trunc_rows[k], {"positions": self.positions}

    # ------------------------------------------------------------------------------------
    def _props(self, want_prop: bool, want_term: bool):
        prop = torch.empty((self.batch_size,), dtype=torch.float32, device=self.device) if want_prop else None
        term = torch.empty((self.batch_size,), dtype=torch.bool, device=self.device) if want_term else None
        with _cabi.on_device(self.device):
            _cabi.check(self._lib.jn_env_props(
                self._visited_words.data_ptr(), self._bbox_words.data_ptr(), self.has_stopped.data_ptr(),
                self.batch_size, self.n_vertical_patches, self.n_horizontal_patches, 1 if self.stop_enabled else 0,
                _cabi.ptr(prop), _cabi.ptr(term), self._stream()))
        return prop, term

    @property
    def terminated(self) -> Tensor:  # general_env.py:235-246
        return self._props(False, True)[1]

    @property
    def prop_patches_found(self) -> Tensor:  # general_env.py:308-315
        return self._props(True, False)[0]

    @property
    def prop_bboxes_found(self) -> Tensor:  # general_env.py:317-319
        return (self.prop_patches_found > 0).to(torch.float32)

    @property
    def rewards(self) -> Tensor:
        """Reward of the CURRENT state (general_env.py:321-358).  ``step`` evaluates it between the move and
        the visited-map update; called on its own it sees whatever the visited map holds now."""
        out = torch.empty((self.batch_size,), dtype=torch.float32, device=self.device)
        with _cabi.on_device(self.device):
            _cabi.check(self._lib.jn_env_rewards(
                self.positions.data_ptr(), self._visited_words.data_ptr(), self._bbox_words.data_ptr(),
                self.has_stopped.data_ptr(), self.batch_size, self.n_vertical_patches, self.n_horizontal_patches,
                self._cost, 1 if self.stop_enabled else 0, out.data_ptr(), self._stream()))
        return out

    def convert_bboxes_to_masks(self, bboxes: Tensor) -> Tensor:
        """``[B, rows, cols]`` bool, patch holds a pixel of some box (general_env.py:360-379), for any
        ``[B, N, 4]`` box tensor -- the closed form of rasterise + max-pool, computed by K0."""
        boxes = torch.as_tensor(bboxes).to(device=self.device, dtype=torch.int64).contiguous()
        b = boxes.shape[0]
        words = torch.empty((b, self._words), dtype=torch.int32, device=self.device)
        out = torch.empty((b, self.n_vertical_patches, self.n_horizontal_patches), dtype=torch.bool, device=self.device)
        with _cabi.on_device(self.device):
            _cabi.check(self._lib.jn_patch_bitmaps(
                boxes.data_ptr(), None, b, boxes.shape[1], self.patch_size, self.n_vertical_patches,
                self.n_horizontal_patches, None, None, _cabi.RULE_ANY_PIXEL, words.data_ptr(), self._words,
                self._stream()))
            _cabi.check(self._lib.jn_bitmap_unpack(words.data_ptr(), b, self.n_vertical_patches,
                                                   self.n_horizontal_patches, out.data_ptr(), self._stream()))
        return out

    def actions_to_movements(self, actions: Tensor) -> Tensor:
        """``[B, 2]`` (dy, dx) of each action code (general_env.py:209-212) -- a table lookup on the
        device instead of one ``.item()`` round trip per episode."""
        table = torch.tensor(DELTA_TABLE, dtype=torch.long, device=actions.device)
        return table[actions.long()]

    def apply_movements(self, actions: Tensor):
        """Move, clamp to the grid, remember STOP (general_env.py:214-233).  ``step`` does this inside K2;
        the stand-alone method exists for callers that drive the pieces themselves."""
        actions = actions.to(self.device)
        moved = self.positions + self.actions_to_movements(actions)
        moved[:, 0].clamp_(0, self.n_vertical_patches - 1)
        moved[:, 1].clamp_(0, self.n_horizontal_patches - 1)
        self.positions = moved
        self.has_stopped |= actions == Action.STOP.value

    @property
    def tiles_reached(self) -> Tensor:  # general_env.py:248-283
        hot = torch.zeros((self.batch_size, self.n_vertical_patches, self.n_horizontal_patches), dtype=torch.bool,
                          device=self.device)
        hot[torch.arange(self.batch_size, device=self.device), self.positions[:, 0], self.positions[:, 1]] = True
        return hot

    # ------------------------------------------------------------------------------------
    # detection side (general_env.py:381-573): K0 split table + host bookkeeping + K1 gather
    def parse_bboxes(self, bboxes=None) -> Tuple[Tensor, Tensor]:
        """``[B, rows, cols, N, 4]`` local boxes and ``[B, rows, cols, N]`` presence
        (general_env.py:381-504)."""
        boxes = self._boxes_dev if bboxes is None else torch.as_tensor(bboxes).to(self.device, torch.int64).contiguous()
        b, n = boxes.shape[0], boxes.shape[1]
        rows, cols = self.n_vertical_patches, self.n_horizontal_patches
        local = torch.empty((b, rows, cols, n, 4), dtype=torch.long, device=self.device)
        present = torch.empty((b, rows, cols, n), dtype=torch.bool, device=self.device)
        with _cabi.on_device(self.device):
            _cabi.check(self._lib.jn_split_boxes(boxes.data_ptr(), b, n, self.patch_size, rows, cols,
                                                 local.data_ptr(), present.data_ptr(), self._status.data_ptr(),
                                                 self._stream()))
        return local, present

    @torch.no_grad()
    def get_detection_batch(self, sample_neg: int = 1):  # general_env.py:506-546
        """Every (patch, box) hit + ``sample_neg`` random misses per image, gathered by K1.  The reference walks
        the images in python (two ``torch.where``, a ``randperm``, two ``cat`` per image: ~40 ms at 1024 images);
        here only the ``torch.randperm`` calls stay per image -- same CPU generator, same order, so the same
        negatives are drawn -- and the bookkeeping is one vectorised pass over the presence table."""
        import numpy as np

        local, present = self.parse_bboxes()
        present = present.squeeze(-1)  # only collapses when there is a single box per image (reference quirk)
        table = present.cpu().numpy()
        b = self.batch_size
        flat = table.reshape(b, -1)  # row-major like torch.where: (row, col[, box])
        inner = flat.shape[1] // (self.n_vertical_patches * self.n_horizontal_patches)
        hit_img, hit_cell = np.nonzero(flat)
        miss_img, miss_cell = np.nonzero(~flat)
        n_miss = np.bincount(miss_img, minlength=b)
        miss_start = np.concatenate(([0], np.cumsum(n_miss)[:-1]))
        picks = [torch.randperm(int(n))[:sample_neg] for n in n_miss]  # CPU generator, one draw per image
        pick_img = np.repeat(np.arange(b), [len(p) for p in picks])
        pick_cell = miss_cell[np.concatenate([p.numpy() + s for p, s in zip(picks, miss_start)])] if len(pick_img) else \
            np.zeros(0, dtype=np.int64)
        img_all = np.concatenate((hit_img, pick_img))
        cell_all = np.concatenate((hit_cell, pick_cell))
        order = np.argsort(img_all * 2 + np.concatenate((np.zeros(len(hit_img), np.int64), np.ones(len(pick_img), np.int64))),
                           kind="stable")  # per image: its hits (table order), then its negatives (draw order)
        img_all, cell_all = img_all[order], cell_all[order] // inner
        host = np.empty((len(img_all), 3), dtype=np.int64)
        host[:, 0] = cell_all // self.n_horizontal_patches
        host[:, 1] = cell_all % self.n_horizontal_patches
        host[:, 2] = img_all
        dev_idx = torch.from_numpy(host).to(self.device)
        positions = dev_idx[:, :2].contiguous()
        src_l = dev_idx[:, 2].contiguous()
        src = src_l.to(torch.int32)
        # glimpse level 0 of image i is image i*G of the stacked set (general_env.py:533-539)
        patches = self._set.gather(positions, src_index=src * self.n_glimps_levels, normalize=self._normalize,
                                   engine=self._engine, status=self._status, shifts=self._shifts,
                                   shifts_aligned=self._shifts_aligned)
        boxes = local[src_l, positions[:, 0], positions[:, 1]]  # [n, N, 4]
        boxes = torch.nn.functional.pad(boxes, (1, 0))  # class id 0 in front
        return patches, boxes

    @torch.no_grad()
    def get_detection_targets(self) -> List[Tensor]:  # general_env.py:548-573
        """Global-coordinate split boxes per image (all-zero entries skipped, general_env.py:560), row-major over
        (y, x, k) like the reference's loops: one masked select for the whole batch and one split, instead of a
        masked select (= a device synchronisation) per image."""
        local, _ = self.parse_bboxes()
        rows, cols, p = self.n_vertical_patches, self.n_horizontal_patches, self.patch_size
        ys = torch.arange(rows, device=self.device).view(1, rows, 1, 1)
        xs = torch.arange(cols, device=self.device).view(1, 1, cols, 1)
        offset = torch.stack((xs.expand(1, rows, cols, 1), ys.expand(1, rows, cols, 1)) * 2, dim=-1) * p
        glob = local + offset  # x1+ox, y1+oy, x2+ox, y2+oy
        keep = local.abs().sum(dim=-1) != 0
        counts = keep.view(self.batch_size, -1).sum(dim=1).tolist()
        if 0 in counts:
            raise RuntimeError("stack expects a non-empty TensorList")  # what torch.stack([]) raises in the reference
        chosen = torch.nn.functional.pad(glob[keep], (1, 0))  # class id 0 in front; batch-major, then (y, x, k)
        return list(torch.split(chosen, counts))
