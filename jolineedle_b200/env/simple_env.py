"""Per-image supervised gaze environment on B200 -- drop-in for the reference's
``NeedleSimpleEnv`` (``src/env/simple_env.py:166-763``) and for the batched entry point
``SupervisedTrainer.generate_trajectories`` (``src/supervised.py:95-136``).

Division of labour
  host    every decision that consumes randomness or depends on python ``set`` iteration
          order: start position, greedy key-point order (ties via ``random.choice``), number
          and placement of random key points, binomial / uniform detours, the replacement
          moves for STOP.  The RNG calls are made in the reference's order, so a seeded env
          reproduces the reference's trajectory exactly.  This is O(key points) integer work:
          the *plan*.
  device  everything proportional to the trajectory length or to pixels: the 5 %-area overlap
          bitmaps (K0), expanding the plan into per-step positions / actions / best actions /
          labels with tail truncation (K3, ``jn_traj_expand``), the per-step local boxes, and
          the glimpse gather of ``[B, T, C, P, P]`` plus the detection patches (K1).
"""
import random
from itertools import product
from math import floor
from typing import Dict, List, Optional, Sequence, Set, Union

import numpy as np
import torch

from .. import _cabi
from ..gather import ImageSet
from ..utils import BBox, Position
from .common import Action, DELTA_TABLE, MOVES, STOP_CODE, direction_code

_LEFT = Action.LEFT.value


def pixel_pos_to_patch_pos(pixel_position: Position, patch_size: int) -> Position:
    """Patch containing a pixel (simple_env.py:13-18)."""
    return Position(y=floor(pixel_position.y / patch_size), x=floor(pixel_position.x / patch_size))


def move_towards(current_position: Position, target_position: Position) -> Action:
    """Greedy 8-neighbour action towards the target, STOP on arrival (simple_env.py:84-125)."""
    return Action(direction_code(target_position[0] - current_position[0], target_position[1] - current_position[1]))


def apply_action(current_position: Position, action: Action) -> Position:
    dy, dx = DELTA_TABLE[action.value]
    return Position(current_position.y + dy, current_position.x + dx)


class EpisodePlan:
    """Host-side description of one supervised episode (input of ``jn_traj_expand``)."""

    __slots__ = ("start", "seg_to", "seg_tgt", "seg_first", "draws", "det_positions")

    def __init__(self, start: Position):
        self.start = start
        self.seg_to: List[Position] = []
        self.seg_tgt: List[Position] = []
        self.seg_first: List[int] = []
        self.draws: List[int] = []
        self.det_positions: List[Position] = []


class NeedleSimpleEnv:
    def __init__(
        self,
        image: torch.Tensor,
        patch_size: int,
        bboxes: List[BBox],
        seed: Optional[int] = None,
        *,
        normalize: bool = False,
        device=None,
    ):
        if device is not None and not image.is_cuda:
            image = image.to(device, non_blocking=True)
        self.image = image
        self.patch_size = patch_size
        self.rng = np.random.default_rng(seed)
        self.raw_bboxes = bboxes
        self.bboxes = [
            BBox(pixel_pos_to_patch_pos(b.up_left, patch_size), pixel_pos_to_patch_pos(b.bottom_right, patch_size))
            for b in bboxes
        ]
        self.position = Position(0, 0)
        self.infos: dict = {}
        self.n_channels, self.height, self.width = image.shape
        self.patch_height = self.height // patch_size
        self.patch_width = self.width // patch_size
        self._normalize = normalize
        # the union is rebuilt box by box so that the set's iteration order is the reference's
        self.bbox_patches: Set[Position] = set()
        for box in bboxes:
            self.bbox_patches = self.bbox_patches | self.bbox_positions(box)
        self.visited_bbox_patches: Set[Position] = set()
        self._image_set: Optional[ImageSet] = None

    def __deepcopy__(self, memo):
        """``deepcopy(env)`` (supervised.py:295) copies the mutable episode state and shares the
        immutable image and its device handle."""
        import copy

        clone = object.__new__(type(self))
        memo[id(self)] = clone
        for key, value in self.__dict__.items():
            if key in ("image", "_image_set"):
                setattr(clone, key, value)
            else:
                setattr(clone, key, copy.deepcopy(value, memo))
        return clone

    # -- geometry (host integers) ------------------------------------------------------------
    def bbox_positions(self, raw_bbox: BBox, area_threshold: float = 0.05) -> Set[Position]:
        """Patches holding more than ``area_threshold`` of P^2 of the box, plus the patch of its
        centre, inside the grid (simple_env.py:270-321).  Insertion order follows the reference
        because set iteration order leaks into tie-breaks and into ``patches_yolox``."""
        p = self.patch_size
        lo = pixel_pos_to_patch_pos(raw_bbox.up_left, p)
        hi = pixel_pos_to_patch_pos(raw_bbox.bottom_right, p)
        cells: Set[Position] = set()
        for y, x in product(range(lo.y, hi.y + 1), range(lo.x, hi.x + 1)):
            oh = min((y + 1) * p, raw_bbox.bottom_right.y) - max(y * p, raw_bbox.up_left.y)
            ow = min((x + 1) * p, raw_bbox.bottom_right.x) - max(x * p, raw_bbox.up_left.x)
            if oh * ow / (p**2) > area_threshold:
                cells.add(Position(y, x))
        centre = Position(
            y=(raw_bbox.up_left.y + raw_bbox.bottom_right.y) // 2,
            x=(raw_bbox.up_left.x + raw_bbox.bottom_right.x) // 2,
        )
        cells.add(pixel_pos_to_patch_pos(centre, p))
        cells = {c for c in cells if 0 <= c.x < self.patch_width}
        cells = {c for c in cells if 0 <= c.y < self.patch_height}
        return cells

    def local_bboxes(self, position: Optional[Position] = None) -> torch.Tensor:
        """``[N, 6]`` rows ``(0, x1, y1, x2, y2, 1)`` of each raw box clipped to the patch, in
        patch-local pixels; zero rows where they do not meet (simple_env.py:231-268).  Single
        query, computed on the host; the batched path uses ``jn_local_boxes``."""
        position = self.position if position is None else position
        p = self.patch_size
        ox, oy = position[1] * p, position[0] * p
        rows = []
        for box in self.raw_bboxes:
            x1, y1 = max(ox, box.up_left.x), max(oy, box.up_left.y)
            x2, y2 = min(ox + p, box.bottom_right.x), min(oy + p, box.bottom_right.y)
            rows.append([0, x1 - ox, y1 - oy, x2 - ox, y2 - oy, 1] if (x1 < x2 and y1 < y2) else [0] * 6)
        return torch.tensor(rows, dtype=torch.float32).reshape(len(self.raw_bboxes), 6)

    # -- single-step API (eval loops of the reference use it) ------------------------------------
    def _ensure_set(self) -> ImageSet:
        if self._image_set is None:
            self._image_set = ImageSet(self.image, self.patch_size)
        return self._image_set

    def get_patch(self, position: Position) -> torch.Tensor:
        """``[C, P, P]`` crop at ``position`` (simple_env.py:55-81) through K1."""
        assert 0 <= position[0] < self.patch_height
        assert 0 <= position[1] < self.patch_width
        s = self._ensure_set()
        pos = torch.tensor([[int(position[0]), int(position[1])]], dtype=torch.long, device=s.device)
        return s.gather(pos, normalize=self._normalize)[0]

    def gather_infos(self) -> dict:  # simple_env.py:208-229
        infos = {
            "position": self.position,
            "number_patches_found": len(self.visited_bbox_patches),
            "local_bboxes": self.local_bboxes(),
            "inside_bbox": self.position in self.bbox_patches,
        }
        self.infos = dict(infos)
        return infos

    def _place(self, position: Optional[Position], visited: Optional[Set[Position]]):
        if position is None:  # y first, then x (simple_env.py:330-333)
            position = Position(
                y=self.rng.integers(low=0, high=self.patch_height),
                x=self.rng.integers(low=0, high=self.patch_width),
            )
        self.position = position
        self.visited_bbox_patches = set() if visited is None else visited
        if self.position in self.bbox_patches:
            self.visited_bbox_patches.add(self.position)

    def reset(self, position: Optional[Position] = None, visited_bbox_patches: Optional[Set[Position]] = None):
        """simple_env.py:323-345 (``visited_bbox_patches=None`` clears the visited set)."""
        self._place(position, visited_bbox_patches)
        return self.get_patch(self.position), self.gather_infos()

    def _move(self, code: int):
        dy, dx = DELTA_TABLE[code]
        self.position = Position(
            min(max(self.position[0] + dy, 0), self.patch_height - 1),
            min(max(self.position[1] + dx, 0), self.patch_width - 1),
        )
        if self.position in self.bbox_patches:
            self.visited_bbox_patches.add(self.position)

    def step(self, move: Union[Action, np.ndarray]):  # simple_env.py:347-376
        move = Action(move.item()) if type(move) is np.ndarray else move
        self._move(move.value)
        infos = self.gather_infos()
        return self.get_patch(self.position), infos

    # -- key points (host RNG, reference draw order) ----------------------------------------------
    def generate_keypoints(self, n_keypoints: int) -> list:  # simple_env.py:666-682
        out = []
        for _ in range(n_keypoints):
            y = self.rng.integers(0, self.patch_height)
            x = self.rng.integers(0, self.patch_width)
            out.append(Position(y, x))
        return out

    def generate_binomial_keypoints(self, n_keypoints: int, target_pos: Position) -> list:  # simple_env.py:684-713
        out = []
        for _ in range(n_keypoints):
            dx = self.rng.binomial(self.patch_width, 0.5) - self.patch_width // 2  # x is drawn first
            dy = self.rng.binomial(self.patch_height, 0.5) - self.patch_height // 2
            out.append(Position((target_pos[0] + dy) % self.patch_height, (target_pos[1] + dx) % self.patch_width))
        return out

    def remove_stop_action(self, action: Action) -> Action:  # simple_env.py:715-718
        return self.rng.choice(MOVES) if action == Action.STOP else action

    def build_keypoints_trajectory(self) -> List[Position]:
        """Greedy nearest-first (L1) order of the unvisited box patches; ties through python's
        global ``random.choice`` over candidates in set order (simple_env.py:590-629)."""
        todo: Set[Position] = set()
        for box in self.raw_bboxes:
            todo |= self.bbox_positions(box)
        for seen in self.visited_bbox_patches:
            todo.remove(seen)
        here, order = self.position, []
        while todo:
            best, ties = None, []
            for c in todo:
                d = abs(c.x - here.x) + abs(c.y - here.y)
                if best is None or d < best:
                    best, ties = d, []
                if d == best:
                    ties.append(c)
            here = random.choice(ties)
            order.append(here)
            todo.remove(here)
        if not order:
            order.append(self.generate_keypoints(1)[0])
            if len(self.visited_bbox_patches) == 0:
                print("Warning: no keypoints found, either the image is empty or the bbox is outside the bound of the image.")
        return order

    # -- planning ----------------------------------------------------------------------------------
    def _plan_walk(self, plan: EpisodePlan, to_visit: Position, true_target: Position, first: int):
        """One straight-line segment (simple_env.py:631-664): only the events that consume
        randomness are simulated here -- a replacement move each time the walk stands on
        ``true_target`` -- the per-step records are produced on the device."""
        plan.seg_to.append(Position(int(to_visit[0]), int(to_visit[1])))
        plan.seg_tgt.append(Position(int(true_target[0]), int(true_target[1])))
        plan.seg_first.append(first)
        self._place(self.position, None)
        ty, tx = int(true_target[0]), int(true_target[1])
        vy, vx = int(to_visit[0]), int(to_visit[1])
        y, x = int(self.position[0]), int(self.position[1])
        while y != vy or x != vx:
            y += (vy > y) - (vy < y)
            x += (vx > x) - (vx < x)
            if y == ty and x == tx:
                plan.draws.append(int(self.rng.choice(8)))  # == rng.choice(MOVES)
        self._place(Position(y, x), None)

    def plan_sample(
        self,
        min_keypoints: int,
        max_keypoints: int,
        binomial_keypoints: bool = False,
        position: Optional[Position] = None,
        visited_bbox_patches: Optional[Set[Position]] = None,
    ) -> EpisodePlan:
        """The host half of ``generate_sample`` (simple_env.py:481-588), RNG calls in the
        reference's order: detection-patch pick, start position, greedy order ties, number of
        random key points, their slots, then per key point the opening replacement (if the
        agent already stands on it), detours and walk events."""
        # init_sample (simple_env.py:397-419): every box patch + one random empty patch
        det: Set[Position] = set()
        for box in self.raw_bboxes:
            for c in self.bbox_positions(box):
                det.add(c)
        empties = [Position(y, x) for y, x in product(range(self.patch_height), range(self.patch_width))
                   if Position(y, x) not in det]
        if empties:
            det.add(empties[self.rng.choice(len(empties))])
        self._place(position, visited_bbox_patches)
        plan = EpisodePlan(Position(int(self.position[0]), int(self.position[1])))
        plan.det_positions = [Position(int(c[0]), int(c[1])) for c in det]
        keypoints = self.build_keypoints_trajectory()
        n_random = self.rng.integers(min_keypoints, max_keypoints + 1)
        slots = sorted(self.rng.integers(0, len(keypoints), size=n_random), reverse=True)
        for k, keypoint in enumerate(keypoints):
            first = 1
            if self.position[0] == keypoint[0] and self.position[1] == keypoint[1]:
                plan.draws.append(int(self.rng.choice(8)))  # opening best action would be STOP
            while k in slots:
                detour = (self.generate_binomial_keypoints(1, keypoint) if binomial_keypoints
                          else self.generate_keypoints(1))[0]
                self._plan_walk(plan, detour, keypoint, first)
                first = 0
                slots.remove(k)
            self._plan_walk(plan, keypoint, keypoint, first)
        return plan

    def generate_sample(
        self,
        max_ep_len: int,
        min_keypoints: int,
        max_keypoints: int,
        binomial_keypoints: bool = False,
        position: Optional[Position] = None,
        visited_bbox_patches: Optional[Set[Position]] = None,
        device: str = "cpu",
    ) -> dict:
        """One episode (simple_env.py:481-588).  Tensors stay on the image's CUDA device; the
        reference's ``device`` argument is accepted and ignored when it says "cpu"."""
        plan = self.plan_sample(min_keypoints, max_keypoints, binomial_keypoints, position, visited_bbox_patches)
        batch = expand_plans([self], [plan], max_ep_len, image_set=self._ensure_set())
        batch.pop("_ep_len")
        batch.pop("_status")
        sample = {k: (v[0] if k not in ("patches_yolox", "bboxes_yolox") else v) for k, v in batch.items()}
        assert sample["patches"].shape[0] == max_ep_len
        return sample

    # -- collate -------------------------------------------------------------------------------------
    @staticmethod
    def collate_fn(batch: List[dict]) -> dict:
        """Pad the box axis to the batch maximum, stack per-step tensors, concatenate the
        detection patches (simple_env.py:720-763)."""
        n_max = max(s["local_bboxes"].shape[1] for s in batch)

        def widen(t: torch.Tensor) -> torch.Tensor:
            pad = torch.zeros((t.shape[0], n_max - t.shape[1], t.shape[2]), dtype=torch.float32, device=t.device)
            return torch.cat((t, pad), dim=1)

        det_tiles = [s.pop("patches_yolox") for s in batch]
        det_boxes = [widen(s.pop("bboxes_yolox")) for s in batch]
        out = {}
        for key in batch[0]:
            out[key] = torch.stack([widen(s[key]) if key == "local_bboxes" else s[key] for s in batch])
        out["patches_yolox"] = torch.cat(det_tiles)
        out["bboxes_yolox"] = torch.cat(det_boxes)
        return out


# ---------------------------------------------------------------------------------------------------
# device half: plans -> tensors
# ---------------------------------------------------------------------------------------------------
def expand_plans(
    envs: Sequence[NeedleSimpleEnv],
    plans: Sequence[EpisodePlan],
    max_ep_len: int,
    image_set: Optional[ImageSet] = None,
    normalize: Optional[bool] = None,
    engine: str = "auto",
) -> Dict[str, torch.Tensor]:
    """Turn host plans into the collated sample dict of the reference (keys ``patches``,
    ``current_actions``, ``next_actions``, ``positions``, ``masks``, ``labels``,
    ``local_bboxes``, ``patches_yolox``, ``bboxes_yolox``) with K0 + K3 + K1."""
    lib = _cabi.lib()
    n, T = len(envs), int(max_ep_len)
    if image_set is None:
        image_set = ImageSet([e.image for e in envs], envs[0].patch_size)
    if normalize is None:
        normalize = envs[0]._normalize
    dev, P = image_set.device, image_set.patch_size
    n_max = max(len(e.raw_bboxes) for e in envs)

    # ---- pack the plans (host) ----
    seg_begin = np.zeros(n + 1, dtype=np.int32)
    draw_begin = np.zeros(n + 1, dtype=np.int32)
    det_begin = np.zeros(n + 1, dtype=np.int32)
    for i, pl in enumerate(plans):
        seg_begin[i + 1] = seg_begin[i] + len(pl.seg_to)
        draw_begin[i + 1] = draw_begin[i] + len(pl.draws)
        det_begin[i + 1] = det_begin[i] + len(pl.det_positions)
    n_seg, n_draw, n_det = int(seg_begin[-1]), int(draw_begin[-1]), int(det_begin[-1])
    start = np.array([pl.start for pl in plans], dtype=np.int32).reshape(n, 2)
    seg_to = np.array([c for pl in plans for c in pl.seg_to], dtype=np.int32).reshape(n_seg, 2)
    seg_tgt = np.array([c for pl in plans for c in pl.seg_tgt], dtype=np.int32).reshape(n_seg, 2)
    rows = np.array([e.patch_height for e in envs], dtype=np.int32)
    cols = np.array([e.patch_width for e in envs], dtype=np.int32)
    n_boxes = np.array([len(e.raw_bboxes) for e in envs], dtype=np.int32)
    det_src = np.repeat(np.arange(n, dtype=np.int32), np.diff(det_begin))
    i32 = np.concatenate([start.ravel(), seg_begin, seg_to.ravel(), seg_tgt.ravel(), draw_begin, rows, cols, n_boxes,
                          det_src])
    u8 = np.concatenate([np.array([f for pl in plans for f in pl.seg_first], dtype=np.uint8),
                         np.array([d for pl in plans for d in pl.draws], dtype=np.uint8),
                         np.zeros(1, dtype=np.uint8)])
    boxes = np.zeros((n, max(n_max, 1), 4), dtype=np.int64)
    for i, e in enumerate(envs):
        for k, b in enumerate(e.raw_bboxes):
            boxes[i, k] = (int(b.up_left.x), int(b.up_left.y), int(b.bottom_right.x), int(b.bottom_right.y))
    det_pos = np.array([c for pl in plans for c in pl.det_positions], dtype=np.int64).reshape(n_det, 2)
    i64 = np.concatenate([boxes.ravel(), det_pos.ravel()])

    d_i32 = torch.from_numpy(i32).to(dev, non_blocking=True)
    d_u8 = torch.from_numpy(u8).to(dev, non_blocking=True)
    d_i64 = torch.from_numpy(i64).to(dev, non_blocking=True)

    def take(buf, offset, count):
        return buf[offset : offset + count], offset + count

    o = 0
    d_start, o = take(d_i32, o, 2 * n)
    d_seg_begin, o = take(d_i32, o, n + 1)
    d_seg_to, o = take(d_i32, o, 2 * n_seg)
    d_seg_tgt, o = take(d_i32, o, 2 * n_seg)
    d_draw_begin, o = take(d_i32, o, n + 1)
    d_rows, o = take(d_i32, o, n)
    d_cols, o = take(d_i32, o, n)
    d_nboxes, o = take(d_i32, o, n)
    d_det_src, o = take(d_i32, o, n_det)
    d_flags, d_draws = d_u8[:n_seg], d_u8[n_seg : n_seg + n_draw + 1]
    d_boxes = d_i64[: boxes.size].view(n, max(n_max, 1), 4)
    d_det_pos = d_i64[boxes.size :].view(n_det, 2)

    words = int(max((r * c + 31) // 32 for r, c in zip(rows.tolist(), cols.tolist())))
    stream = _cabi.stream_ptr(dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    area = torch.empty((n, words), dtype=torch.int32, device=dev)
    out = {
        "patches": torch.empty((n, T) + image_set.out_shape(1, False)[1:], dtype=torch.float32, device=dev),
        "current_actions": torch.empty((n, T), dtype=torch.long, device=dev),
        "next_actions": torch.empty((n, T), dtype=torch.long, device=dev),
        "positions": torch.empty((n, T, 2), dtype=torch.long, device=dev),
        "masks": torch.empty((n, T), dtype=torch.float32, device=dev),
        "labels": torch.empty((n, T), dtype=torch.long, device=dev),
        "local_bboxes": torch.empty((n, T, n_max, 6), dtype=torch.float32, device=dev),
    }
    if image_set.out_dtype(normalize) != torch.float32:
        raise ValueError("supervised samples are float32: pass float32 images, or uint8 images with normalize=True")
    gather_src = torch.empty((n, T), dtype=torch.int32, device=dev)
    ep_len = torch.empty((n,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        # K0: 5 %-area bitmaps (labels = inside_bbox, simple_env.py:225,478)
        _cabi.check(lib.jn_patch_bitmaps(d_boxes.data_ptr(), d_nboxes.data_ptr(), n, max(n_max, 1), P, 0, 0,
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
            _cabi.check(lib.jn_local_boxes(d_boxes.data_ptr(), d_nboxes.data_ptr(), n_max, P,
                                           out["positions"].data_ptr(), gather_src.data_ptr(), n * T,
                                           out["local_bboxes"].data_ptr(), stream))
    # K1: the glimpses themselves, straight into [B, T, C, P, P]; padded slots are zero-filled
    image_set.gather(out["positions"].view(n * T, 2), src_index=gather_src.view(n * T),
                     out=out["patches"].view((n * T,) + out["patches"].shape[2:]), normalize=normalize,
                     engine=engine, status=status, tag="trajectory")
    # detection patches: every box patch + one random empty patch per image (simple_env.py:397-441)
    out["patches_yolox"] = image_set.gather(d_det_pos, src_index=d_det_src, normalize=normalize, engine=engine,
                                            status=status, tag="detection")
    det_boxes = torch.empty((n_det, n_max, 6), dtype=torch.float32, device=dev)
    if n_max > 0 and n_det > 0:
        with torch.cuda.device(dev):
            _cabi.check(lib.jn_local_boxes(d_boxes.data_ptr(), d_nboxes.data_ptr(), n_max, P, d_det_pos.data_ptr(),
                                           d_det_src.data_ptr(), n_det, det_boxes.data_ptr(), stream))
    out["bboxes_yolox"] = det_boxes
    out["_ep_len"] = ep_len
    out["_status"] = status
    return out


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
) -> Dict[str, torch.Tensor]:
    """Batched supervised trajectories (``SupervisedTrainer.generate_trajectories``,
    supervised.py:95-136): ``batch`` holds lists ``image`` ([C,H,W] tensors), ``bboxes`` (lists
    of ``BBox``) and ``class_id``.  Returns the collated dict of the reference on the GPU.
    ``seeds`` (one per image) makes the plans reproducible; the reference builds unseeded envs."""
    images = batch["image"]
    envs = [
        NeedleSimpleEnv(images[i], patch_size, batch["bboxes"][i], None if seeds is None else seeds[i],
                        normalize=normalize, device=device)
        for i in range(len(images))
    ]
    plans = [e.plan_sample(min_keypoints, max_keypoints, binomial_keypoints, position) for e in envs]
    out = expand_plans(envs, plans, max_seq_len, normalize=normalize, engine=engine)
    dev = out["patches"].device
    out["class_id"] = torch.tensor([int(c) for c in batch["class_id"]], dtype=torch.long, device=dev)
    out.pop("_ep_len")
    out.pop("_status")
    return out
