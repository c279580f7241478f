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


def get_patch(image: torch.Tensor, patch_size: int, position: Position) -> torch.Tensor:
    """``[C, P, P]`` window of patch ``position`` = (y, x) (simple_env.py:55-81).  Like the reference this is
    a *view* of the image (no copy, no kernel); the batched paths use the K1 gather instead."""
    _, height, width = image.shape
    assert height % patch_size == 0
    assert width % patch_size == 0
    assert 0 <= position[0] < height // patch_size
    assert 0 <= position[1] < width // patch_size
    y, x = int(position[0]) * patch_size, int(position[1]) * patch_size
    return image[:, y:y + patch_size, x:x + patch_size]


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

    def visit_point(self, sample: dict, to_visit: Position, true_target: Position, device: str = "cpu"):
        """Walk straight to ``to_visit`` recording every step into ``sample`` (simple_env.py:631-664), for
        callers that build samples incrementally; ``generate_sample`` itself plans the walk on the host and
        expands it on the device.  The recorded best action points at ``true_target`` (a random move
        replaces STOP); the visited set is cleared after each step, as in the reference."""
        self.reset(self.position)
        index = int(sample["masks"].long().sum().item())
        while self.position != to_visit:
            action = move_towards(self.position, to_visit)
            patch, infos = self.step(action)
            infos["best_action"] = self.remove_stop_action(move_towards(self.position, true_target))
            self._place(self.position, None)
            self.add_to_sample(sample, action, patch, infos, index)
            index += 1

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

    # -- incremental sample API (the reference's eval loop fills a sample step by step) ------------------
    def init_sample(self, max_ep_len: int, device=None) -> dict:
        """Zeroed sample buffers plus the detection patches (simple_env.py:378-441).  Consumes one
        ``rng.choice`` like the reference; tensors live on the image's CUDA device."""
        s = self._ensure_set()
        dev, p, n = s.device, self.patch_size, len(self.raw_bboxes)
        sample = {
            "patches": torch.zeros((max_ep_len, self.n_channels, p, p), dtype=torch.float, device=dev),
            "current_actions": torch.zeros((max_ep_len,), dtype=torch.long, device=dev),
            "next_actions": torch.zeros((max_ep_len,), dtype=torch.long, device=dev),
            "positions": torch.zeros((max_ep_len, 2), dtype=torch.long, device=dev),
            "masks": torch.zeros((max_ep_len,), dtype=torch.float, device=dev),
            "labels": torch.zeros((max_ep_len,), dtype=torch.long, device=dev),
            "local_bboxes": torch.zeros((max_ep_len, n, 6), device=dev),
        }
        det: Set[Position] = set()
        for box in self.raw_bboxes:
            for c in self.bbox_positions(box):
                det.add(c)
        empties = [Position(y, x) for y, x in product(range(self.patch_height), range(self.patch_width))
                   if Position(y, x) not in det]
        if empties:
            det.add(empties[self.rng.choice(len(empties))])
        cells = list(det)
        if cells:
            pos = torch.tensor([[int(c[0]), int(c[1])] for c in cells], dtype=torch.long, device=dev)
            src = torch.zeros((len(cells),), dtype=torch.int32, device=dev)
            sample["patches_yolox"] = s.gather(pos, src_index=src, normalize=self._normalize).float()
            sample["bboxes_yolox"] = torch.stack([self.local_bboxes(c) for c in cells]).to(dev)
        else:  # fictitious entry (simple_env.py:421-436)
            sample["patches_yolox"] = torch.zeros((1, self.n_channels, p, p), dtype=torch.float, device=dev)
            sample["bboxes_yolox"] = torch.zeros((1, n, 6), dtype=torch.float, device=dev)
        return sample

    def add_to_sample(self, sample: dict, action_taken: Action, patch: torch.Tensor, infos: dict, index: int):
        """Record one step (simple_env.py:443-479); buffers double in length when full."""
        if sample["patches"].shape[0] <= index:
            for key in sample:
                if key in ("patches_yolox", "bboxes_yolox"):
                    continue
                sample[key] = torch.cat([sample[key], torch.zeros_like(sample[key])], dim=0)
        sample["patches"][index] = patch
        sample["current_actions"][index] = action_taken.value
        sample["next_actions"][index] = infos["best_action"].value
        sample["positions"][index, 0] = int(infos["position"][0])
        sample["positions"][index, 1] = int(infos["position"][1])
        sample["masks"][index] = 1.0
        sample["labels"][index] = int(infos["inside_bbox"])
        sample["local_bboxes"][index] = infos["local_bboxes"].to(sample["local_bboxes"].device)

    def best_next_action(self, position: Optional[Position] = None,
                         visited_bbox_patches: Optional[Set[Position]] = None) -> Action:
        """The expert's next action from ``position``: what the reference's eval loop obtains by generating a
        whole 50-step ``generate_sample(50, 0, 0, position, visited)`` and reading ``next_actions[0]``
        (supervised.py:301-309,340-348).  Only the host plan is made here -- no buffer, no pixel -- and the
        same random draws are consumed (one ``random.choice`` per greedy key point, the ``rng`` calls of the
        plan), so interleaving it with the reference's call order keeps both streams in step."""
        plan = self.plan_sample(0, 0, False, position, visited_bbox_patches)
        target = plan.seg_tgt[0]
        code = direction_code(target[0] - plan.start[0], target[1] - plan.start[1])
        if code == STOP_CODE:
            code = plan.draws[0]
        return Action(code)

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
        sample = {k: (v[0] if k not in ("patches_yolox", "bboxes_yolox") else v) for k, v in batch.items()
                  if not k.startswith("_")}
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
# device half (plans -> tensors) and the batched entry point live in trajectories.py
# ---------------------------------------------------------------------------------------------------
def expand_plans(envs, plans, max_ep_len, image_set=None, normalize=None, engine="auto"):
    """Turn host plans of these envs into the collated sample dict (see trajectories.py)."""
    from .trajectories import expand_packed, pack_python_plans

    if image_set is None:
        image_set = ImageSet([e.image for e in envs], envs[0].patch_size)
    if normalize is None:
        normalize = envs[0]._normalize
    return expand_packed(image_set, pack_python_plans(envs, plans), max_ep_len, normalize, engine)


def generate_trajectories(*args, **kwargs):
    """Batched supervised trajectories (supervised.py:95-136); see trajectories.py."""
    from .trajectories import generate_trajectories as impl

    return impl(*args, **kwargs)
