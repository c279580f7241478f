"""Oracle for the per-image supervised env (reference: ``src/env/simple_env.py``).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  The trajectory generator draws from a
numpy ``Generator`` and from python's global ``random`` and its decisions depend on the
iteration order of python ``set`` objects, so this restatement deliberately performs the
same sequence of set constructions / unions and the same sequence of RNG calls as the
reference; everything else is written as plain integer code on (y, x) tuples.

``Pos`` is a NamedTuple of two ints: it hashes like the reference's ``Position`` (a tuple
hash), which is what makes the set orders coincide.
"""
import math
import random
from itertools import product
from typing import Dict, List, NamedTuple, Optional, Sequence, Set

import numpy as np
import torch


class Pos(NamedTuple):
    y: int
    x: int


class Box(NamedTuple):
    up_left: Pos
    bottom_right: Pos


# action codes (src/env/common.py:4-15)
LEFT, RIGHT, UP, DOWN, LEFT_UP, RIGHT_UP, LEFT_DOWN, RIGHT_DOWN, STOP = range(9)
_DELTA = ((0, -1), (0, 1), (-1, 0), (1, 0), (-1, -1), (-1, 1), (1, -1), (1, 1), (0, 0))


def patch_of(pixel: Sequence, patch: int) -> Pos:
    """simple_env.py:13-18 -- floor of a true division (works for float pixels too)."""
    return Pos(math.floor(pixel[0] / patch), math.floor(pixel[1] / patch))


def heading(src: Sequence[int], dst: Sequence[int]) -> int:
    """simple_env.py:84-125 -- greedy 8-neighbour direction, diagonal while both deltas are
    non-zero, STOP on arrival."""
    gy, gx = dst[0] - src[0], dst[1] - src[1]
    if gx == 0:
        return DOWN if gy > 0 else UP if gy < 0 else STOP
    if gy == 0:
        return RIGHT if gx > 0 else LEFT
    if gy < 0:
        return RIGHT_UP if gx > 0 else LEFT_UP
    return RIGHT_DOWN if gx > 0 else LEFT_DOWN


def tile_view(image: torch.Tensor, patch: int, pos: Sequence[int]) -> torch.Tensor:
    """simple_env.py:55-81 -- the [C, P, P] window of patch (y, x); a view, like the einops
    rearrange + index of the reference."""
    _, h, w = image.shape
    assert h % patch == 0 and w % patch == 0
    assert 0 <= pos[0] < h // patch and 0 <= pos[1] < w // patch
    return image[:, pos[0] * patch : (pos[0] + 1) * patch, pos[1] * patch : (pos[1] + 1) * patch]


class TrajectoryOracle:
    """Restatement of ``NeedleSimpleEnv`` (simple_env.py:166-763)."""

    def __init__(self, image: torch.Tensor, patch_size: int, bboxes: List, seed: Optional[int] = None):
        # simple_env.py:167-206
        self.image, self.patch_size = image, patch_size
        self.rng = np.random.default_rng(seed)
        self.raw_bboxes = [Box(Pos(*b[0]), Pos(*b[1])) for b in bboxes]
        self.n_channels, self.height, self.width = image.shape
        self.grid_h, self.grid_w = self.height // patch_size, self.width // patch_size
        self.position = Pos(0, 0)
        self.bbox_patches: Set[Pos] = set()
        for box in self.raw_bboxes:
            self.bbox_patches = self.bbox_patches | self.patches_of_box(box)
        self.visited: Set[Pos] = set()

    # -- geometry ---------------------------------------------------------------------
    def patches_of_box(self, box: Box, area_threshold: float = 0.05) -> Set[Pos]:
        """simple_env.py:270-321 -- patches holding > 5 % of P^2 of the box, plus the patch of
        the box centre, restricted to the grid.  Built with the reference's insertion order."""
        p = self.patch_size
        first, last = patch_of(box.up_left, p), patch_of(box.bottom_right, p)
        found: Set[Pos] = set()
        for y, x in product(range(first.y, last.y + 1), range(first.x, last.x + 1)):
            top, left = max(y * p, box.up_left.y), max(x * p, box.up_left.x)
            bottom, right = min((y + 1) * p, box.bottom_right.y), min((x + 1) * p, box.bottom_right.x)
            if (bottom - top) * (right - left) / (p**2) > area_threshold:
                found.add(Pos(y, x))
        centre = Pos((box.up_left.y + box.bottom_right.y) // 2, (box.up_left.x + box.bottom_right.x) // 2)
        found.add(patch_of(centre, p))
        found = {q for q in found if 0 <= q.x < self.grid_w}
        found = {q for q in found if 0 <= q.y < self.grid_h}
        return found

    def local_boxes(self, pos: Optional[Pos] = None) -> torch.Tensor:
        """simple_env.py:231-268 -- per raw box its intersection with the patch in local
        x1,y1,x2,y2 pixels as ``[0, x1, y1, x2, y2, 1]``; zero row when they do not overlap."""
        pos = self.position if pos is None else pos
        p = self.patch_size
        out = torch.zeros((len(self.raw_bboxes), 6), dtype=torch.float32)
        px1, py1 = pos[1] * p, pos[0] * p
        px2, py2 = px1 + p, py1 + p
        for k, box in enumerate(self.raw_bboxes):
            x1, y1 = max(px1, box.up_left.x), max(py1, box.up_left.y)
            x2, y2 = min(px2, box.bottom_right.x), min(py2, box.bottom_right.y)
            if px1 <= x1 < x2 <= px2 and py1 <= y1 < y2 <= py2:
                out[k] = torch.tensor([0, x1 - px1, y1 - py1, x2 - px1, y2 - py1, 1], dtype=torch.float32)
        return out

    # -- single-step API ----------------------------------------------------------------
    def _infos(self) -> dict:  # simple_env.py:208-229
        return {
            "position": self.position,
            "number_patches_found": len(self.visited),
            "local_bboxes": self.local_boxes(),
            "inside_bbox": self.position in self.bbox_patches,
        }

    def reset(self, position: Optional[Pos] = None, visited: Optional[Set[Pos]] = None):
        # simple_env.py:323-345 -- NB: `visited=None` clears the visited set.
        if position is None:
            position = Pos(int(self.rng.integers(low=0, high=self.grid_h)), int(self.rng.integers(low=0, high=self.grid_w)))
        self.position = Pos(*position)
        tile = tile_view(self.image, self.patch_size, self.position)
        self.visited = set() if visited is None else visited
        if self.position in self.bbox_patches:
            self.visited.add(self.position)
        return tile, self._infos()

    def step(self, move: int):  # simple_env.py:347-376
        dy, dx = _DELTA[int(move)]
        y = min(max(self.position.y + dy, 0), self.grid_h - 1)
        x = min(max(self.position.x + dx, 0), self.grid_w - 1)
        self.position = Pos(y, x)
        if self.position in self.bbox_patches:
            self.visited.add(self.position)
        infos = self._infos()
        return tile_view(self.image, self.patch_size, self.position), infos

    # -- sample buffers -------------------------------------------------------------------
    def _blank_sample(self, max_ep_len: int) -> Dict[str, torch.Tensor]:
        # simple_env.py:378-441
        p, n = self.patch_size, len(self.raw_bboxes)
        sample = {
            "patches": torch.zeros((max_ep_len, self.n_channels, p, p), dtype=torch.float),
            "current_actions": torch.zeros((max_ep_len,), dtype=torch.long),
            "next_actions": torch.zeros((max_ep_len,), dtype=torch.long),
            "positions": torch.zeros((max_ep_len, 2), dtype=torch.long),
            "masks": torch.zeros((max_ep_len,), dtype=torch.float),
            "labels": torch.zeros((max_ep_len,), dtype=torch.long),
            "local_bboxes": torch.zeros((max_ep_len, n, 6)),
        }
        det_positions: Set[Pos] = set()
        for box in self.raw_bboxes:
            for q in self.patches_of_box(box):
                det_positions.add(q)
        empties = [Pos(y, x) for y, x in product(range(self.grid_h), range(self.grid_w)) if Pos(y, x) not in det_positions]
        if empties:
            det_positions.add(empties[self.rng.choice(len(empties))])
        tiles = [tile_view(self.image, p, q) for q in det_positions]
        boxes = [self.local_boxes(q) for q in det_positions]
        if not tiles:
            tiles.append(torch.zeros((self.n_channels, p, p), dtype=torch.float))
            boxes.append(torch.zeros((n, 6), dtype=torch.float))
        sample["patches_yolox"] = torch.stack(tiles)
        sample["bboxes_yolox"] = torch.stack(boxes)
        sample["_det_positions"] = list(det_positions)  # oracle-only: order of patches_yolox
        return sample

    @staticmethod
    def _record(sample, index: int, action: int, tile, infos, best_action: int):
        # simple_env.py:443-479 -- buffers double when full.
        if sample["patches"].shape[0] <= index:
            for key in list(sample):
                if key in ("patches_yolox", "bboxes_yolox", "_det_positions"):
                    continue
                sample[key] = torch.cat([sample[key], torch.zeros_like(sample[key])], dim=0)
        sample["patches"][index] = tile
        sample["current_actions"][index] = action
        sample["next_actions"][index] = best_action
        sample["positions"][index, 0] = infos["position"].y
        sample["positions"][index, 1] = infos["position"].x
        sample["masks"][index] = 1.0
        sample["labels"][index] = int(infos["inside_bbox"])
        sample["local_bboxes"][index] = infos["local_bboxes"]

    def _no_stop(self, action: int) -> int:  # simple_env.py:715-718
        if action == STOP:
            return int(self.rng.choice(8))  # == rng.choice(MOVES): one bounded draw over 8 items
        return action

    # -- keypoints -------------------------------------------------------------------------
    def _uniform_keypoint(self) -> Pos:  # simple_env.py:666-682 (y first)
        y = int(self.rng.integers(0, self.grid_h))
        x = int(self.rng.integers(0, self.grid_w))
        return Pos(y, x)

    def _binomial_keypoint(self, target: Pos) -> Pos:  # simple_env.py:684-713 (x first)
        dx = int(self.rng.binomial(self.grid_w, 0.5)) - self.grid_w // 2
        dy = int(self.rng.binomial(self.grid_h, 0.5)) - self.grid_h // 2
        return Pos((target[0] + dy) % self.grid_h, (target[1] + dx) % self.grid_w)

    def keypoint_order(self) -> List[Pos]:
        """simple_env.py:590-629 -- greedy L1-nearest ordering of the not-yet-visited box
        patches; ties are broken with python's global ``random.choice`` over the candidates
        in set-iteration order; if nothing is left, one uniform keypoint."""
        todo: Set[Pos] = set()
        for box in self.raw_bboxes:
            todo |= self.patches_of_box(box)
        for q in self.visited:
            todo.remove(q)
        here, order = self.position, []
        while todo:
            best, nearest = float("+inf"), []
            for q in todo:
                d = abs(q.x - here.x) + abs(q.y - here.y)
                if d < best:
                    best, nearest = d, []
                if d == best:
                    nearest.append(q)
            pick = random.choice(nearest)
            order.append(pick)
            todo.remove(pick)
            here = pick
        if not order:
            order.append(self._uniform_keypoint())
        return order

    def _walk(self, sample, to_visit: Pos, true_target: Pos):
        """simple_env.py:631-664 -- straight-line walk to ``to_visit``; the recorded best
        action points at ``true_target`` (a random move replaces STOP)."""
        self.reset(self.position)
        index = int(sample["masks"].long().sum().item())
        while self.position != to_visit:
            action = heading(self.position, to_visit)
            tile, infos = self.step(action)
            best = self._no_stop(heading(self.position, true_target))
            self.reset(self.position)
            self._record(sample, index, action, tile, infos, best)
            index += 1

    def generate_sample(
        self,
        max_ep_len: int,
        min_keypoints: int,
        max_keypoints: int,
        binomial_keypoints: bool = False,
        position: Optional[Pos] = None,
        visited: Optional[Set[Pos]] = None,
    ) -> Dict[str, torch.Tensor]:
        """simple_env.py:481-588."""
        sample = self._blank_sample(max_ep_len)
        tile, infos = self.reset(position, visited)
        self._record(sample, 0, LEFT, tile, infos, LEFT)
        keypoints = self.keypoint_order()
        n_random = self.rng.integers(min_keypoints, max_keypoints + 1)
        slots = sorted(self.rng.integers(0, len(keypoints), size=n_random), reverse=True)
        for k, keypoint in enumerate(keypoints):
            last = int(sample["masks"].long().sum().item()) - 1
            sample["next_actions"][last] = self._no_stop(heading(self.position, keypoint))
            while k in slots:
                detour = self._binomial_keypoint(keypoint) if binomial_keypoints else self._uniform_keypoint()
                self._walk(sample, detour, keypoint)
                slots.remove(k)
            self._walk(sample, keypoint, keypoint)
        ep_len = int(sample["masks"].long().sum().item())
        if ep_len > max_ep_len:  # keep the LAST max_ep_len steps (simple_env.py:573-584)
            for key in list(sample):
                if key not in ("patches_yolox", "bboxes_yolox", "_det_positions"):
                    sample[key] = sample[key][ep_len - max_ep_len : ep_len]
        assert sample["patches"].shape[0] == max_ep_len
        sample["_ep_len"] = ep_len
        return sample


def collate_oracle(batch: List[Dict[str, torch.Tensor]]) -> Dict[str, torch.Tensor]:
    """simple_env.py:720-763 -- pad the box axis to the batch maximum, stack the per-step
    tensors, concatenate the detection patches."""
    batch = [{k: v for k, v in s.items() if not k.startswith("_")} for s in batch]
    n_max = max(s["local_bboxes"].shape[1] for s in batch)

    def widen(t: torch.Tensor) -> torch.Tensor:
        pad = torch.zeros((t.shape[0], n_max - t.shape[1], t.shape[2]), dtype=torch.float32)
        return torch.cat((t, pad), dim=1)

    det_tiles = [s.pop("patches_yolox") for s in batch]
    det_boxes = [widen(s.pop("bboxes_yolox")) for s in batch]
    out = {}
    for key in batch[0]:
        vals = [widen(s[key]) if key == "local_bboxes" else s[key] for s in batch]
        out[key] = torch.stack(vals)
    out["patches_yolox"] = torch.cat(det_tiles)
    out["bboxes_yolox"] = torch.cat(det_boxes)
    return out


def generate_trajectories_oracle(
    images: List[torch.Tensor],
    bboxes: List[List],
    class_ids: List[int],
    patch_size: int,
    max_seq_len: int,
    min_keypoints: int,
    max_keypoints: int,
    binomial_keypoints: bool,
    position: Optional[Pos] = None,
    seeds: Optional[List[int]] = None,
) -> Dict[str, torch.Tensor]:
    """The batched entry point (reference: ``src/supervised.py:95-136``): one env per image,
    serial loop, collate.  ``seeds`` (one per image) makes it reproducible; the reference
    constructs its envs unseeded."""
    samples = []
    for i, image in enumerate(images):
        env = TrajectoryOracle(image, patch_size, bboxes[i], None if seeds is None else seeds[i])
        s = env.generate_sample(max_seq_len, min_keypoints, max_keypoints, binomial_keypoints, position)
        s["class_id"] = torch.tensor(class_ids[i], dtype=torch.long)
        samples.append(s)
    return collate_oracle(samples)
