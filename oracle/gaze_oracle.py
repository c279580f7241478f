"""Oracle for the batched RL gaze environment (reference: ``src/env/general_env.py``).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  State lives in numpy arrays, crops are
taken with torch-CPU slicing exactly like the reference does (a per-episode slice followed
by one stack), so that timing this class on the host cores is a fair "port" baseline.

Every method cites the reference lines it restates.  Known quirks that are preserved on
purpose are listed in SURVEY.md appendix A.
"""
from typing import List, Optional, Tuple

import numpy as np
import torch

# (dy, dx) per action code -- src/env/common.py:17-27.
_DELTAS = np.array(
    [(0, -1), (0, 1), (-1, 0), (1, 0), (-1, -1), (-1, 1), (1, -1), (1, 1), (0, 0)],
    dtype=np.int64,
)
STOP = 8


def bbox_patch_mask_closed_form(bboxes: np.ndarray, height: int, width: int, patch: int) -> np.ndarray:
    """Patch x bbox containment under the any-pixel rule, without rasterising.

    general_env.py:360-379: boxes are rasterised with inclusive x2/y2 after clamping to the
    image (kornia ``xyxy_plus`` -> ``to_mask``), OR-ed over the N boxes, then max-pooled with
    a PxP window.  A patch is therefore set iff it intersects the clamped half-open pixel
    rectangle [x1c, x2c) x [y1c, y2c) of some box.
    """
    b, n, _ = bboxes.shape
    rows, cols = height // patch, width // patch
    out = np.zeros((b, rows, cols), dtype=bool)
    for i in range(b):
        for j in range(n):
            x1, y1, x2, y2 = (int(v) for v in bboxes[i, j])
            x1c, x2c = min(max(x1, 0), width), min(max(x2 + 1, 0), width)
            y1c, y2c = min(max(y1, 0), height), min(max(y2 + 1, 0), height)
            if x1c >= x2c or y1c >= y2c:
                continue
            out[i, y1c // patch : (y2c - 1) // patch + 1, x1c // patch : (x2c - 1) // patch + 1] = True
    return out


def bbox_patch_mask_raster(bboxes: np.ndarray, height: int, width: int, patch: int) -> np.ndarray:
    """Same result as the closed form, but following the reference's data flow
    (general_env.py:373-379): a ``[B, N, H, W]`` float mask, max over N, ``max_pool2d``.
    Used to cross-check the closed form and as the honest cost of the port baseline."""
    b, n, _ = bboxes.shape
    mask = torch.zeros((b, n, height, width), dtype=torch.float32)
    for i in range(b):
        for j in range(n):
            x1, y1, x2, y2 = (int(v) for v in bboxes[i, j])
            x1c, x2c = min(max(x1, 0), width), min(max(x2 + 1, 0), width)
            y1c, y2c = min(max(y1, 0), height), min(max(y2 + 1, 0), height)
            mask[i, j, y1c:y2c, x1c:x2c] = 1
    merged = mask.long().max(dim=1).values
    pooled = torch.nn.functional.max_pool2d(merged.float(), patch)
    return pooled.bool().numpy()


def split_boxes_per_patch(bboxes: np.ndarray, rows: int, cols: int, patch: int) -> Tuple[np.ndarray, np.ndarray]:
    """Per-patch local boxes (general_env.py:381-504 ``parse_bboxes``).

    The reference places a box in the patch holding its top-left corner, clamps the far
    corner to ``patch - 1`` and recurses right / down / diagonally for the remainder.  The
    recursion is restated iteratively with an explicit work stack, in the same visiting
    order (so that overlapping writes land identically).  Boxes whose corner lies outside
    the grid raise ``IndexError`` like the reference does.
    """
    b, n, _ = bboxes.shape
    local = np.zeros((b, rows, cols, n, 4), dtype=np.int64)
    present = np.zeros((b, rows, cols, n), dtype=bool)

    def place(i: int, k: int, box):
        bx1, by1, bx2, by2 = box
        lx1, ly1 = bx1 % patch, by1 % patch
        lx2, ly2 = lx1 + (bx2 - bx1), ly1 + (by2 - by1)
        px, py = bx1 // patch, by1 // patch
        cx2, cy2 = min(lx2, patch - 1), min(ly2, patch - 1)
        if not (-rows <= py < rows and -cols <= px < cols):
            raise IndexError(f"bbox {box} falls outside the {rows}x{cols} patch grid")
        local[i, py, px, k] = (lx1, ly1, cx2, cy2)
        present[i, py, px, k] = True
        over_x, over_y = lx2 - cx2 > 0, ly2 - cy2 > 0
        if over_x:
            place(i, k, ((px + 1) * patch, by1, bx2, py * patch + cy2))
        if over_y:
            place(i, k, (bx1, (py + 1) * patch, px * patch + cx2, by2))
        if over_x and over_y:
            place(i, k, ((px + 1) * patch, (py + 1) * patch, bx2, by2))

    for i in range(b):
        for k in range(n):
            # `.int()` in the reference truncates float boxes (general_env.py:495).
            place(i, k, tuple(int(v) for v in bboxes[i, k]))
    return local, present


class GazeOracle:
    """Restatement of ``NeedleGeneralEnv`` (general_env.py:14-573) for ``n_glimps_levels == 1``."""

    def __init__(
        self,
        images: torch.Tensor,
        bboxes,
        patch_size: int,
        max_ep_len: int,
        n_glimps_levels: int = 1,
        stop_enabled: bool = False,
        raster_masks: bool = False,
    ):
        """``images`` may be a bare ``(B, C, H, W)`` shape instead of a tensor: the oracle then tracks every
        integer / float output of the env but returns ``None`` for the crops -- full-size batches (1024 LARD
        images are 74 GB as float32) are checked that way, with a sub-batch replayed with pixels."""
        bboxes = np.asarray(bboxes, dtype=np.int64)
        self.has_pixels = isinstance(images, torch.Tensor)
        shape = tuple(images.shape) if self.has_pixels else tuple(int(v) for v in images)
        assert shape[0] == bboxes.shape[0]
        assert len(shape) == 4
        assert n_glimps_levels == 1, "the oracle covers the single-level env only"
        self.batch_size, self.n_channels, self.height, self.width = shape
        assert self.height % patch_size == 0 and self.width % patch_size == 0
        self.patch_size, self.max_ep_len = patch_size, max_ep_len
        self.stop_enabled = stop_enabled
        self.rows, self.cols = self.height // patch_size, self.width // patch_size
        maker = bbox_patch_mask_raster if raster_masks else bbox_patch_mask_closed_form
        self.bbox_masks = maker(bboxes, self.height, self.width, patch_size)
        self.bboxes = bboxes
        self.images = images  # level 0 of init_glimps_images (general_env.py:84-115)
        self._clear()

    # -- state -----------------------------------------------------------------------
    def _clear(self):  # general_env.py:117-142
        b = self.batch_size
        self.positions = np.zeros((b, 2), dtype=np.int64)
        self.visited = np.zeros((b, self.rows, self.cols), dtype=bool)
        self.steps = np.zeros((b,), dtype=np.int64)
        self.has_stopped = np.zeros((b,), dtype=bool)

    def _here(self) -> np.ndarray:  # general_env.py:248-283 (one-hot of the current patch)
        hot = np.zeros_like(self.visited)
        hot[np.arange(self.batch_size), self.positions[:, 0], self.positions[:, 1]] = True
        return hot

    # -- API -------------------------------------------------------------------------
    def reset(self, positions: Optional[np.ndarray] = None):  # general_env.py:144-170
        self._clear()
        if positions is None:
            # CPU default generator, rows first then columns (general_env.py:158-163).
            ys = torch.randint(low=0, high=self.rows, size=(self.batch_size,))
            xs = torch.randint(low=0, high=self.cols, size=(self.batch_size,))
            self.positions = np.stack([ys.numpy(), xs.numpy()], axis=1).astype(np.int64)
        else:
            self.positions = np.asarray(positions, dtype=np.int64).copy()
        self.visited |= self._here()
        return self.crops(), {"positions": self.positions.copy()}

    def step(self, actions):  # general_env.py:172-207
        actions = np.asarray(actions, dtype=np.int64)
        # move + clamp, sticky stop flag (general_env.py:209-233)
        moved = self.positions + _DELTAS[actions]
        moved[:, 0] = np.clip(moved[:, 0], 0, self.rows - 1)
        moved[:, 1] = np.clip(moved[:, 1], 0, self.cols - 1)
        self.positions = moved
        self.has_stopped |= actions == STOP
        rewards = self.rewards()  # uses `visited` BEFORE the new patch is marked
        self.visited |= self._here()
        self.steps += 1
        truncated = self.steps >= self.max_ep_len
        return self.crops(), rewards, self.terminated(), truncated, {"positions": self.positions.copy()}

    def terminated(self) -> np.ndarray:  # general_env.py:235-246
        if self.stop_enabled:
            return self.has_stopped.copy()
        missing = (self.bbox_masks & self.visited) != self.bbox_masks
        return missing.sum(axis=(1, 2)) == 0

    def rewards(self) -> np.ndarray:  # general_env.py:321-358
        ids = np.arange(self.batch_size)
        ys, xs = self.positions[:, 0], self.positions[:, 1]
        fresh = self.bbox_masks[ids, ys, xs] & ~self.visited[ids, ys, xs]
        cost = np.float32(-1 / self.max_ep_len)
        total = fresh.astype(np.float32) + cost  # fp32 add #1
        if self.stop_enabled:
            found = (self.visited & self.bbox_masks).sum(axis=(1, 2)).astype(np.int64)
            every = self.bbox_masks.sum(axis=(1, 2)).astype(np.int64)
            stop_eval = np.where(found == every, found, found - every) * self.has_stopped
            total = total + stop_eval.astype(np.float32)  # fp32 add #2
        return total.astype(np.float32)

    def crops(self) -> Optional[torch.Tensor]:  # general_env.py:285-306
        if not self.has_pixels:
            return None
        p = self.patch_size
        tiles = [
            self.images[i, :, y * p : (y + 1) * p, x * p : (x + 1) * p]
            for i, (y, x) in enumerate(self.positions.tolist())
        ]
        return torch.stack(tiles).unsqueeze(1)  # [B, G=1, C, P, P]

    def prop_patches_found(self) -> np.ndarray:  # general_env.py:308-315
        count = (self.bbox_masks & self.visited).sum(axis=(1, 2)).astype(np.int64)
        tot = self.bbox_masks.sum(axis=(1, 2)).astype(np.int64)
        tot[tot == 0] = 1
        return (count.astype(np.float32) / tot.astype(np.float32)).astype(np.float32)

    def prop_bboxes_found(self) -> np.ndarray:  # general_env.py:317-319
        return (self.prop_patches_found() > 0).astype(np.float32)

    # -- detection side (general_env.py:506-573) --------------------------------------
    def detection_targets(self) -> List[np.ndarray]:
        """Global-coordinate split boxes per image, class id 0 in column 0; all-zero
        entries are skipped (general_env.py:548-573)."""
        local, _ = split_boxes_per_patch(self.bboxes, self.rows, self.cols, self.patch_size)
        out = []
        for i in range(self.batch_size):
            rows_i = []
            for y in range(self.rows):
                for x in range(self.cols):
                    for k in range(local.shape[3]):
                        box = local[i, y, x, k]
                        if np.abs(box).sum() == 0:
                            continue
                        off = np.array([x, y, x, y], dtype=np.int64) * self.patch_size
                        rows_i.append(np.concatenate([[0], box + off]))
            out.append(np.stack(rows_i).astype(np.int64))
        return out

    def detection_batch(self, sample_neg: int = 1):
        """Every (patch, box) hit plus ``sample_neg`` random misses per image
        (general_env.py:506-546), including the ``squeeze(-1)`` quirk: with a single box per
        image the box axis disappears and entries are patches; with several boxes each
        (patch, box) pair is its own entry.  Consumes ``torch.randperm`` once per image."""
        local, present = split_boxes_per_patch(self.bboxes, self.rows, self.cols, self.patch_size)
        if present.shape[-1] == 1:
            present = present[..., 0]
        p = self.patch_size
        tiles, targets = [], []
        for i in range(self.batch_size):
            hits = np.nonzero(present[i])
            misses = np.nonzero(~present[i])
            pick = torch.randperm(len(misses[0]))[:sample_neg].numpy()
            misses = tuple(m[pick] for m in misses)
            ids = tuple(np.concatenate([h, m]) for h, m in zip(hits, misses))
            for j in range(len(ids[0])):
                r, c = int(ids[0][j]), int(ids[1][j])
                tiles.append(self.images[i, :, r * p : (r + 1) * p, c * p : (c + 1) * p])
                targets.append(np.concatenate([np.zeros((local.shape[3], 1), np.int64), local[i, r, c]], axis=1))
        return torch.stack(tiles), np.stack(targets)


def returns_oracle(rewards: torch.Tensor, masks: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Per-episode returns of a rollout (reference: ``src/reinforce.py:186-202``).

    ``rewards`` is ``[B, T']`` fp32, ``masks`` is ``[B, T'+1]`` bool with column 0 True and
    column t+1 = ``~terminated`` after step t.  The reward of the terminating step counts,
    later ones do not: ``logit_masks[:, 0] = True``, ``logit_masks[:, t] = masks[:, t]``.
    Returns are the reverse cumulative sum of the masked rewards; torch's CPU ``cumsum``
    accumulates fp32 inputs in fp64 and rounds each output, which this follows by calling
    the same torch op.
    """
    logit_masks = masks[:, :-1].clone()
    logit_masks[:, 0] = True
    masked = rewards * logit_masks
    returns = torch.flip(torch.cumsum(torch.flip(masked, dims=(1,)), dim=1), dims=(1,))
    return returns, logit_masks


def returns_oracle_numpy(rewards: np.ndarray, logit_masks: np.ndarray) -> np.ndarray:
    """Explicit arithmetic of :func:`returns_oracle`: fp32 product, fp64 running sum from the
    last column, one fp32 rounding per output element."""
    b, t = rewards.shape
    out = np.zeros((b, t), dtype=np.float32)
    for i in range(b):
        acc = np.float64(0.0)
        for s in range(t - 1, -1, -1):
            term = np.float32(rewards[i, s]) * np.float32(1.0 if logit_masks[i, s] else 0.0)
            acc = acc + np.float64(term)
            out[i, s] = np.float32(acc)
    return out


def translate_oracle(images: torch.Tensor, shifts_xy) -> torch.Tensor:
    """Integer translation with zero fill, per image: ``out[.., y, x] = in[.., y - ty, x - tx]``.

    This is what the reference's dataset does for ``--augment-translate`` (dataset.py:207-214):
    ``torchvision.transforms.functional.affine(image, angle=0, translate=[tx, ty], scale=1.0,
    shear=0.0, fill=0.0)`` with its default nearest interpolation (checked against torchvision in
    tests/test_oracle_cpu.py).  ``images`` is [B, C, H, W]; ``shifts_xy`` is [B, 2] = (tx, ty)."""
    out = torch.zeros_like(images)
    _, _, h, w = images.shape
    for i, (tx, ty) in enumerate(np.asarray(shifts_xy).tolist()):
        ys0, ys1 = max(0, -ty), min(h, h - ty)  # source rows that stay inside
        xs0, xs1 = max(0, -tx), min(w, w - tx)
        if ys0 < ys1 and xs0 < xs1:
            out[i, :, ys0 + ty : ys1 + ty, xs0 + tx : xs1 + tx] = images[i, :, ys0:ys1, xs0:xs1]
    return out
