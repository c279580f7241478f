"""Value types that cross the env API (reference: ``src/utils.py:10-12,95-106``)."""
from typing import List, NamedTuple

import torch


class Position(NamedTuple):
    """Patch- or pixel-space coordinate, row first (y, x)."""

    y: int
    x: int


class BBox(NamedTuple):
    """Axis-aligned box given by its two corners (each a ``Position``)."""

    up_left: Position
    bottom_right: Position


def bboxes_to_tensor(bboxes: List[BBox]) -> torch.Tensor:
    """``[N, 4]`` tensor in x1, y1, x2, y2 order (note: x first, unlike ``Position``)."""
    rows = [(b.up_left.x, b.up_left.y, b.bottom_right.x, b.bottom_right.y) for b in bboxes]
    return torch.tensor(rows)
