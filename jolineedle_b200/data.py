"""Dataset -> env hand-off helpers: uint8, page-locked batches.

The reference's dataset turns every image into float32 ``[0, 1]`` on the CPU (``ToTensor``, dataset.py:240),
zero-pads and stacks the batch (``padded_collate_fn``, dataset.py:307-347) and the trainer uploads it whole
(reinforce.py:313-315): 4 bytes per pixel cross PCIe for images of which an episode looks at a few patches.
The env in this package normalises on the fly (``normalize=True``: crops equal ``ToTensor`` then crop, bit for
bit) and can read page-locked host images in place (``zero_copy=True``), so the cheapest hand-off is a pinned
**uint8** batch.  The two collate functions below produce exactly that, with the reference's padding rules:

    loader = DataLoader(dataset, batch_size=B, collate_fn=partial(pinned_u8_collate, patch_size=P))
    for batch in loader:                                    # reinforce.py:302-326
        env = NeedleGeneralEnv(batch["image"], batch["bboxes"], P, T, 1, stop, device=dev,
                               normalize=True, zero_copy=True, history=True)

``dataset[i]["image"]`` may be what the reference yields (float32 CHW in [0, 1], converted back to bytes exactly:
``round(x * 255)``) or, cheaper, the raw uint8 image (HWC as loaded, or CHW) -- see :func:`to_uint8_chw`.
"""
from typing import Dict, List, Sequence

import numpy as np
import torch

from .utils import bboxes_to_tensor


def to_uint8_chw(image) -> torch.Tensor:
    """``[C, H, W]`` uint8 tensor of an image given as HWC / CHW uint8 (numpy or torch) or as the float32 CHW
    ``[0, 1]`` tensor ``ToTensor`` makes (dataset.py:240).  ``ToTensor`` computes ``byte / 255`` in float32 and
    ``round(x * 255)`` inverts that for every byte value, so a float image that came from 8-bit pixels loses
    nothing."""
    t = torch.from_numpy(np.ascontiguousarray(image)) if isinstance(image, np.ndarray) else image
    if t.dtype == torch.uint8:
        if t.dim() == 3 and t.shape[-1] in (1, 3, 4) and t.shape[0] not in (1, 3, 4):
            t = t.permute(2, 0, 1)  # HWC as cv2 / PIL hand it over
        return t
    if t.dtype == torch.float32:
        return (t * 255).round_().clamp_(0, 255).to(torch.uint8)
    raise ValueError(f"images must be uint8 or float32 in [0, 1], got {t.dtype}")


def _alloc(shape, pin: bool) -> torch.Tensor:
    return torch.empty(shape, dtype=torch.uint8, pin_memory=pin)


def pinned_u8_collate(batch: Sequence[Dict], patch_size: int, pin: bool = True) -> Dict:
    """``padded_collate_fn`` (dataset.py:307-347) for uint8 images: every image is zero-padded at the bottom /
    right to the largest height / width of the batch rounded up to a multiple of ``patch_size``, written
    straight into ONE page-locked ``[B, C, H, W]`` uint8 buffer (no per-image padded copies, no float32); boxes
    become a zero-padded int64 ``[B, N, 4]`` tensor (x1, y1, x2, y2), class ids an int64 ``[B]`` tensor.
    ``pin=False`` for DataLoader workers (no CUDA context there): let the loader's ``pin_memory=True`` pin it."""
    images = [to_uint8_chw(s["image"]) for s in batch]
    channels = images[0].shape[0]
    max_h = max(im.shape[1] for im in images)
    max_w = max(im.shape[2] for im in images)
    final_h = -(-max_h // patch_size) * patch_size
    final_w = -(-max_w // patch_size) * patch_size
    out = _alloc((len(images), channels, final_h, final_w), pin)
    for i, im in enumerate(images):
        h, w = im.shape[1:]
        out[i, :, :h, :w] = im
        if h < final_h:
            out[i, :, h:, :] = 0
        if w < final_w:
            out[i, :, :h, w:] = 0
    max_boxes = max(len(s["bboxes"]) for s in batch)
    boxes = torch.zeros((len(images), max_boxes, 4), dtype=torch.long)
    for i, s in enumerate(batch):
        if len(s["bboxes"]):
            boxes[i, : len(s["bboxes"])] = bboxes_to_tensor(s["bboxes"])
    return {"image": out, "bboxes": boxes, "class_id": torch.tensor([s["class_id"] for s in batch])}


def pinned_u8_list_collate(batch: Sequence[Dict], pin: bool = True) -> Dict:
    """``list_collate_fn`` (dataset.py:298-305) for the supervised trainer: lists, not stacks; every image a
    page-locked uint8 ``[C, H, W]`` tensor that ``generate_trajectories(..., device=dev, normalize=True)`` reads in
    place (only the glimpsed tiles cross PCIe)."""
    out: Dict[str, List] = {key: [s[key] for s in batch] for key in batch[0].keys()}
    images = []
    for im in out["image"]:
        im = to_uint8_chw(im)
        if pin and not im.is_pinned():
            im = _alloc(im.shape, True).copy_(im)
        elif not im.is_contiguous():
            im = im.contiguous()
        images.append(im)
    out["image"] = images
    return out
