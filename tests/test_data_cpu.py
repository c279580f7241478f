"""Dataset -> env hand-off helpers (jolineedle_b200/data.py) against the reference's collate rules
(dataset.py:298-347), restated here with the same torch calls on float32 images."""
import numpy as np
import torch

from jolineedle_b200.data import pinned_u8_collate, pinned_u8_list_collate, to_uint8_chw
from jolineedle_b200.utils import BBox, Position, bboxes_to_tensor


def padded_collate_restated(batch, patch_size):
    images = [s["image"] for s in batch]
    max_h, max_w = max(i.shape[1] for i in images), max(i.shape[2] for i in images)
    max_bbox = max(len(s["bboxes"]) for s in batch)
    dh, dw = patch_size - max_h % patch_size, patch_size - max_w % patch_size
    fh, fw = max_h + (dh if dh != patch_size else 0), max_w + (dw if dw != patch_size else 0)
    out_i, out_b = [], []
    for s in batch:
        h, w = s["image"].shape[1:]
        out_i.append(torch.nn.functional.pad(s["image"], (0, fw - w, 0, fh - h), mode="constant", value=0))
        bbox = bboxes_to_tensor(s["bboxes"]) if len(s["bboxes"]) else torch.zeros((0, 4), dtype=torch.long)
        out_b.append(torch.nn.functional.pad(bbox, (0, 0, 0, max_bbox - bbox.shape[0]), mode="constant", value=0))
    return {"image": torch.stack(out_i), "bboxes": torch.stack(out_b), "class_id": torch.tensor([s["class_id"] for s in batch])}


def samples(rng, n, sizes):
    out = []
    for i in range(n):
        h, w = sizes[i % len(sizes)]
        hwc = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
        boxes = [BBox(Position(int(rng.integers(0, h // 2)), int(rng.integers(0, w // 2))), Position(h - 1, w - 1))
                 for _ in range(int(rng.integers(1, 4)))]
        out.append({"hwc": hwc, "bboxes": boxes, "class_id": i % 3})
    return out


def test_every_byte_survives_the_float_round_trip():
    b = torch.arange(256, dtype=torch.uint8).view(1, 16, 16)
    as_float = b.float() / 255  # ToTensor
    assert torch.equal(to_uint8_chw(as_float), b)
    hwc = np.arange(2 * 5 * 3, dtype=np.uint8).reshape(2, 5, 3)
    assert torch.equal(to_uint8_chw(hwc), torch.from_numpy(hwc).permute(2, 0, 1))


def test_padded_collate_matches_the_reference_rules():
    rng = np.random.default_rng(0)
    raw = samples(rng, 5, [(60, 100), (64, 96), (33, 47)])
    P = 32
    ref = padded_collate_restated([{"image": torch.from_numpy(s["hwc"]).permute(2, 0, 1).float() / 255, **s} for s in raw], P)
    for feed in ("hwc uint8", "float chw"):
        batch = [{"image": s["hwc"] if feed == "hwc uint8" else torch.from_numpy(s["hwc"]).permute(2, 0, 1).float() / 255,
                  "bboxes": s["bboxes"], "class_id": s["class_id"]} for s in raw]
        got = pinned_u8_collate(batch, P, pin=False)
        assert got["image"].dtype == torch.uint8 and got["image"].shape == ref["image"].shape
        assert got["image"].shape[2] % P == 0 and got["image"].shape[3] % P == 0
        assert torch.equal(got["image"].float() / 255, ref["image"]), feed  # normalize-on-gather sees the same pixels
        assert torch.equal(got["bboxes"], ref["bboxes"]) and torch.equal(got["class_id"], ref["class_id"])


def test_list_collate_keeps_lists():
    rng = np.random.default_rng(1)
    raw = samples(rng, 3, [(64, 96), (32, 64)])
    got = pinned_u8_list_collate([{"image": s["hwc"], "bboxes": s["bboxes"], "class_id": s["class_id"]} for s in raw], pin=False)
    assert isinstance(got["image"], list) and [tuple(i.shape) for i in got["image"]] == [(3, 64, 96), (3, 32, 64), (3, 64, 96)]
    assert all(i.dtype == torch.uint8 and i.is_contiguous() for i in got["image"])
    assert got["bboxes"][1] == raw[1]["bboxes"] and got["class_id"] == [0, 1, 2]
