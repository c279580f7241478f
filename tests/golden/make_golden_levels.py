#!/usr/bin/env python
"""Golden fixture for the multi-level glimpse pyramid (n_glimps_levels > 1, general_env.py:84-115), produced by
the UNMODIFIED reference env on the CPU.  Build container only:  python tests/golden/make_golden_levels.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _ref_loader import load_reference  # noqa: E402
from make_golden import synth_u8, to_f32  # noqa: E402

ge, _, _, _ = load_reference()


def main():
    fx = {}
    for name, levels in (("lv2", 2), ("lv3", 3)):
        u8 = synth_u8(3, 3, 80, 96, salt=40 + levels)
        boxes = np.array([[[5, 5, 30, 30]], [[40, 20, 70, 60]], [[0, 0, 0, 0]]], dtype=np.int64)
        env = ge.NeedleGeneralEnv(to_f32(u8), torch.from_numpy(boxes), 16, 6, levels, True)
        start = np.array([[0, 0], [2, 3], [4, 5]], dtype=np.int64)
        actions = np.array([[1, 3, 0], [7, 2, 4], [8, 8, 1]], dtype=np.int64)
        patches = [env.reset(torch.from_numpy(start))[0].numpy()]
        rewards = []
        for a in actions:
            out = env.step(torch.from_numpy(a))
            patches.append(out[0].numpy())
            rewards.append(out[1].numpy())
        fx.update({f"{name}/u8": u8, f"{name}/boxes": boxes, f"{name}/levels": levels, f"{name}/start": start,
                   f"{name}/actions": actions, f"{name}/patches": np.stack(patches), f"{name}/rewards": np.stack(rewards),
                   f"{name}/images": env.images.numpy()})
    # uint8 images (what the reference's docstring promises, general_env.py:28): torchvision resizes them through
    # float32 + torch.round, level after level
    u8 = synth_u8(3, 3, 80, 96, salt=47)
    boxes = np.array([[[5, 5, 30, 30]], [[40, 20, 70, 60]], [[0, 0, 0, 0]]], dtype=np.int64)
    env = ge.NeedleGeneralEnv(torch.from_numpy(u8), torch.from_numpy(boxes), 16, 6, 3, True)
    start = np.array([[1, 1], [2, 3], [4, 0]], dtype=np.int64)
    actions = np.array([[1, 3, 0], [7, 2, 4]], dtype=np.int64)
    patches = [env.reset(torch.from_numpy(start))[0].numpy()]
    rewards = []
    for a in actions:
        out = env.step(torch.from_numpy(a))
        patches.append(out[0].numpy())
        rewards.append(out[1].numpy())
    assert env.images.dtype == torch.uint8
    fx.update({"lv3u8/u8": u8, "lv3u8/boxes": boxes, "lv3u8/levels": 3, "lv3u8/start": start, "lv3u8/actions": actions,
               "lv3u8/patches": np.stack(patches), "lv3u8/rewards": np.stack(rewards), "lv3u8/images": env.images.numpy()})
    np.savez_compressed(os.path.join(HERE, "glimpse_levels.npz"), **fx)
    print("glimpse_levels.npz", os.path.getsize(os.path.join(HERE, "glimpse_levels.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
