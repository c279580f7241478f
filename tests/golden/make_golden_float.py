#!/usr/bin/env python
"""Golden fixture for boxes that are NOT whole pixels (build container only, needs /root/reference).

The reference's dataset rescales small images and their boxes (dataset.py:258-270), so ``NeedleSimpleEnv`` can
receive float coordinates and keeps them as python floats: the 5 % area rule, the centre patch and the local
boxes are all evaluated on the un-truncated values.  This runs the UNMODIFIED reference on such boxes and stores
its samples in ``simple_env_float.npz``.

    python tests/golden/make_golden_float.py
"""
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

from make_golden import SAMPLE_KEYS, se, synth_u8, to_f32, ut  # noqa: E402


def main():
    rng = np.random.default_rng(77)
    fx, names = {}, []
    # boxes scaled by a resize factor (integers x 1.37 and friends), including values that sit right at the
    # 5 % area threshold and at patch edges, mixed with whole-pixel boxes
    for idx in range(12):
        P = int(rng.choice([16, 32]))
        gh, gw = int(rng.integers(3, 8)), int(rng.integers(3, 8))
        h, w = gh * P, gw * P
        T = int(rng.choice([6, 8, 12]))
        scale = float(rng.choice([1.37, 0.83, 1.5, 2.0 / 3.0]))
        raw = []
        for _ in range(int(rng.integers(1, 4))):
            bw, bh = (int(v) for v in rng.integers(2, 3 * P, size=2))
            x1, y1 = int(rng.integers(0, w)), int(rng.integers(0, h))
            box = [x1 * scale, y1 * scale, min((x1 + bw) * scale, w - 1), min((y1 + bh) * scale, h - 1)]
            if box[2] <= box[0] or box[3] <= box[1]:
                continue
            raw.append(box)
        if idx % 4 == 0:  # a sliver whose overlap with its patch is just above / below 5 % of P^2
            x0, y0 = P * int(rng.integers(0, gw - 1)), P * int(rng.integers(0, gh - 1))
            raw.append([x0 + 0.25, y0 + 0.5, x0 + 0.25 + 0.05 * P + (0.01 if idx % 8 else -0.01), y0 + 0.5 + P - 1])
        seed = 3000 + idx
        binomial = bool(idx % 2)
        kmin, kmax = ((0, 3), (1, 2), (0, 0))[idx % 3]
        u8 = synth_u8(1, 3, h, w, salt=50 + idx)[0]
        random.seed(seed * 7 + 1)
        boxes = [ut.BBox(ut.Position(y1, x1), ut.Position(y2, x2)) for (x1, y1, x2, y2) in raw]
        env = se.NeedleSimpleEnv(to_f32(u8), P, boxes, seed=seed)
        s = env.generate_sample(T, kmin, kmax, binomial_keypoints=binomial, position=None)
        name = f"f{idx:02d}"
        fx[f"{name}/u8"] = u8
        fx[f"{name}/raw_boxes"] = np.array(raw, dtype=np.float64).reshape(-1, 4)
        fx[f"{name}/cfg"] = np.array([P, T, kmin, kmax, int(binomial), seed, -1, -1], dtype=np.int64)
        fx[f"{name}/bbox_patches"] = np.array(sorted(env.bbox_patches), dtype=np.int64).reshape(-1, 2)
        for k in SAMPLE_KEYS:
            fx[f"{name}/{k}"] = s[k].numpy()
        names.append(name)
    fx["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "simple_env_float.npz"), **fx)
    print("wrote simple_env_float.npz:", len(names), "cases")


if __name__ == "__main__":
    main()
