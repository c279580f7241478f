#!/usr/bin/env python
"""Host + device time of the detection side at BASELINE cfg-3 shape (1024 images of 2240x2688, P = 448):
``get_detection_batch`` (K0 split table, per-image ``torch.randperm`` in the reference's order, one K1 gather of
every positive patch + one negative per image) and ``get_detection_targets``.

    python tools/microbench_detection.py [batch]

Round 2, one B200: batch 10.8 ms per call for 7862 patches (23.7 GB of crops; the per-image python loop of the
round-1 version took 33.5 ms), targets 0.8 ms (round 1: 50.8 ms, one masked select = one sync per image)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from jolineedle_b200.env.general_env import NeedleGeneralEnv  # noqa: E402


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    dev = torch.device("cuda", 0)
    wl = bench.ReinforceWorkload(batch, 0, dev, "u8")
    wl.to_device()
    env = NeedleGeneralEnv(wl.images, wl.boxes_dev, 448, 20, 1, True, normalize=True)
    for name, fn in (("get_detection_batch", env.get_detection_batch), ("get_detection_targets", env.get_detection_targets)):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            out = fn()
        torch.cuda.synchronize()
        what = tuple(out[0].shape) if name.endswith("batch") else f"{len(out)} images"
        print(f"{name}: {(time.perf_counter() - t0) / 5 * 1e3:.2f} ms per call ({what})")


if __name__ == "__main__":
    main()
