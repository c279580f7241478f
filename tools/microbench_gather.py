#!/usr/bin/env python
"""BASELINE cfg 5: gather microbench -- patch 128..1024 x batch 64..8192, uint8 vs fp32 source,
per engine; prints achieved algorithmic GB/s (bytes = C*P^2*(s_in+s_out) per tile) as JSON lines.

Working set per launch exceeds L2 (>= 1 GB of crops where memory allows) and source tiles do not
repeat within a launch where the image pool allows it."""
import argparse
import json
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jolineedle_b200.gather import ImageSet  # noqa: E402


def bench_one(P, n, dtype, normalize, focus, engine, iters=10, pool_bytes=6 << 30, translate=False):
    elem = 1 if dtype == torch.uint8 else 4
    gh = gw = max(2, 2048 // P)
    img_bytes = 3 * gh * P * gw * P * elem
    n_img = max(1, min(n, pool_bytes // img_bytes))
    g = torch.Generator(device="cuda").manual_seed(0)
    u8 = torch.randint(0, 256, (n_img, 3, gh * P, gw * P), dtype=torch.uint8, device="cuda", generator=g)
    images = u8 if dtype == torch.uint8 else u8.float().div_(255)
    del u8
    s = ImageSet(images, P)
    # distinct tiles as far as the pool allows
    idx = torch.arange(n, device="cuda")
    src = (idx % n_img).to(torch.int32)
    cell = (idx // n_img) % (gh * gw)
    pos = torch.stack([cell // gw, cell % gw], 1).contiguous()
    out_elem = 4 if (normalize or dtype == torch.float32) else 1
    shifts = None
    if translate:  # arbitrary integer (ty, tx) per image within half a patch (zero fill at the borders)
        shifts = torch.randint(-P // 2, P // 2 + 1, (n_img, 2), generator=torch.Generator().manual_seed(1)).int().cuda()
    out = torch.empty(s.out_shape(n, focus), dtype=s.out_dtype(normalize), device="cuda")
    for _ in range(3):
        s.gather(pos, src_index=src, out=out, normalize=normalize, focus=focus, engine=engine, shifts=shifts)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        s.gather(pos, src_index=src, out=out, normalize=normalize, focus=focus, engine=engine, shifts=shifts)
        ev[i + 1].record()
    torch.cuda.synchronize()
    times = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(iters))
    ms = times[len(times) // 2]
    bytes_ = n * 3 * P * P * (elem + out_elem)
    return {"P": P, "n": n, "src": "u8" if elem == 1 else "f32", "normalize": normalize, "focus": focus,
            "engine": engine, "translate": translate, "ms": round(ms, 4), "GBps": round(bytes_ / ms / 1e6, 1),
            "tiles_per_s": round(n / ms * 1e3), "out_MB": round(n * 3 * P * P * out_elem / 1e6)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--engines", default="tensor,bulk")
    ap.add_argument("--out", default=None)
    ap.add_argument("--patches", default="")
    ap.add_argument("--batches", default="")
    ap.add_argument("--modes", default="f32,u8", help="f32 = fp32 copy, u8 = uint8 normalised")
    ap.add_argument("--layouts", default="plain,focus")
    ap.add_argument("--label", default="")
    ap.add_argument("--translate", action="store_true", help="per-image integer translation folded into the gather")
    args = ap.parse_args()
    peak = 6465.2
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    combos = []
    patches = [448, 256] if args.quick else [128, 256, 448, 1024]
    if args.patches:
        patches = [int(v) for v in args.patches.split(",")]
    for P in patches:
        batches = [64, 512, 2048, 8192] if not args.quick else [256, 2048]
        if args.batches:
            batches = [int(v) for v in args.batches.split(",")]
        batches = [b for b in batches if b * 3 * P * P * 4 <= 24 << 30]
        modes = [(torch.float32, False)] * ("f32" in args.modes) + [(torch.uint8, True)] * ("u8" in args.modes)
        for n in batches:
            for dtype, normalize in modes:
                for focus in [False] * ("plain" in args.layouts) + [True] * ("focus" in args.layouts):
                    for engine in args.engines.split(","):
                        combos.append((P, n, dtype, normalize, focus, engine))
    lines = []
    for c in combos:
        try:
            r = bench_one(*c, translate=args.translate)
            r["frac_of_measured_peak"] = round(r["GBps"] / peak, 3)
            if args.label:
                r["label"] = args.label
        except Exception as e:  # keep sweeping
            r = {"P": c[0], "n": c[1], "engine": c[5], "error": repr(e)[:200]}
        print(json.dumps(r), flush=True)
        lines.append(r)
        torch.cuda.empty_cache()
    if args.out:
        with open(args.out, "a") as f:
            for r in lines:
                f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
