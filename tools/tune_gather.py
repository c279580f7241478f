#!/usr/bin/env python
"""Sweep the gather pipeline shape (JN_GATHER_TUNE) per patch size / dtype / layout / engine and
print the best settings.  Images are built once per shape; each setting runs 3 warm-up + 7 timed
launches (median).  Output: gpurun_out/tune_sweep.jsonl (all) and a top-5 table per shape."""
import argparse, itertools, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jolineedle_b200.gather import ImageSet  # noqa: E402

PEAK = 6465.2


def time_gather(s, pos, src, out, normalize, focus, engine, iters=7, shifts=None):
    for _ in range(3):
        s.gather(pos, src_index=src, out=out, normalize=normalize, focus=focus, engine=engine, shifts=shifts)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        s.gather(pos, src_index=src, out=out, normalize=normalize, focus=focus, engine=engine, shifts=shifts)
        ev[i + 1].record()
    torch.cuda.synchronize()
    t = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(iters))
    return t[len(t) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--patches", default="448,256,128,1024")
    ap.add_argument("--out", default="gpurun_out/tune_sweep.jsonl")
    ap.add_argument("--crop-gb", type=float, default=2.0)
    ap.add_argument("--skip-copy", action="store_true", help="only the xform kernel's modes")
    ap.add_argument("--translate", action="store_true", help="per-image integer translation (xform kernel, tensor engine)")
    args = ap.parse_args()
    results = []
    for P in [int(v) for v in args.patches.split(",")]:
        gh = gw = max(2, 2048 // P)
        for dtype in (torch.float32, torch.uint8):
            elem = 1 if dtype == torch.uint8 else 4
            n = int(args.crop_gb * (1 << 30) / (3 * P * P * 4))
            n_img = max(1, min(n, int(5 * (1 << 30) / (3 * gh * P * gw * P * elem))))
            g = torch.Generator(device="cuda").manual_seed(0)
            u8 = torch.randint(0, 256, (n_img, 3, gh * P, gw * P), dtype=torch.uint8, device="cuda", generator=g)
            images = u8 if dtype == torch.uint8 else u8.float().div_(255)
            del u8
            s = ImageSet(images, P)
            idx = torch.arange(n, device="cuda")
            src = (idx % n_img).to(torch.int32)
            cell = (idx // n_img) % (gh * gw)
            pos = torch.stack([cell // gw, cell % gw], 1).contiguous()
            normalize = dtype == torch.uint8
            shifts = None
            if args.translate:
                shifts = torch.randint(-P // 2, P // 2 + 1, (n_img, 2), generator=torch.Generator().manual_seed(1)).int().cuda()
            for focus in (False, True):
                out = torch.empty(s.out_shape(n, focus), dtype=torch.float32, device="cuda")
                copy_mode = (not normalize) and (not focus) and not args.translate
                if copy_mode and args.skip_copy:
                    continue
                row = P * elem
                if copy_mode:
                    grid = [(S, D, ch, c) for (S, D) in ((2, 1), (3, 1), (3, 2), (4, 2), (4, 3), (6, 3), (6, 4), (8, 4), (12, 6))
                            for ch in (row * 4, row * 8, row * 16, row * 32, row * 64) for c in (1, 2, 3, 4)
                            if 4096 <= ch <= 65536 and (S * ch + 4200) * c <= 227 * 1024 and ch // row <= min(P, 256) and P % (ch // row) == 0]
                    tunes = [f"{S},{D},{ch},{c},0,0,0" for (S, D, ch, c) in grid]
                else:
                    grid = [(S, ch, c) for S in (2, 3, 4, 6, 8) for ch in (row * 8, row * 16, row * 32, row * 64, row * 128)
                            for c in (1, 2, 3, 4)
                            if 4096 <= ch <= 65536 and (S * ch + 256) * c <= 227 * 1024 and ch // row <= min(P, 256) and P % (ch // row) == 0]
                    tunes = [f"0,0,0,0,{S},{ch},{c}" for (S, ch, c) in grid]
                nbytes = n * 3 * P * P * (elem + 4)
                for engine in (("tensor",) if args.translate else ("tensor", "bulk")):
                    shape_rows = []
                    for t in tunes:
                        os.environ["JN_GATHER_TUNE"] = t
                        try:
                            ms = time_gather(s, pos, src, out, normalize, focus, engine, shifts=shifts)
                        except Exception as e:
                            continue
                        r = {"P": P, "src": "u8" if elem == 1 else "f32", "focus": focus, "engine": engine, "translate": bool(args.translate), "tune": t,
                             "ms": round(ms, 4), "GBps": round(nbytes / ms / 1e6, 1)}
                        shape_rows.append(r)
                    results.extend(shape_rows)
                    shape_rows.sort(key=lambda r: r["ms"])
                    print(f"== P={P} src={'u8' if elem == 1 else 'f32'} focus={int(focus)} engine={engine} n={n} ({len(shape_rows)} settings)")
                    for r in shape_rows[:5] + shape_rows[-1:]:
                        print(f"   {r['tune']:28s} {r['ms']:8.4f} ms {r['GBps']:8.1f} GB/s  {r['GBps'] / PEAK:.3f}")
                    sys.stdout.flush()
                del out
            del s, images
            torch.cuda.empty_cache()
    os.environ.pop("JN_GATHER_TUNE", None)
    with open(args.out, "w") as f:
        for r in results:
            f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
