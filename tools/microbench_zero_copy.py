#!/usr/bin/env python
"""PCIe side of the end-to-end path: gather straight out of pinned host images (zero-copy) vs a plain
pinned H2D memcpy of the same number of bytes."""
import argparse, json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jolineedle_b200.gather import ImageSet  # noqa: E402

P, gh, gw = 448, 5, 6  # defaults: the LARD grid (cfg 2 / cfg 3); --patch 256 --grid 8 for cfg-4 style tiles


def timed(fn, iters=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def main():
    global P, gh, gw
    ap = argparse.ArgumentParser()
    ap.add_argument("--patch", type=int, default=P)
    ap.add_argument("--grid", type=int, default=0, help="square patch grid per image (default: 5 x 6)")
    ap.add_argument("--tiles", type=int, default=480)
    ap.add_argument("--translate", action="store_true", help="per-image integer shifts (superset loads)")
    ap.add_argument("--default-tune-only", action="store_true")
    args = ap.parse_args()
    P = args.patch
    if args.grid:
        gh = gw = args.grid
    n_img, n = 48, args.tiles
    shifts = None
    if args.translate:
        shifts = torch.randint(-P // 2, P // 2 + 1, (n_img, 2), generator=torch.Generator().manual_seed(1)).int().cuda()
    for dtype in (torch.float32, torch.uint8):
        elem = 4 if dtype == torch.float32 else 1
        g = torch.Generator().manual_seed(0)
        host = torch.randint(0, 256, (n_img, 3, gh * P, gw * P), dtype=torch.uint8, generator=g)
        host = (host.float() / 255 if elem == 4 else host).pin_memory()
        idx = torch.arange(n)
        src = (idx % n_img).to(torch.int32).cuda()
        cell = (idx // n_img) % (gh * gw)
        pos = torch.stack([cell // gw, cell % gw], 1).cuda()
        nbytes = n * 3 * P * P * elem
        flat = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        ms = timed(lambda: dev.copy_(flat, non_blocking=True))
        print(json.dumps({"what": "pinned H2D memcpy", "bytes": nbytes, "ms": round(ms, 3), "GBps": round(nbytes / ms / 1e6, 1)}), flush=True)
        out = torch.empty((n, 3, P, P), dtype=torch.float32, device="cuda")
        for slabs, label in ((host, "one slab"), ([host[i] for i in range(n_img)], "list")):
            s = ImageSet(slabs, P, device="cuda")
            engines = ("tensor", "bulk", "ldg") if label == "one slab" else ("bulk",)
            if args.translate:
                engines = ("auto",) if label == "one slab" else ()
            for engine in engines:
                tunes = ["0"] if (engine == "ldg" or args.default_tune_only) else (
                    ["3,1,28672,2,3,14336,1", "6,3,28672,1,4,14336,2", "3,1,57344,2,3,28672,2", "6,4,14336,2,6,7168,3",
                     "12,9,14336,1,8,14336,1", "2,1,7168,4,2,7168,4", "8,6,24576,1,8,14336,1"])
                for t in tunes:
                    os.environ["JN_GATHER_TUNE"] = t
                    try:
                        ms = timed(lambda: s.gather(pos, src_index=src, out=out, normalize=(elem == 1), engine=engine,
                                                    shifts=shifts), iters=3)
                    except Exception as e:
                        print(json.dumps({"engine": engine, "tune": t, "error": repr(e)[:150]})); continue
                    print(json.dumps({"what": f"zero-copy gather {label}", "P": P, "tiles": n, "translate": args.translate,
                                      "src": "f32" if elem == 4 else "u8", "engine": engine,
                                      "tune": t, "ms": round(ms, 3), "pcie_GBps": round(nbytes / ms / 1e6, 1)}), flush=True)
            del s
        os.environ.pop("JN_GATHER_TUNE", None)


if __name__ == "__main__":
    main()
