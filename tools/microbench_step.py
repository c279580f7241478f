#!/usr/bin/env python
"""BASELINE cfg 5, step side: the env step kernel (K2) over grid sizes 5x6 .. 40x36 and batches 64 .. 8192.

Prints one JSON line per (grid, batch): median launch time of `jn_env_step` over back-to-back launches (CUDA
events on the launch stream), ns per episode-step, and the state bytes a step touches.  The kernel is latency /
launch bound (report ns per episode and warp-instruction efficiency, not flops); its
`smsp__thread_inst_executed_per_inst_executed` comes from ncu over this same script (tools/gpu_ci.sh ncu).

    python tools/microbench_step.py [--out profiles/r02/micro_step.jsonl] [--fused]

`--fused` times the whole native step call of the batched env instead (jn_env_step_gather: step kernel + the
gather of the new glimpses behind it, P = 64 tiles out of small images), i.e. the per-step cost a rollout sees.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jolineedle_b200 import _cabi  # noqa: E402

GRIDS = [(5, 6), (8, 8), (16, 16), (32, 32), (40, 36)]
BATCHES = [64, 256, 1024, 4096, 8192]


def bench_step(rows, cols, n, iters=200, stop_enabled=True):
    lib, dev = _cabi.lib(), torch.device("cuda", torch.cuda.current_device())
    g = torch.Generator(device="cuda").manual_seed(rows * 100 + cols + n)
    words = (rows * cols + 31) // 32
    pos = torch.stack([torch.randint(0, rows, (n,), device="cuda", generator=g),
                       torch.randint(0, cols, (n,), device="cuda", generator=g)], 1).contiguous()
    pos2 = torch.empty_like(pos)
    actions = torch.randint(0, 9, (iters + 8, n), device="cuda", generator=g)
    visited = torch.zeros((n, words), dtype=torch.int32, device="cuda")
    bbox = torch.randint(-2**31, 2**31 - 1, (n, words), dtype=torch.int32, device="cuda", generator=g)
    steps = torch.zeros(n, dtype=torch.long, device="cuda")
    stopped = torch.zeros(n, dtype=torch.bool, device="cuda")
    rewards = torch.empty(n, dtype=torch.float32, device="cuda")
    term = torch.empty(n, dtype=torch.bool, device="cuda")
    trunc = torch.empty(n, dtype=torch.bool, device="cuda")
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    stream = _cabi.stream_ptr(dev)
    bufs = [pos, pos2]

    def launch(i):
        a, b = bufs[i & 1], bufs[(i + 1) & 1]
        _cabi.check(lib.jn_env_step(a.data_ptr(), actions[i].data_ptr(), b.data_ptr(), visited.data_ptr(),
                                    bbox.data_ptr(), steps.data_ptr(), stopped.data_ptr(), rewards.data_ptr(),
                                    term.data_ptr(), trunc.data_ptr(), n, rows, cols, 1 << 20, -0.05,
                                    1 if stop_enabled else 0, status.data_ptr(), stream))

    for i in range(8):
        launch(i)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        launch(8 + i)
        ev[i + 1].record()
    torch.cuda.synchronize()
    times = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(iters))
    us = times[len(times) // 2] * 1e3
    # back-to-back launches without events in between: the throughput a rollout loop sees
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(iters):
        launch(8 + i)
    t1.record()
    torch.cuda.synchronize()
    us_stream = t0.elapsed_time(t1) * 1e3 / iters
    state_bytes = n * (16 + 16 + 8 + 2 * 4 * words + 4 + 8 + 8 + 1 + 1 + 1 + 1)  # pos in/out, action, bitmaps, outputs
    return {"kernel": "env_step_kernel", "grid": f"{rows}x{cols}", "words": words, "n": n,
            "us_per_launch": round(us, 2), "us_per_launch_streamed": round(us_stream, 2),
            "ns_per_episode": round(us_stream * 1e3 / n, 2), "state_MB": round(state_bytes / 1e6, 3),
            "state_GBps": round(state_bytes / us_stream / 1e3, 1)}


def bench_fused(rows, cols, n, iters=100, P=64):
    """The env's whole native step call at this grid: K2 + K1 (uint8 images normalised by the gather)."""
    from jolineedle_b200.env.general_env import NeedleGeneralEnv

    g = torch.Generator(device="cuda").manual_seed(7)
    images = torch.randint(0, 256, (n, 3, rows * P, cols * P), dtype=torch.uint8, device="cuda", generator=g)
    boxes = torch.zeros((n, 1, 4), dtype=torch.long, device="cuda")
    boxes[:, 0, 2:] = P
    T = iters + 8
    env = NeedleGeneralEnv(images, boxes, P, T, 1, stop_enabled=True, normalize=True)
    actions = torch.randint(0, 8, (T, n), device="cuda", generator=g).unbind(0)
    env.reset()
    for i in range(8):
        env.step(actions[i])
    torch.cuda.synchronize()
    import time
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h0 = time.perf_counter()
    t0.record()
    for i in range(iters):
        env.step(actions[8 + i])
    t1.record()
    h1 = time.perf_counter()
    torch.cuda.synchronize()
    us = t0.elapsed_time(t1) * 1e3 / iters
    tile = 3 * P * P * 5
    return {"kernel": "jn_env_step_gather (K2 + K1, P=64 uint8 -> float32)", "grid": f"{rows}x{cols}", "n": n,
            "us_per_step": round(us, 2), "host_us_per_step": round((h1 - h0) * 1e6 / iters, 2),
            "ns_per_episode": round(us * 1e3 / n, 2), "gather_GBps": round(n * tile / us / 1e3, 1)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--fused", action="store_true")
    ap.add_argument("--grids", default="")
    ap.add_argument("--batches", default="")
    args = ap.parse_args()
    grids = [tuple(int(v) for v in g.split("x")) for g in args.grids.split(",")] if args.grids else GRIDS
    batches = [int(v) for v in args.batches.split(",")] if args.batches else BATCHES
    lines = []
    for rows, cols in grids:
        for n in batches:
            if args.fused and n * 3 * rows * 64 * cols * 64 > (24 << 30):  # one image per episode: bounded pool
                continue
            rec = bench_fused(rows, cols, n) if args.fused else bench_step(rows, cols, n)
            print(json.dumps(rec), flush=True)
            lines.append(rec)
    if args.out:
        with open(args.out, "a") as f:
            for rec in lines:
                f.write(json.dumps(rec) + "\n")


if __name__ == "__main__":
    main()
