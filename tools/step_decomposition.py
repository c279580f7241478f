#!/usr/bin/env python
"""Where does the time of an env step go?  K1 alone vs the fused native call, per engine / pipeline shape.

    python tools/step_decomposition.py fused|split|contig [auto|tensor|bulk] [reinforce|aerial]
    JN_PDL=0 python tools/step_decomposition.py fused            # without programmatic dependent launch
    JN_GATHER_TUNE=0,0,0,0,stages,chunk_bytes,ctas,batch python tools/step_decomposition.py fused tensor

fused  = the env's native step call (K2 + K1 behind it), CUDA events around each call;
split  = the state update alone, then K1 by itself into the history slot (events around K1 only);
contig = like split, into a fresh contiguous tensor.   Prints one JSON line (median / min / p90 over 4 rollouts);
results of round 2: profiles/r02/step_decomposition.jsonl."""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from jolineedle_b200 import _cabi
from jolineedle_b200.env.general_env import NeedleGeneralEnv

def main():
    mode = sys.argv[1]
    engine = sys.argv[2] if len(sys.argv) > 2 else "auto"
    wlname = sys.argv[3] if len(sys.argv) > 3 else "reinforce"
    dev = torch.device("cuda", 0)
    cls, batch = bench.WORKLOADS[wlname]
    wl = cls(batch, 0, dev, "u8"); wl.to_device()
    T = wl.T
    res = []
    for rep in range(6):
        env = NeedleGeneralEnv(wl.images, wl.boxes_dev, wl.PATCH, T, 1, stop_enabled=True, normalize=True, history=(mode != "contig"),
                               translate=wl.translate_dev, engine=engine)
        g = torch.Generator(device=dev).manual_seed(rep)
        actions = torch.randint(0, 9, (T, batch), device=dev, generator=g).unbind(0)
        env.reset()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(T)]
        if mode == "fused":
            for t in range(T):
                evs[t][0].record(); env.step(actions[t]); evs[t][1].record()
        else:  # K2 by itself (state only), then K1 by itself, each bracketed
            launch = env._set.bind(normalize=True, engine=engine, status=env._status, tag="x", shifts=env._shifts, shifts_aligned=env._shifts_aligned)
            for t in range(T):
                saved = env._set_handle
                env._set_handle = None  # state only
                env.step(actions[t])
                env._set_handle = saved
                out = env._history[:, t + 1] if mode == "split" else torch.empty(env._tile_shape, dtype=env._tile_dtype, device=dev)
                evs[t][0].record(); launch(env.positions, out); evs[t][1].record()
        torch.cuda.synchronize()
        if rep >= 2:
            res += [a.elapsed_time(b) for a, b in evs]
    res.sort()
    print(json.dumps({"mode": mode, "engine": engine, "workload": wlname, "tune": os.environ.get("JN_GATHER_TUNE"), "pdl": os.environ.get("JN_PDL", "1"),
                      "median_ms": round(res[len(res) // 2], 4), "mean_ms": round(sum(res) / len(res), 4), "min_ms": round(res[0], 4),
                      "p90_ms": round(res[int(len(res) * 0.9)], 4)}), flush=True)

main()
