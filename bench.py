#!/usr/bin/env python
"""Benchmark of the gaze-environment hot path (metric: gaze-steps/s = glimpses cropped + scored).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA path)
    python bench.py --impl reference --gpus N --steps K ...  # reference arm: the CPU env on host cores

Default workload = BASELINE.json configs[1]: supervised trajectories on LARD-shaped images
(2048x2448 zero-padded to 2240x2688), patch 448, max-seq-len 8, 256 images per GPU, binomial
key points 0-3; uint8 images normalised by the gather (``--src f32`` feeds pre-normalised float32
images instead: same crops bit for bit, 4x the source bytes).  One *step* = one call of ``generate_trajectories`` on a fresh batch of seeds
(host plan + K0 + K3 + K1).  ``--workload reinforce`` runs BASELINE configs[2] instead (B=1024
episodes, T=20, STOP enabled): one step = reset + T env steps with seeded random actions.

Under torchrun every rank owns one GPU and its own shard of episodes (weak scaling, no data-path
collective); the step time is the max over ranks, measured with CUDA events on the launch stream.
"""
import argparse
import json
import os
import random
import statistics
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

P, GH, GW = 448, 5, 6  # LARD 2048x2448 padded to 2240x2688 (dataset.py:379-406)


# ----------------------------------------------------------------------------------------------
# synthetic workload
# ----------------------------------------------------------------------------------------------
def synth_boxes(rng, n_images, h, w):
    """1, 2 or 4 boxes per image, sides U[8, 448), inside the image; x1,y1,x2,y2."""
    out = []
    for _ in range(n_images):
        raw = []
        for _ in range(int(rng.choice([1, 2, 4]))):
            bw, bh = (int(v) for v in rng.integers(8, 448, size=2))
            x1, y1 = int(rng.integers(0, w - bw)), int(rng.integers(0, h - bh))
            raw.append((x1, y1, x1 + bw, y1 + bh))
        out.append(raw)
    return out


def device_images(b, h, w, seed, device, dtype):
    g = torch.Generator(device=device).manual_seed(seed)
    u8 = torch.randint(0, 256, (b, 3, h, w), dtype=torch.uint8, device=device, generator=g)
    if dtype == "u8":
        return u8
    out = torch.empty((b, 3, h, w), dtype=torch.float32, device=device)
    table = (torch.arange(256, dtype=torch.uint8).float() / 255).to(device)  # exact ToTensor values
    for i in range(0, b, 16):
        out[i:i + 16] = table[u8[i:i + 16].long()]
    return out


def read_back(tensors, non_blocking):
    """Device -> pinned host copies of a step's results; returns (host tensors, bytes)."""
    host = {}
    for k, v in tensors.items():
        dst = torch.empty(v.shape, dtype=v.dtype, pin_memory=True)
        dst.copy_(v, non_blocking=non_blocking)
        host[k] = dst
    return host, sum(v.numel() * v.element_size() for v in host.values())


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index, self.rows, self.proc = gpu_index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def mark(self, name):
        setattr(self, name, time.perf_counter())

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        # samples taken inside the timed region (the sampler itself runs from before the warm-up, so that
        # nvidia-smi's start-up does not land in the timed steps); if the region was shorter than the sampling
        # period, the two samples that bracket it
        t0, t1 = getattr(self, "t0", 0.0), getattr(self, "t1", float("inf"))
        inside = [r for (t, r) in self.rows if t0 <= t <= t1 + 0.3]
        if not inside:
            before = [r for (t, r) in self.rows if t < t0][-1:]
            after = [r for (t, r) in self.rows if t > t1][:1]
            inside = before + after
        rows = [r for r in inside if len(r) >= 9]
        sm = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in rows if r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            if len(r) >= 9:
                for name, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# supervised workload (cfg 2)
# ----------------------------------------------------------------------------------------------
class SupervisedWorkload:
    name = ("cfg2 supervised: LARD-shaped 2048x2448 synthetic zero-padded to 2240x2688, patch 448, max-seq-len 8, "
            "binomial keypoints 0-3")
    T, KMIN, KMAX, BINOMIAL = 8, 0, 3, True

    def __init__(self, batch, rank, device, src_dtype):
        from jolineedle_b200.utils import BBox, Position

        self.batch, self.rank, self.device, self.src_dtype = batch, rank, device, src_dtype
        self.h, self.w = GH * P, GW * P
        rng = np.random.default_rng(1234 + rank)
        self.raw_boxes = synth_boxes(rng, batch, self.h, self.w)
        self.bboxes = [[BBox(Position(y1, x1), Position(y2, x2)) for (x1, y1, x2, y2) in r] for r in self.raw_boxes]
        self.class_ids = [0] * batch
        self.images = None

    def to_device(self):
        slab = device_images(self.batch, self.h, self.w, 1234 + self.rank, self.device, self.src_dtype)
        self.slab = slab
        self.images = [slab[i] for i in range(self.batch)]

    def seeds(self, step):
        return [1_000_003 * (step + 1) + 7919 * self.rank + i for i in range(self.batch)]

    def run(self, step, images=None, device=None):
        from jolineedle_b200.env.simple_env import generate_trajectories

        random.seed(step * 31 + self.rank)
        batch = {"image": self.images if images is None else images, "bboxes": self.bboxes, "class_id": self.class_ids}
        self.stats = {}
        return generate_trajectories(batch, P, self.T, self.KMIN, self.KMAX, binomial_keypoints=self.BINOMIAL,
                                     seeds=self.seeds(step), normalize=(self.src_dtype == "u8"), device=device,
                                     stats=self.stats)

    def gaze_steps(self, out):
        return out["masks"].sum()  # recorded glimpses (padded slots are not glimpses)

    def host_tiles(self, out, stats):
        """Tiles that crossed PCIe: first occurrences of trajectory slots + detection patches no glimpse held."""
        zero = torch.zeros((), dtype=torch.long, device=out["masks"].device)
        return stats.get("host_traj_tiles", zero) + stats.get("host_det_tiles", zero)

    def gather_bytes(self, n_items, valid_items, tag):
        s_in = 1 if self.src_dtype == "u8" else 4
        tile = 3 * P * P
        return valid_items * tile * (s_in + 4) + (n_items - valid_items) * tile * 4  # padded slots: zero-fill writes

    def d2h(self, out, non_blocking=False):
        """What a trainer reads back on the host per step (actions/positions/masks/labels)."""
        keys = ("current_actions", "next_actions", "positions", "masks", "labels")
        return read_back({k: out[k] for k in keys}, non_blocking)

    # --- CPU reference port (oracle) on a bounded sample
    def cpu_sample(self, n_episodes, step):
        from oracle.traj_oracle import generate_trajectories_oracle

        cpu_images = getattr(self, "_cpu_images", None)
        if cpu_images is None or len(cpu_images) < n_episodes:
            g = torch.Generator().manual_seed(99 + self.rank)
            cpu_images = []
            for _ in range(n_episodes):
                u8 = torch.randint(0, 256, (3, self.h, self.w), dtype=torch.uint8, generator=g)
                cpu_images.append(u8.float() / 255)
            self._cpu_images = cpu_images
        boxes = [[((y1, x1), (y2, x2)) for (x1, y1, x2, y2) in r] for r in self.raw_boxes[:n_episodes]]
        random.seed(step * 31 + self.rank)
        t0 = time.perf_counter()
        out = generate_trajectories_oracle(cpu_images[:n_episodes], boxes, self.class_ids[:n_episodes], P, self.T,
                                           self.KMIN, self.KMAX, self.BINOMIAL, seeds=self.seeds(step)[:n_episodes])
        dt = time.perf_counter() - t0
        return float(out["masks"].sum()), dt


# ----------------------------------------------------------------------------------------------
# reinforce workload (cfg 3)
# ----------------------------------------------------------------------------------------------
class ReinforceWorkload:
    name = ("cfg3 reinforce: LARD-shaped 2240x2688 (padded), patch 448, max-seq-len 20, enable-stop, seeded random "
            "actions, uint8-resident images normalised on gather")
    T, PATCH, GRID, TRANSLATE = 20, P, (GH, GW), False

    def __init__(self, batch, rank, device, src_dtype):
        self.batch, self.rank, self.device, self.src_dtype = batch, rank, device, src_dtype
        self.h, self.w = self.GRID[0] * self.PATCH, self.GRID[1] * self.PATCH
        rng = np.random.default_rng(4321 + rank)
        raw = synth_boxes(rng, batch, self.h, self.w)
        nmax = max(len(r) for r in raw)
        boxes = np.zeros((batch, nmax, 4), dtype=np.int64)  # zero-padded rows like padded_collate_fn
        for i, r in enumerate(raw):
            boxes[i, :len(r)] = r
        self.translate = None
        if self.TRANSLATE:
            # augment-translate (dataset.py:157-226): per image an integer (tx, ty) inside the margins that keep
            # every box in the image, capped at a third of the image; boxes move with it, pixels are shifted by
            # the gather itself (zero fill), the translated image is never materialised
            shifts = np.zeros((batch, 2), dtype=np.int64)
            for i, r in enumerate(raw):
                a = np.array(r)
                lo_x, lo_y = min(self.w // 3, a[:, 0].min()), min(self.h // 3, a[:, 1].min())
                hi_x, hi_y = min(self.w // 3, self.w - a[:, 2].max()), min(self.h // 3, self.h - a[:, 3].max())
                tx = 0 if lo_x == 0 and hi_x == 0 else int(rng.integers(-lo_x, hi_x))
                ty = 0 if lo_y == 0 and hi_y == 0 else int(rng.integers(-lo_y, hi_y))
                shifts[i] = (tx, ty)
                boxes[i, :len(r)] += (tx, ty, tx, ty)
            self.translate = torch.from_numpy(shifts)
        self.boxes = torch.from_numpy(boxes)

    def to_device(self):
        self.images = device_images(self.batch, self.h, self.w, 4321 + self.rank, self.device, self.src_dtype)
        self.gen = torch.Generator(device=self.device)

    def run(self, step, images=None, device=None):
        from jolineedle_b200.env.general_env import NeedleGeneralEnv
        from jolineedle_b200.reinforce import rollout_tail

        imgs = self.images if images is None else images
        env = NeedleGeneralEnv(imgs, self.boxes, self.PATCH, self.T, 1, stop_enabled=True,
                               normalize=(self.src_dtype == "u8"), history=True, device=device,
                               translate=self.translate, zero_copy=images is not None)
        # (seeds the CPU generator reset() draws its start positions from; torch.manual_seed would also walk
        # through every accelerator backend, 0.1 ms a call)
        torch.default_generator.manual_seed(step * 31 + self.rank)
        self.gen.manual_seed(step * 31 + self.rank)
        b = self.batch
        actions = torch.randint(0, 9, (self.T, b), device=self.device, generator=self.gen)  # the policy's stand-in
        rewards, term = [], []
        env.reset()
        for t in range(self.T):
            _, r, te, _, _ = env.step(actions[t])
            rewards.append(r)
            term.append(te)
        rewards, term = torch.stack(rewards), torch.stack(term)  # [T, B] as the trainer collects them
        out = rollout_tail(rewards, term)
        out["positions"] = env.positions
        out["host_tiles"] = env.host_tiles  # tiles read over PCIe (zero-copy env only)
        return out

    def host_tiles(self, out, stats):
        return out["host_tiles"]

    def gaze_steps(self, out):
        return torch.tensor(float(self.batch * (self.T + 1)), device=self.device)

    def gather_bytes(self, n_items, valid_items, tag):
        s_in = 1 if self.src_dtype == "u8" else 4
        return n_items * 3 * self.PATCH * self.PATCH * (s_in + 4)

    def d2h(self, out, non_blocking=False):
        return read_back({k: out[k] for k in ("rewards", "returns", "masks")}, non_blocking)

    def cpu_sample(self, n_episodes, step):
        from oracle.gaze_oracle import GazeOracle, returns_oracle

        g = torch.Generator().manual_seed(99 + self.rank)
        imgs = getattr(self, "_cpu_images", None)
        if imgs is None or imgs.shape[0] != n_episodes:
            imgs = torch.randint(0, 256, (n_episodes, 3, self.h, self.w), dtype=torch.uint8, generator=g).float() / 255
            self._cpu_images = imgs
        boxes = self.boxes[:n_episodes].numpy()
        rng = np.random.default_rng(step)
        torch.manual_seed(step)
        t0 = time.perf_counter()
        env = GazeOracle(imgs, boxes, self.PATCH, self.T, 1, True, raster_masks=True)
        env.reset()
        rew, term = [], []
        for t in range(self.T):
            o = env.step(rng.integers(0, 9, size=n_episodes))
            rew.append(torch.from_numpy(o[1])); term.append(torch.from_numpy(o[2]))
        masks = torch.cat([torch.ones((n_episodes, 1), dtype=torch.bool), ~torch.stack(term, 1)], dim=1)
        returns_oracle(torch.stack(rew, 1), masks)
        dt = time.perf_counter() - t0
        return float(n_episodes * (self.T + 1)), dt


class AerialWorkload(ReinforceWorkload):
    name = ("cfg4 aerial: 8192x8192 synthetic images, patch 256 (32x32 grid, 1024-bit bitmaps), max-seq-len 32, "
            "enable-stop, augment-translate folded into the gather, seeded random actions, uint8-resident images "
            "normalised on gather")
    T, PATCH, GRID, TRANSLATE = 32, 256, (32, 32), True


WORKLOADS = {"supervised": (SupervisedWorkload, 256), "reinforce": (ReinforceWorkload, 1024),
             "aerial": (AerialWorkload, 256)}  # 256 x 201 MB uint8 = 48 GiB resident (+ the same again in the e2e arm)


# ----------------------------------------------------------------------------------------------
def reference_arm(args, rank, world):
    """The reference's CPU implementation of the path (oracle port: the reference is pure Python
    and cannot travel to the GPU box), all host threads, bounded sample per step."""
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    cls, default_batch = WORKLOADS[args.workload]
    sample = args.cpu_sample or {"supervised": 24, "reinforce": 8, "aerial": 2}[args.workload]
    wl = cls(sample, 0, "cpu", "f32")
    for s in range(args.warmup):
        wl.cpu_sample(sample, s)
    units, secs = 0.0, 0.0
    for s in range(args.steps):
        u, dt = wl.cpu_sample(sample, args.warmup + s)
        units += u
        secs += dt
    value = units / secs
    line = {
        "impl": "reference", "metric": "gaze_steps_per_sec", "value": value, "unit": "gaze-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl.name, "sample": f"{sample} episodes per step (bounded sample of the workload)"},
        "cpu_baseline": {"value": value, "unit": "gaze-steps/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{args.steps} steps x {sample} episodes, oracle port of the reference CPU env"},
        "e2e": {"value": value, "unit": "gaze-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="supervised", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="episodes per GPU (default: the BASELINE config's)")
    ap.add_argument("--src", default="u8", choices=["f32", "u8"],
                    help="image dtype: u8 = uint8 images normalised by the gather (x/255 like ToTensor, bit-identical "
                         "crops; SURVEY 8d's primary synthetic input), f32 = pre-normalised float32 images as the "
                         "reference's dataset hands them over")
    ap.add_argument("--cpu-sample", type=int, default=0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-pool", type=int, default=0,
                    help="distinct pinned host images per rank in the e2e arm (default: all of them up to 2 ranks, "
                         "64 beyond, to bound page-locked host memory at 8 ranks)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-timing", action="store_true", help="also time every gather of the e2e arm (diagnostics)")
    ap.add_argument("--no-clocks", action="store_true", help="do not sample nvidia-smi (A/B of the sampler's own cost)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    from jolineedle_b200.sharding import dist_env, max_over_ranks, sum_over_ranks

    rank, local_rank, world = dist_env()
    if args.impl == "reference":
        return reference_arm(args, rank, world)

    import torch.distributed as dist

    from jolineedle_b200 import _cabi, gather

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize(device)

    cls, default_batch = WORKLOADS[args.workload]
    src = args.src if args.workload == "supervised" else "u8"  # RL workloads keep uint8-resident images
    wl = cls(args.batch or default_batch, rank, device, src)
    # nvidia-smi is started here, seconds before the timed region: its NVML start-up stalls CUDA calls
    clocks = ClockSampler(local_rank)
    if not args.no_clocks and rank == 0:  # rank 0's GPU stands for the box: one poller, not one per rank
        clocks.__enter__()
    wl.to_device()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak, peak_kind = (peaks["hbm_gbs"], "measured") if "hbm_gbs" in peaks else (6650.0, "fallback")

    # ---- device-resident arm ------------------------------------------------------------------
    try:
        # warm-up runs the timed loop body verbatim (event-timed gathers, unit accounting) so that lazily
        # loaded kernels, allocator pools and pinned staging buffers all exist before the clock starts
        gather.TIMING = []
        units = torch.zeros((), dtype=torch.float64, device=device)
        out = None
        for s in range(args.warmup):
            out = wl.run(s)  # held across the next call like in the timed loop: two result sets are alive at once
            units += wl.gaze_steps(out)
        barrier()
        gather.TIMING = []
        units = torch.zeros((), dtype=torch.float64, device=device)
        valid = []
        launches0 = _cabi.launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        clocks.mark("t0")
        ev0.record()
        step_host = [time.perf_counter()]
        for s in range(args.steps):
            out = wl.run(args.warmup + s)
            g = wl.gaze_steps(out)
            units += g
            valid.append(g)
            step_host.append(time.perf_counter())
        ev1.record()
        barrier()
        clocks.mark("t1")
    finally:
        clocks.__exit__(None, None, None)
    host_ms = [round(1e3 * (b - a), 2) for a, b in zip(step_host, step_host[1:])]
    launches = _cabi.launch_count() - launches0
    timing, gather.TIMING = gather.TIMING, None
    ms = max_over_ranks(ev0.elapsed_time(ev1), device)
    total_units = sum_over_ranks(float(units.item()), device)
    value = total_units / (ms / 1e3)

    # roofline of the dominant kernel: the trajectory / step gather
    main_tag = "trajectory" if args.workload == "supervised" else "step"
    dur, byts = [], []
    per_step = [t for t in timing if t[0] == main_tag]
    k = len(per_step) // max(args.steps, 1)
    for i, (tag, n_items, e0, e1) in enumerate(per_step):
        v = int(valid[i // k].item()) if args.workload == "supervised" else n_items
        dur.append(e0.elapsed_time(e1))
        byts.append(wl.gather_bytes(n_items, v, tag))
    achieved = (sum(byts) / len(byts)) / (sum(dur) / len(dur) / 1e3) / 1e9 if dur else 0.0
    by_tag = {}
    for tag, n_items, e0, e1 in timing:
        by_tag.setdefault(tag, []).append(e0.elapsed_time(e1))
    gather_ms = {tag: round(sum(v) / len(v), 4) for tag, v in by_tag.items()}  # mean launch time of every gather
    items_by_tag = {}
    for tag, n_items, e0, e1 in timing:
        items_by_tag.setdefault(tag, []).append(n_items)
    gather_items = {tag: round(sum(v) / len(v), 1) for tag, v in items_by_tag.items()}
    # DRAM traffic per launch of that kernel from the committed `ncu --set full` capture of this workload at its
    # default batch (profiles/r01/ncu_summary.json, produced by tools/gpu_ci.sh ncu); null when there is none
    traffic, traffic_src = None, None
    try:
        summary = json.load(open(os.path.join(ROOT, "profiles", "r01", "ncu_summary.json")))
        capture = summary[f"{args.workload}_{src}"]
        prefix = "gather_xform_kernel" if src == "u8" else "gather_copy_kernel"
        if wl.batch == default_batch:  # the capture was taken at the default batch
            recs = [r for k_, v in capture.items() if k_.startswith(prefix) for r in v]
            rec = max(recs, key=lambda r: r["duration_s"])  # the trajectory / step gather (detection gathers are smaller)
            traffic = int(rec["dram_traffic_bytes"])
            traffic_src = f"profiles/r01/ncu_summary.json[{args.workload}_{src}] (ncu --set full)"
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": f"{peak_kind} (MEASURED_PEAKS.json)",
                "kernel": "gather (K1), tag=" + main_tag, "launches_timed": len(dur),
                "avg_launch_ms": round(sum(dur) / len(dur), 4) if dur else None,
                "algorithmic_bytes_per_launch": int(sum(byts) / len(byts)) if byts else 0}

    # ---- end-to-end arm: HOST buffers in, host-visible results out ---------------------------------
    e2e = None
    if not args.no_e2e:
        def pinned(t):
            return torch.empty(t.shape, dtype=t.dtype, pin_memory=True).copy_(t)

        if args.workload == "supervised":
            pool = args.e2e_pool or (wl.batch if world <= 2 else 64)
            distinct = [pinned(img) for img in wl.images[:pool]]
            host = [distinct[i % len(distinct)] for i in range(wl.batch)]
        else:
            host = pinned(wl.images)
        full = sum(t.numel() * t.element_size() for t in host) if isinstance(host, list) else host.numel() * host.element_size()
        zero_copy = True  # pinned images are gathered in place (supervised lists and the batched RL env alike)
        wl_patch = getattr(wl, "PATCH", P)
        e2e_steps = max(2, min(args.steps, 10))
        eunits, d2h_bytes, h2d = 0.0, 0, 0
        side = torch.cuda.Stream(device)
        tile_bytes = 3 * wl_patch * wl_patch * (1 if src == "u8" else 4)

        def launch(step):
            """Host plan + launches of one step (asynchronous); its device -> host read is queued on a side
            stream behind an event, so that it overlaps the next step's PCIe reads like a prefetching trainer."""
            out = wl.run(step, images=host, device=device)
            stats = dict(getattr(wl, "stats", {}))
            done = torch.cuda.Event()
            done.record()
            with torch.cuda.stream(side):
                side.wait_event(done)
                hostres, nbytes = wl.d2h(out, non_blocking=True)
                counts = read_back({"tiles": wl.host_tiles(out, stats).reshape(1)}, True)[0]["tiles"]
                read = torch.cuda.Event()
                read.record(side)
            # `out` and `stats` were allocated on the launch stream and are read on the side stream: both stay
            # referenced until finish() so that the allocator cannot hand their memory to the next step early
            return (out, stats), hostres, nbytes, counts, read

        def finish(pending):
            nonlocal eunits, d2h_bytes, h2d
            out, hostres, nbytes, counts, read = pending
            read.synchronize()  # the step's results are on the host
            d2h_bytes = nbytes
            eunits += float(hostres["masks"].sum()) if args.workload == "supervised" else float(wl.batch * (wl.T + 1))
            h2d += float(counts.sum()) * tile_bytes

        def pipeline(first_step, n_steps):
            pending = None
            for s in range(n_steps):
                nxt = launch(first_step + s)
                if pending is not None:
                    finish(pending)
                pending = nxt
            finish(pending)
            torch.cuda.current_stream(device).wait_stream(side)

        # warm-up runs the timed loop verbatim: pinned read-back buffers, the side stream's allocator pool and
        # the zero-copy path's scratch all exist before the clock starts (a cudaHostAlloc or cudaMalloc inside a
        # 0.2 s window would halve the figure)
        pipeline(0, 3)
        barrier()
        eunits, d2h_bytes, h2d = 0.0, 0, 0
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        gather.TIMING = [] if args.e2e_timing else None
        t0.record()
        pipeline(args.warmup, e2e_steps)
        t1.record()
        barrier()
        ems = max_over_ranks(t0.elapsed_time(t1), device)
        e2e_tags = {}
        for tag, n_items, ev_a, ev_b in (gather.TIMING or []):
            e2e_tags.setdefault(tag, []).append((n_items, ev_a.elapsed_time(ev_b)))
        gather.TIMING = None
        how = "read in place by the gather kernels (zero-copy over PCIe: only glimpsed tiles move, each once)"
        e2e = {"value": sum_over_ranks(eunits, device) / (ems / 1e3), "unit": "gaze-steps/s",
               "h2d_bytes_per_step": int(h2d / e2e_steps), "d2h_bytes_per_step": int(d2h_bytes), "steps": e2e_steps,
               **({"gather_by_tag": {k: {"items": round(sum(n for n, _ in v) / len(v), 1),
                                          "ms": round(sum(m for _, m in v) / len(v), 3)} for k, v in e2e_tags.items()}}
                  if e2e_tags else {}),
               "host_buffers": f"pinned {src} images ({full} bytes addressed on the host"
                               + (f", {len(distinct)} distinct" if isinstance(host, list) else "") + f"), {how}"}
        del host

    # ---- CPU baseline beside it (rank 0, N=1 only) ------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        torch.set_num_threads(os.cpu_count() or 1)
        sample = args.cpu_sample or {"supervised": 32, "reinforce": 8, "aerial": 2}[args.workload]
        budget_s = 12.0  # bounded sample: ~10-30 s of CPU work
        wl.cpu_sample(sample, 0)
        u_sum, t_sum, reps = 0.0, 0.0, 0
        while t_sum < budget_s and reps < 400:
            u, dt = wl.cpu_sample(sample, 100 + reps)
            u_sum += u
            t_sum += dt
            reps += 1
        cpu = {"value": u_sum / t_sum, "unit": "gaze-steps/s", "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"{reps} x {sample} episodes of the same workload through the oracle port of the reference "
                         f"CPU env ({t_sum:.1f} s of CPU work)"}

    if rank == 0:
        line = {
            "metric": "gaze_steps_per_sec", "value": value, "unit": "gaze-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl.name, "episodes_per_gpu": wl.batch, "resident_images": src,
                       "l2": "inputs larger than L2: every step reads fresh tiles of a multi-GB image pool and "
                             "writes GBs of crops",
                       "parallelism": f"episodes sharded over {world} GPU(s), no data-path collective"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches,
            "clocks": clocks.summary(), "host_ms_per_step": host_ms, "gather_ms_by_tag": gather_ms,
            "gather_items_by_tag": gather_items,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
