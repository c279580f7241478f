#!/usr/bin/env python
"""Benchmark of the gaze-environment hot path (metric: gaze-steps/s = glimpses cropped + scored).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA path)
    python bench.py --impl reference --gpus N --steps K ...  # reference arm: the unmodified reference env on host cores

The parsed line is BASELINE.json configs[2], the configuration the metric's "1/2/4/8 B200" clause is quoted on:
reinforce mode, LARD-shaped 2048x2448 images zero-padded to 2240x2688, patch 448, max-seq-len 20, STOP enabled,
1024 episodes per GPU, uint8-resident images normalised by the gather.  One *step* = one rollout of the env path:
env construction (K0 overlap bitmaps) + reset + 20 env steps with seeded random actions (the policy's stand-in;
the model forward is outside the path) + the returns tail; every env step is one native call (K2 + K1).
The same run also measures configs[1] (supervised trajectories, 256 images, T=8, binomial key points 0-3) and
configs[3] (8192x8192 aerial images, patch 256, T=32, augment-translate) and nests their results under "also".
``--workload`` picks another primary, ``--also none`` drops the secondary workloads.

Under torchrun every rank owns one GPU and its own shard of episodes (weak scaling, no data-path collective);
the step time is the max over ranks, measured with CUDA events on the launch stream.
"""
import argparse
import gc
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
PROFILE_ROUND = "r02"  # profiles/<round>/ncu_summary.json supplies roofline.traffic


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


def cpu_images(n, h, w, seed):
    """float32 [0, 1] images as the reference's dataset hands them over (ToTensor, dataset.py:240)."""
    g = torch.Generator().manual_seed(seed)
    return [torch.randint(0, 256, (3, h, w), dtype=torch.uint8, generator=g).float() / 255 for _ in range(n)]


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
        self.windows = []

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

    def close(self):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
            self.proc = None

    def summary(self, t0, t1):
        # samples taken inside the timed region (the sampler itself runs from before the warm-up, so that
        # nvidia-smi's start-up does not land in the timed steps); if the region was shorter than the sampling
        # period, the two samples that bracket it
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
            for name, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# the reference itself on the host cores (baseline/_ref, unmodified), with the oracle port as fallback
# ----------------------------------------------------------------------------------------------
def reference_kind():
    from baseline import ref_env

    return "reference" if ref_env.available() else "port"


def returns_tail_cpu(rewards, terminated):
    """reinforce.py:186-202 on stacked [B, T] CPU tensors (the trainer module itself cannot be imported)."""
    b = rewards.shape[0]
    masks = torch.cat([torch.ones((b, 1), dtype=torch.bool), ~terminated], dim=1)
    logit_masks = torch.roll(masks[:, 1:], shifts=1, dims=(1,))
    logit_masks[:, 0] = True
    back = torch.cumsum(torch.flip(rewards, dims=(1,)) * torch.flip(logit_masks, dims=(1,)), dim=1)
    return torch.flip(back, dims=(1,)), masks, logit_masks


# ----------------------------------------------------------------------------------------------
# supervised workload (cfg 2)
# ----------------------------------------------------------------------------------------------
class SupervisedWorkload:
    key = "cfg2"
    name = ("cfg2 supervised: LARD-shaped 2048x2448 synthetic zero-padded to 2240x2688, patch 448, max-seq-len 8, "
            "binomial keypoints 0-3")
    T, KMIN, KMAX, BINOMIAL, PATCH = 8, 0, 3, True, P
    main_tag = "trajectory"
    launches_per_step = 4

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

    def host_images(self, pool):
        def pinned(t):
            return torch.empty(t.shape, dtype=t.dtype, pin_memory=True).copy_(t)

        distinct = [pinned(img) for img in self.images[:pool]]
        return [distinct[i % len(distinct)] for i in range(self.batch)], len(distinct)

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

    def host_gaze_steps(self, hostres):
        return float(hostres["masks"].sum())

    def host_tiles(self, out, stats):
        """Tiles that crossed PCIe: first occurrences of trajectory slots + detection patches no glimpse held."""
        zero = torch.zeros((), dtype=torch.long, device=out["masks"].device)
        return stats.get("host_traj_tiles", zero) + stats.get("host_det_tiles", zero)

    def gather_bytes(self, n_items, valid_items):
        s_in = 1 if self.src_dtype == "u8" else 4
        tile = 3 * P * P
        return valid_items * tile * (s_in + 4) + (n_items - valid_items) * tile * 4  # padded slots: zero-fill writes

    def d2h(self, out, non_blocking=False):
        """What a trainer reads back on the host per step (actions/positions/masks/labels)."""
        keys = ("current_actions", "next_actions", "positions", "masks", "labels")
        return read_back({k: out[k] for k in keys}, non_blocking)

    # --- CPU arm: the reference's generate_trajectories loop (supervised.py:116-136) on a bounded sample
    cpu_default_sample = 24

    def cpu_prepare(self, n_episodes):
        if getattr(self, "_cpu_images", None) is None or len(self._cpu_images) < n_episodes:
            self._cpu_images = cpu_images(n_episodes, self.h, self.w, 99 + self.rank)

    def cpu_sample(self, n_episodes, step):
        self.cpu_prepare(n_episodes)
        imgs, seeds = self._cpu_images[:n_episodes], self.seeds(step)[:n_episodes]
        random.seed(step * 31 + self.rank)
        if reference_kind() == "reference":
            from baseline import ref_env

            _, se, _, ut = ref_env.load()
            t0 = time.perf_counter()
            samples = []
            for i in range(n_episodes):
                boxes = [ut.BBox(ut.Position(y1, x1), ut.Position(y2, x2)) for (x1, y1, x2, y2) in self.raw_boxes[i]]
                env = se.NeedleSimpleEnv(imgs[i], P, boxes, seeds[i])
                sample = env.generate_sample(self.T, min_keypoints=self.KMIN, max_keypoints=self.KMAX, position=None,
                                             binomial_keypoints=self.BINOMIAL)
                sample["class_id"] = torch.tensor(self.class_ids[i], dtype=torch.long, device=sample["patches"].device)
                samples.append(sample)
            out = se.NeedleSimpleEnv.collate_fn(samples)
            dt = time.perf_counter() - t0
            return float(out["masks"].sum()), dt
        from oracle.traj_oracle import generate_trajectories_oracle

        boxes = [[((y1, x1), (y2, x2)) for (x1, y1, x2, y2) in r] for r in self.raw_boxes[:n_episodes]]
        t0 = time.perf_counter()
        out = generate_trajectories_oracle(imgs, boxes, self.class_ids[:n_episodes], P, self.T, self.KMIN, self.KMAX,
                                           self.BINOMIAL, seeds=seeds)
        dt = time.perf_counter() - t0
        return float(out["masks"].sum()), dt


# ----------------------------------------------------------------------------------------------
# reinforce workload (cfg 3)
# ----------------------------------------------------------------------------------------------
class ReinforceWorkload:
    key = "cfg3"
    name = ("cfg3 reinforce: LARD-shaped 2048x2448 synthetic zero-padded to 2240x2688, patch 448, max-seq-len 20, "
            "enable-stop, seeded random actions")
    T, PATCH, GRID, TRANSLATE = 20, P, (GH, GW), False
    main_tag = "step"

    def __init__(self, batch, rank, device, src_dtype):
        self.batch, self.rank, self.device, self.src_dtype = batch, rank, device, src_dtype
        self.h, self.w = self.GRID[0] * self.PATCH, self.GRID[1] * self.PATCH
        self.launches_per_step = self.T + 1
        self.host_split = []
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
        # the trainer moves the collated boxes to the device with the images (reinforce.py:313-315)
        self.boxes_dev = self.boxes.to(self.device)
        self.boxes_pinned = torch.empty(self.boxes.shape, dtype=self.boxes.dtype, pin_memory=True).copy_(self.boxes)
        self.translate_dev = None if self.translate is None else self.translate.to(self.device)

    def host_images(self, pool):
        host = torch.empty(self.images.shape, dtype=self.images.dtype, pin_memory=True).copy_(self.images)
        return host, self.batch

    def run(self, step, images=None, device=None):
        from jolineedle_b200.env.general_env import NeedleGeneralEnv
        from jolineedle_b200.reinforce import rollout_tail

        e2e = images is not None
        c0 = time.perf_counter()
        env = NeedleGeneralEnv(self.images if not e2e else images, self.boxes_pinned if e2e else self.boxes_dev,
                               self.PATCH, self.T, 1, stop_enabled=True, normalize=(self.src_dtype == "u8"),
                               history=True, device=device, translate=self.translate_dev, zero_copy=e2e)
        # (seeds the CPU generator reset() draws its start positions from; torch.manual_seed would also walk
        # through every accelerator backend, 0.1 ms a call)
        torch.default_generator.manual_seed(step * 31 + self.rank)
        self.gen.manual_seed(step * 31 + self.rank)
        # the policy's stand-in: one row of action codes per env step
        actions = torch.randint(0, 9, (self.T, self.batch), device=self.device, generator=self.gen).unbind(0)
        c1 = time.perf_counter()
        env.reset()
        c2 = time.perf_counter()
        for t in range(self.T):
            env.step(actions[t])
        c3 = time.perf_counter()
        rewards_tn, terminated_tn, _ = env.rollout_buffers()  # [T, B] rings the steps wrote: nothing to stack
        out = rollout_tail(rewards_tn, terminated_tn)
        out["positions"] = env.positions
        out["host_tiles"] = env.host_tiles  # tiles read over PCIe (zero-copy env only)
        del env
        c4 = time.perf_counter()
        # where the launching thread spends a rollout (diagnostic; ms): env construction + actions, reset, the T
        # step calls, returns + teardown
        self.host_split.append((c1 - c0, c2 - c1, c3 - c2, c4 - c3))
        return out

    def host_tiles(self, out, stats):
        return out["host_tiles"]

    def gaze_steps(self, out):
        return float(self.batch * (self.T + 1))

    def host_gaze_steps(self, hostres):
        return float(self.batch * (self.T + 1))

    def gather_bytes(self, n_items, valid_items):
        s_in = 1 if self.src_dtype == "u8" else 4
        return n_items * 3 * self.PATCH * self.PATCH * (s_in + 4)

    def d2h(self, out, non_blocking=False):
        return read_back({k: out[k] for k in ("rewards", "returns", "masks")}, non_blocking)

    # --- CPU arm: the reference's NeedleGeneralEnv (construction + reset + T steps + returns tail)
    cpu_default_sample = 8

    def cpu_prepare(self, n_episodes):
        imgs = getattr(self, "_cpu_images", None)
        if imgs is None or imgs.shape[0] != n_episodes:
            self._cpu_images = torch.stack(cpu_images(n_episodes, self.h, self.w, 99 + self.rank))

    def cpu_sample(self, n_episodes, step):
        self.cpu_prepare(n_episodes)
        imgs, boxes = self._cpu_images, self.boxes[:n_episodes]
        torch.manual_seed(step)
        actions = torch.randint(0, 9, (self.T, n_episodes), generator=torch.Generator().manual_seed(step))
        if reference_kind() == "reference":
            from baseline import ref_env

            ge = ref_env.load()[0]
            t0 = time.perf_counter()
            env = ge.NeedleGeneralEnv(imgs, boxes, self.PATCH, self.T, 1, True)
            env.reset()
            rew, term = [], []
            for t in range(self.T):
                _, r, te, _, _ = env.step(actions[t])
                rew.append(r); term.append(te)
            returns_tail_cpu(torch.stack(rew, dim=1), torch.stack(term, dim=1))
            dt = time.perf_counter() - t0
            return float(n_episodes * (self.T + 1)), dt
        from oracle.gaze_oracle import GazeOracle, returns_oracle

        t0 = time.perf_counter()
        env = GazeOracle(imgs, boxes.numpy(), self.PATCH, self.T, 1, True, raster_masks=True)
        env.reset()
        rew, term = [], []
        for t in range(self.T):
            o = env.step(actions[t].numpy())
            rew.append(torch.from_numpy(o[1])); term.append(torch.from_numpy(o[2]))
        masks = torch.cat([torch.ones((n_episodes, 1), dtype=torch.bool), ~torch.stack(term, 1)], dim=1)
        returns_oracle(torch.stack(rew, 1), masks)
        dt = time.perf_counter() - t0
        return float(n_episodes * (self.T + 1)), dt


class AerialWorkload(ReinforceWorkload):
    key = "cfg4"
    name = ("cfg4 aerial: 8192x8192 synthetic images, patch 256 (32x32 grid, 1024-bit bitmaps), max-seq-len 32, "
            "enable-stop, augment-translate folded into the gather, seeded random actions")
    T, PATCH, GRID, TRANSLATE = 32, 256, (32, 32), True
    cpu_default_sample = 2  # (the CPU arm crops the untranslated images: the shift itself is dataset-side work there)


# name -> (class, episodes per GPU)
WORKLOADS = {"supervised": (SupervisedWorkload, 256), "reinforce": (ReinforceWorkload, 1024),
             "aerial": (AerialWorkload, 256)}  # 256 x 201 MB uint8 = 48 GiB resident


def config_of(wl_cls, batch, src, n_gpus):
    """The `config` object of a line -- the same for our arm and the reference arm of a workload."""
    return {"workload": wl_cls.name, "episodes_per_gpu": batch, "resident_images": src,
            "l2": "inputs larger than L2: every step reads fresh tiles of a multi-GB image pool and writes GBs of crops",
            "parallelism": f"episodes sharded over {n_gpus} GPU(s), no data-path collective"}


# ----------------------------------------------------------------------------------------------
def cpu_arm(wl, sample, steps, warmup=1, budget_s=None):
    """Time the CPU implementation of the path on a bounded sample: `steps` samples (or as many as fit in
    `budget_s` seconds).  Returns (gaze-steps/s, seconds per sample, samples run, total seconds)."""
    torch.set_num_threads(os.cpu_count() or 1)
    wl.cpu_prepare(sample)
    for s in range(warmup):
        wl.cpu_sample(sample, s)
    units, secs, reps = 0.0, 0.0, 0
    while (reps < steps) if budget_s is None else (secs < budget_s and reps < 400):
        u, dt = wl.cpu_sample(sample, 100 + reps)
        units += u
        secs += dt
        reps += 1
    return units / secs, secs / max(reps, 1), reps, secs


def cpu_description(kind, reps, sample, secs):
    what = ("the unmodified reference env (baseline/_ref: src/env/*.py of jolibrain/jolineedle, three third-party "
            "stand-ins)" if kind == "reference" else "the oracle port of the reference CPU env")
    return f"{reps} x {sample} episodes of the same workload through {what} ({secs:.1f} s of CPU work)"


def reference_arm(args, rank, world):
    """The reference's own CPU implementation of the path (unmodified, from baseline/_ref; the oracle port only
    when that copy is absent), all host threads, a bounded sample of the workload per step."""
    if rank != 0:
        return
    cls, default_batch = WORKLOADS[args.workload]
    src = args.src if args.workload == "supervised" else "u8"
    sample = args.cpu_sample or cls.cpu_default_sample
    wl = cls(sample, 0, "cpu", "f32")
    value, per_step, reps, secs = cpu_arm(wl, sample, args.steps, warmup=args.warmup)
    kind = reference_kind()
    line = {
        "impl": "reference", "metric": "gaze_steps_per_sec", "value": value, "unit": "gaze-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_of(cls, args.batch or default_batch, src, args.gpus),
        "cpu_baseline": {"value": value, "unit": "gaze-steps/s", "cores": torch.get_num_threads(), "kind": kind,
                         "sample": cpu_description(kind, reps, sample, secs)},
        "e2e": {"value": value, "unit": "gaze-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
def ncu_capture(workload_key, src, batch, default_batch, kernel_prefix):
    """(DRAM bytes, duration in seconds, source note) of one launch of the dominant kernel from the committed
    `ncu --set full` capture of this workload at its config batch -- only when the capture was taken from the
    sources the loaded library was built from; (None, None, None) otherwise."""
    from jolineedle_b200 import buildinfo

    try:
        path = os.path.join(ROOT, "profiles", PROFILE_ROUND, "ncu_summary.json")
        summary = json.load(open(path))
        name = f"{workload_key}_{src}"
        meta = summary.get("_meta", {}).get(name, {})
        if batch != default_batch or meta.get("source_hash") != buildinfo.library_source_hash():
            return None, None, None
        recs = [r for k_, v in summary[name].items() if k_.startswith(kernel_prefix) for r in v]
        rec = max(recs, key=lambda r: r["duration_s"])  # the trajectory / step gather (detection gathers are smaller)
        return int(rec["dram_traffic_bytes"]), float(rec["duration_s"]), (
            f"profiles/{PROFILE_ROUND}/ncu_summary.json[{name}] (ncu --set full, source hash {meta['source_hash']})")
    except Exception:
        return None, None, None


def measure_pcie(device, barrier, reps=4, nbytes=1 << 30):
    """Pinned host -> device cudaMemcpyAsync bandwidth of this rank while every rank does the same (what the box
    can give the end-to-end arm)."""
    host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    dev = torch.empty(nbytes, dtype=torch.uint8, device=device)
    dev.copy_(host, non_blocking=True)
    barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(reps):
        dev.copy_(host, non_blocking=True)
    t1.record()
    barrier()
    return reps * nbytes / (t0.elapsed_time(t1) / 1e3) / 1e9


def run_workload(wl_name, args, ctx, primary):
    """Device-resident arm + end-to-end arm of one workload; returns the fields of its JSON object."""
    import torch.distributed as dist  # noqa: F401

    from jolineedle_b200 import _cabi, gather
    from jolineedle_b200.sharding import max_over_ranks, sum_over_ranks

    rank, world, device, barrier, clocks, peak, peak_kind = (ctx[k] for k in (
        "rank", "world", "device", "barrier", "clocks", "peak", "peak_kind"))
    cls, default_batch = WORKLOADS[wl_name]
    src = args.src if wl_name == "supervised" else "u8"  # RL workloads keep uint8-resident images
    batch = (args.batch if primary else 0) or default_batch
    wl = cls(batch, rank, device, src)
    wl.to_device()

    # ---- device-resident arm ------------------------------------------------------------------
    # warm-up runs the timed loop body verbatim (event-timed gathers, unit accounting) so that lazily
    # loaded kernels, allocator pools and pinned staging buffers all exist before the clock starts
    gather.TIMING = gather.LaunchTimer((args.warmup + 1) * wl.launches_per_step)
    units = torch.zeros((), dtype=torch.float64, device=device)
    out = None
    for s in range(args.warmup):
        out = wl.run(s)  # held across the next call like in the timed loop: two result sets are alive at once
        units += wl.gaze_steps(out)
    barrier()
    gather.TIMING = gather.LaunchTimer((args.steps + 1) * wl.launches_per_step)
    units = torch.zeros((), dtype=torch.float64, device=device)
    valid = []
    launches0 = _cabi.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if hasattr(wl, "host_split"):
        wl.host_split.clear()
    mem0, gc0 = torch.cuda.memory_stats(device), [g["collections"] for g in gc.get_stats()]
    barrier()
    t_begin = time.perf_counter()
    ev0.record()
    step_host, step_cpu = [time.perf_counter()], [time.thread_time()]
    for s in range(args.steps):
        out = wl.run(args.warmup + s)
        g = wl.gaze_steps(out)
        units += g
        valid.append(g)
        step_host.append(time.perf_counter())
        step_cpu.append(time.thread_time())
    ev1.record()
    barrier()
    t_end = time.perf_counter()
    host_ms = [round(1e3 * (b - a), 2) for a, b in zip(step_host, step_host[1:])]
    # CPU time of the launching thread per step: well below host_ms = the thread was waiting (driver, queue), not working
    host_cpu_ms = [round(1e3 * (b - a), 2) for a, b in zip(step_cpu, step_cpu[1:])]
    launches = _cabi.launch_count() - launches0
    mem1 = torch.cuda.memory_stats(device)
    # host-side diagnostics of the timed region: cudaMalloc / cudaFree calls of the caching allocator, python GC
    # passes per generation, and (RL workloads) the split of a rollout's host time
    host_diag = {"device_allocs": mem1.get("num_device_alloc", 0) - mem0.get("num_device_alloc", 0),
                 "device_frees": mem1.get("num_device_free", 0) - mem0.get("num_device_free", 0),
                 "alloc_retries": mem1.get("num_alloc_retries", 0) - mem0.get("num_alloc_retries", 0),
                 "gc_collections": [g["collections"] - g0 for g, g0 in zip(gc.get_stats(), gc0)]}
    if getattr(wl, "host_split", None):
        rows = wl.host_split
        host_diag["rollout_split_ms"] = dict(zip(
            ("construct", "reset", "steps", "tail"),
            (round(1e3 * sum(r[i] for r in rows) / len(rows), 3) for i in range(4))))
    timing, gather.TIMING = gather.TIMING.records, None
    ms = max_over_ranks(ev0.elapsed_time(ev1), device)
    total_units = sum_over_ranks(float(units.item()), device)
    value = total_units / (ms / 1e3)

    # roofline of the dominant kernel: the trajectory / step gather
    dur, byts = [], []
    per_step = [t for t in timing if t[0] == wl.main_tag]
    k = max(len(per_step) // max(args.steps, 1), 1)
    for i, (tag, n_items, e0, e1) in enumerate(per_step):
        v = int(valid[min(i // k, len(valid) - 1)].item()) if wl_name == "supervised" else n_items
        dur.append(e0.elapsed_time(e1))
        byts.append(wl.gather_bytes(n_items, v))
    achieved = (sum(byts) / len(byts)) / (sum(dur) / len(dur) / 1e3) / 1e9 if dur else 0.0
    by_tag, items_by_tag = {}, {}
    for tag, n_items, e0, e1 in timing:
        by_tag.setdefault(tag, []).append(e0.elapsed_time(e1))
        items_by_tag.setdefault(tag, []).append(n_items)
    gather_ms = {tag: round(sum(v) / len(v), 4) for tag, v in by_tag.items()}  # mean launch time of every gather
    gather_items = {tag: round(sum(v) / len(v), 1) for tag, v in items_by_tag.items()}
    prefix = "gather_xform_kernel" if src == "u8" else "gather_copy_kernel"
    traffic, ncu_s, traffic_src = ncu_capture(wl_name, src, batch, default_batch, prefix)
    fused = wl_name != "supervised"
    roofline = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": f"{peak_kind} (MEASURED_PEAKS.json)",
                "kernel": f"{prefix} (K1), tag={wl.main_tag}" + (
                    "; events bracket the native step call: the gather + env_step_kernel launched behind it "
                    "with programmatic dependent launch" if fused else ""),
                "launches_timed": len(dur), "avg_launch_ms": round(sum(dur) / len(dur), 4) if dur else None,
                "algorithmic_bytes_per_launch": int(sum(byts) / len(byts)) if byts else 0}
    if ncu_s and byts:  # the kernel by itself (ncu's gpu__time_duration of the same capture): no launch latency, no K2
        roofline["kernel_ms_ncu"] = round(ncu_s * 1e3, 4)
        roofline["frac_kernel_ncu"] = round(sum(byts) / len(byts) / ncu_s / 1e9 / peak, 4)
    result = {"value": value, "ms_per_step": ms / args.steps, "roofline": roofline, "gpu_launches": launches,
              "host_ms_per_step": host_ms, "host_cpu_ms_per_step": host_cpu_ms, "gather_ms_by_tag": gather_ms,
              "gather_items_by_tag": gather_items, "host_diag": host_diag,
              "clocks": clocks.summary(t_begin, t_end) if rank == 0 else None}
    if fused:  # host time of one env step (python + one native call), the launch-bound regime of small batches
        result["host_us_per_env_step"] = round(1e3 * statistics.median(host_ms) / (wl.T + 1), 1)
    del out

    # ---- end-to-end arm: HOST buffers in, host-visible results out ---------------------------------
    run_e2e = not args.no_e2e and (primary or world == 1)
    if run_e2e:
        pool = args.e2e_pool or (batch if wl_name != "supervised" else min(batch, 256))
        host, n_distinct = wl.host_images(pool)
        full = (sum(t.numel() * t.element_size() for t in host[:n_distinct]) if isinstance(host, list)
                else host.numel() * host.element_size())
        e2e_steps = max(2, min(args.steps, 10))
        side = torch.cuda.Stream(device)
        tile_bytes = 3 * wl.PATCH * wl.PATCH * (1 if src == "u8" else 4)
        acc = {"units": 0.0, "d2h": 0, "h2d": 0.0}

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
            _, hostres, nbytes, counts, read = pending
            read.synchronize()  # the step's results are on the host
            acc["d2h"] = nbytes
            acc["units"] += wl.host_gaze_steps(hostres)
            acc["h2d"] += float(counts.sum()) * tile_bytes

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
        acc.update(units=0.0, d2h=0, h2d=0.0)
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        gather.TIMING = gather.LaunchTimer() if args.e2e_timing else None
        t0.record()
        pipeline(args.warmup, e2e_steps)
        t1.record()
        barrier()
        ems = max_over_ranks(t0.elapsed_time(t1), device)
        e2e_tags = {}
        for tag, n_items, ev_a, ev_b in (gather.TIMING.records if gather.TIMING else []):
            e2e_tags.setdefault(tag, []).append((n_items, ev_a.elapsed_time(ev_b)))
        gather.TIMING = None
        how = "read in place by the gather kernels (zero-copy over PCIe: only glimpsed tiles move, each once)"
        h2d_per_step = acc["h2d"] / e2e_steps
        result["e2e"] = {
            "value": sum_over_ranks(acc["units"], device) / (ems / 1e3), "unit": "gaze-steps/s",
            "h2d_bytes_per_step": int(h2d_per_step), "d2h_bytes_per_step": int(acc["d2h"]), "steps": e2e_steps,
            "ms_per_step": ems / e2e_steps,
            **({"gather_by_tag": {k_: {"items": round(sum(n for n, _ in v) / len(v), 1),
                                       "ms": round(sum(m for _, m in v) / len(v), 3)} for k_, v in e2e_tags.items()}}
               if e2e_tags else {}),
            "host_buffers": f"pinned {src} images ({full} bytes addressed on the host, {n_distinct} distinct images "
                            f"per rank at every GPU count), {how}"}
        # PCIe side: bytes that crossed per second of the e2e arm vs a pinned cudaMemcpyAsync measured in the same
        # run with every rank copying at once
        memcpy_gbs = ctx["pcie_memcpy_gbs"]
        got = h2d_per_step / (ems / e2e_steps / 1e3) / 1e9
        result["pcie"] = {"bound": "pcie", "achieved": round(got, 2), "peak": round(memcpy_gbs, 2), "unit": "GB/s",
                          "frac": round(got / memcpy_gbs, 4) if memcpy_gbs else None,
                          "peak_source": f"pinned host->device cudaMemcpyAsync of 1 GiB x 4 in this run, all {world} "
                                         "rank(s) copying concurrently (per-rank figure of this rank)",
                          "aggregate_peak": round(sum_over_ranks(memcpy_gbs, device), 1)}
        del host
    elif not args.no_e2e:
        result["e2e"] = None
    result["_wl"] = wl
    return result, cls, batch, src


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="reinforce", choices=sorted(WORKLOADS))
    ap.add_argument("--also", default="auto",
                    help="secondary workloads measured in the same run and nested under \"also\": auto = the other two "
                         "BASELINE configs when the primary is the default one, none, or a comma-separated list")
    ap.add_argument("--batch", type=int, default=0, help="episodes per GPU (default: the BASELINE config's)")
    ap.add_argument("--src", default="u8", choices=["f32", "u8"],
                    help="image dtype of the supervised workload: u8 = uint8 images normalised by the gather (x/255 "
                         "like ToTensor, bit-identical crops; SURVEY 8d's primary synthetic input), f32 = pre-normalised "
                         "float32 images as the reference's dataset hands them over")
    ap.add_argument("--cpu-sample", type=int, default=0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-pool", type=int, default=0,
                    help="distinct pinned host images per rank in the supervised e2e arm (default: all 256, at every "
                         "GPU count)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-timing", action="store_true", help="also time every gather of the e2e arm (diagnostics)")
    ap.add_argument("--no-clocks", action="store_true", help="do not sample nvidia-smi (A/B of the sampler's own cost)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    from jolineedle_b200.sharding import dist_env

    rank, local_rank, world = dist_env()
    if args.impl == "reference":
        return reference_arm(args, rank, world)

    import torch.distributed as dist

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

    # nvidia-smi is started here, seconds before the timed region: its NVML start-up stalls CUDA calls
    clocks = ClockSampler(local_rank)
    if not args.no_clocks and rank == 0:  # rank 0's GPU stands for the box: one poller, not one per rank
        clocks.__enter__()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak, peak_kind = (peaks["hbm_gbs"], "measured") if "hbm_gbs" in peaks else (6650.0, "fallback")
    ctx = {"rank": rank, "world": world, "device": device, "barrier": barrier, "clocks": clocks, "peak": peak,
           "peak_kind": peak_kind, "pcie_memcpy_gbs": None}
    if args.also == "auto":
        also = [w for w in ("supervised", "aerial") if w != args.workload] if args.workload == "reinforce" else []
    elif args.also in ("none", ""):
        also = []
    else:
        also = [w for w in args.also.split(",") if w in WORKLOADS and w != args.workload]
    try:
        if not args.no_e2e:
            ctx["pcie_memcpy_gbs"] = measure_pcie(device, barrier)
        main_res, cls, batch, src = run_workload(args.workload, args, ctx, primary=True)
        wl = main_res.pop("_wl")
        # ---- CPU baseline beside it (rank 0, N=1 only): before the secondary workloads take the memory ----
        cpu = None
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            sample = args.cpu_sample or cls.cpu_default_sample
            v, _, reps, secs = cpu_arm(wl, sample, 0, warmup=1, budget_s=12.0)  # bounded: ~10-30 s of CPU work
            kind = reference_kind()
            cpu = {"value": v, "unit": "gaze-steps/s", "cores": torch.get_num_threads(), "kind": kind,
                   "sample": cpu_description(kind, reps, sample, secs)}
        del wl
        torch.cuda.empty_cache()
        also_res = {}
        for name in also:
            res, acls, abatch, asrc = run_workload(name, args, ctx, primary=False)
            res.pop("_wl")
            res["config"] = config_of(acls, abatch, asrc, world)
            res["metric"], res["unit"] = "gaze_steps_per_sec", "gaze-steps/s"
            also_res[acls.key] = res
            torch.cuda.empty_cache()
    finally:
        clocks.close()

    if rank == 0:
        line = {
            "metric": "gaze_steps_per_sec", "value": main_res.pop("value"), "unit": "gaze-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": main_res.pop("ms_per_step"),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_of(cls, batch, src, world),
            "roofline": main_res.pop("roofline"), "cpu_baseline": cpu, "e2e": main_res.pop("e2e", None),
            "gpu_launches": main_res.pop("gpu_launches"), "clocks": main_res.pop("clocks"),
            **main_res, "also": also_res,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
