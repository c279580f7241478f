#!/usr/bin/env python
"""Generate the committed golden fixtures by running the UNMODIFIED reference env.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py            # rewrite tests/golden/*.npz
    python tests/golden/make_golden.py --fuzz 300 # + randomized reference-vs-oracle cross-check

The fixtures hold inputs and the reference's outputs for the scenarios listed below.  The
CPU test-suite checks ``oracle/`` against them, the GPU test-suite checks the CUDA path
against them, and neither needs the reference at run time.
"""
import argparse
import hashlib
import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

from _ref_loader import load_reference  # noqa: E402

ge, se, co, ut = load_reference()


def synth_u8(b, c, h, w, salt=0):
    """Closed-form deterministic uint8 image batch (no RNG stream dependence)."""
    bb, cc, yy, xx = np.meshgrid(np.arange(b), np.arange(c), np.arange(h), np.arange(w), indexing="ij")
    v = xx * 131 + yy * 71 + cc * 29 + bb * 17 + (xx * yy) % 251 + ((xx ^ yy) * 7) % 13 + salt * 101
    return (v % 256).astype(np.uint8)


def sha(t: torch.Tensor) -> str:
    return hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest()


def to_f32(u8: np.ndarray) -> torch.Tensor:
    # ToTensor semantics (dataset.py:240): uint8 -> float32 / 255
    return torch.from_numpy(u8).float() / 255


# ------------------------------------------------------------------------------------------
# general env
# ------------------------------------------------------------------------------------------
def run_general(name, u8, bboxes, P, T, stop_enabled, actions, start=None, seed=None, keep_steps=(0, 1)):
    images = to_f32(u8)
    env = ge.NeedleGeneralEnv(images, torch.from_numpy(bboxes), P, T, 1, stop_enabled)
    if start is None:
        torch.manual_seed(seed)
        patches, infos = env.reset()
    else:
        patches, infos = env.reset(torch.from_numpy(start))
    out = {
        "u8": u8, "bboxes": bboxes, "P": P, "T": T, "stop_enabled": int(stop_enabled),
        "actions": actions, "seed": -1 if seed is None else seed,
        "bbox_masks": env.bbox_masks.numpy(),
    }
    pos = [infos["positions"].numpy().copy()]
    vis = [env.visited_patches.numpy().copy()]
    prop = [env.prop_patches_found.numpy().copy()]
    propb = [env.prop_bboxes_found.numpy().copy()]
    shas = [sha(patches)]
    kept = {0: patches.numpy().copy()} if 0 in keep_steps else {}
    rew, term, trunc = [], [], []
    for t in range(actions.shape[0]):
        patches, r, te, tr, infos = env.step(torch.from_numpy(actions[t]))
        pos.append(infos["positions"].numpy().copy())
        vis.append(env.visited_patches.numpy().copy())
        prop.append(env.prop_patches_found.numpy().copy())
        propb.append(env.prop_bboxes_found.numpy().copy())
        rew.append(r.numpy().copy()); term.append(te.numpy().copy()); trunc.append(tr.numpy().copy())
        shas.append(sha(patches))
        if t + 1 in keep_steps:
            kept[t + 1] = patches.numpy().copy()
    out.update(
        positions=np.stack(pos), visited=np.stack(vis), prop_patches=np.stack(prop), prop_bboxes=np.stack(propb),
        rewards=np.stack(rew), terminated=np.stack(term), truncated=np.stack(trunc),
        patch_sha=np.array(shas), kept_steps=np.array(sorted(kept)),
        kept_patches=np.stack([kept[k] for k in sorted(kept)]),
    )
    assert out["rewards"].dtype == np.float32 and out["positions"].dtype == np.int64
    return {f"{name}/{k}": v for k, v in out.items()}


def random_boxes(rng, b, n, h, w, max_side, zero_rows=True):
    boxes = np.zeros((b, n, 4), dtype=np.int64)
    for i in range(b):
        for j in range(n):
            bw, bh = rng.integers(2, max_side, size=2)
            x1 = rng.integers(0, max(w - bw, 1)); y1 = rng.integers(0, max(h - bh, 1))
            boxes[i, j] = (x1, y1, min(x1 + bw, w - 1), min(y1 + bh, h - 1))
        if zero_rows and n > 1 and i % 3 == 2:
            boxes[i, n - 1] = 0  # zero-padded row from padded_collate_fn (dataset.py:338-341)
    return boxes


def general_fixtures():
    rng = np.random.default_rng(20240601)
    fx = {}
    # B: 5x6 grid, STOP enabled, 9 actions, T=20 (cfg-3 shaped, shrunk to P=16)
    u8 = synth_u8(6, 3, 80, 96, salt=1)
    boxes = random_boxes(rng, 6, 3, 80, 96, 30)
    boxes[0, 0] = (10, 10, 40, 40)      # spans 3x3 patches
    boxes[1, 1] = (90, 70, 120, 100)    # sticks out of the image -> clamped (kornia to_mask)
    actions = rng.integers(0, 9, size=(20, 6)).astype(np.int64)
    actions[5:, 2] = rng.integers(0, 8, size=15)  # episode 2 never stops after step 5 ...
    actions[3, 2] = 8                              # ... but stops once at step 3 (sticky)
    start = np.stack([rng.integers(0, 5, size=6), rng.integers(0, 6, size=6)], axis=1).astype(np.int64)
    fx.update(run_general("stop9", u8, boxes, 16, 20, True, actions, start=start, keep_steps=(0, 1, 20)))
    # C: no STOP, 8 actions, T=8 (cfg-1 shaped), greedy-ish actions so some episodes terminate
    u8 = synth_u8(4, 3, 80, 80, salt=2)
    boxes = random_boxes(rng, 4, 2, 80, 80, 20)
    actions = rng.integers(0, 8, size=(8, 4)).astype(np.int64)
    start = np.array([[by // 16, max(bx // 16 - 1, 0)] for bx, by in boxes[:, 0, :2]], dtype=np.int64)
    actions[0, :] = 1  # RIGHT: walk into the first box's patch
    fx.update(run_general("nostop8", u8, boxes, 16, 8, False, actions, start=start, keep_steps=(0, 8)))
    # D: random start positions drawn by reset() from the CPU default generator
    actions = rng.integers(0, 9, size=(4, 6)).astype(np.int64)
    fx.update(run_general("randstart", synth_u8(6, 3, 80, 96, salt=3), random_boxes(rng, 6, 2, 80, 96, 30),
                          16, 6, True, actions, seed=777, keep_steps=(0,)))
    # E: 32x32 grid (1024-bit bitmaps, cfg-4 shaped: T=32), P=8
    u8 = synth_u8(3, 3, 256, 256, salt=4)
    boxes = random_boxes(rng, 3, 4, 256, 256, 60)
    actions = rng.integers(0, 9, size=(32, 3)).astype(np.int64)
    fx.update(run_general("grid32", u8, boxes, 8, 32, True, actions, seed=31337, keep_steps=(0, 32)))
    # F: 40x36 grid (more than 1024 patches -> several bitmap words per lane), P=4
    u8 = synth_u8(2, 3, 160, 144, salt=5)
    boxes = random_boxes(rng, 2, 5, 160, 144, 50, zero_rows=False)
    actions = rng.integers(0, 8, size=(12, 2)).astype(np.int64)
    fx.update(run_general("grid40", u8, boxes, 4, 12, False, actions, seed=5, keep_steps=(0,)))
    # G: uint8 images handed to the env as they are (crops keep dtype and values)
    u8 = synth_u8(3, 3, 64, 96, salt=6)
    env = ge.NeedleGeneralEnv(torch.from_numpy(u8), torch.from_numpy(random_boxes(rng, 3, 1, 64, 96, 30)), 32, 4, 1, False)
    p0, _ = env.reset(torch.tensor([[0, 0], [1, 2], [1, 1]]))
    p1 = env.step(torch.tensor([1, 0, 7]))[0]
    fx.update({"u8env/u8": u8, "u8env/p0": p0.numpy(), "u8env/p1": p1.numpy()})
    return fx


def detection_fixtures():
    rng = np.random.default_rng(99)
    fx = {}
    # KAT of the reference's tests/test_map.py:9-34 (zeros image, P=448) -- recorded, not trusted blindly
    env = ge.NeedleGeneralEnv(torch.zeros((1, 3, 1792, 2240)), torch.tensor([[[410, 410, 500, 500], [1500, 1500, 1600, 1600]]]), 448, 20, 1)
    fx["kat_map/targets0"] = env.get_detection_targets()[0].numpy()
    for name, n in (("det1", 1), ("det3", 3)):
        u8 = synth_u8(4, 3, 80, 96, salt=10 + n)
        boxes = random_boxes(rng, 4, n, 80, 96, 40, zero_rows=(n > 1))
        boxes[0, 0] = (14, 14, 50, 34)  # splits over 4x2 patches
        env = ge.NeedleGeneralEnv(to_f32(u8), torch.from_numpy(boxes), 16, 8, 1)
        local, present = env.parse_bboxes(env.bboxes)
        targets = env.get_detection_targets()
        torch.manual_seed(4242 + n)
        patches, tb = env.get_detection_batch(sample_neg=2)
        fx.update({
            f"{name}/u8": u8, f"{name}/bboxes": boxes, f"{name}/P": 16,
            f"{name}/local": local.numpy(), f"{name}/present": present.numpy(),
            f"{name}/targets_cat": torch.cat(targets).numpy(),
            f"{name}/targets_len": np.array([len(t) for t in targets]),
            f"{name}/seed": 4242 + n, f"{name}/batch_patches_sha": np.array(sha(patches)),
            f"{name}/batch_patches": patches.numpy(), f"{name}/batch_boxes": tb.numpy(),
        })
    return fx


# ------------------------------------------------------------------------------------------
# simple env
# ------------------------------------------------------------------------------------------
def mk_bboxes(raw):
    return [ut.BBox(ut.Position(y1, x1), ut.Position(y2, x2)) for (x1, y1, x2, y2) in raw]


SAMPLE_KEYS = ("patches", "current_actions", "next_actions", "positions", "masks", "labels", "local_bboxes",
               "patches_yolox", "bboxes_yolox")


def run_simple(u8_img, raw_boxes, P, T, kmin, kmax, binomial, seed, position):
    random.seed(seed * 7 + 1)
    env = se.NeedleSimpleEnv(to_f32(u8_img), P, mk_bboxes(raw_boxes), seed=seed)
    pos = None if position is None else ut.Position(*position)
    s = env.generate_sample(T, kmin, kmax, binomial_keypoints=binomial, position=pos)
    return s, env


def simple_fixtures():
    rng = np.random.default_rng(424242)
    fx, meta = {}, []
    cases = []
    # (grid_h, grid_w, P, T): cfg-1/2 shaped (5x5 / 5x6, T=8), a long one, and a wide grid with short T
    for gh, gw, P, T in ((5, 5, 16, 8), (5, 6, 16, 8), (5, 6, 16, 20), (8, 9, 8, 6)):
        for rep in range(8):
            cases.append((gh, gw, P, T, rep))
    for idx, (gh, gw, P, T, rep) in enumerate(cases):
        h, w = gh * P, gw * P
        n = int(rng.integers(0, 4)) if rep != 0 else 0
        raw = []
        for _ in range(n):
            bw, bh = (int(v) for v in rng.integers(2, 3 * P, size=2))
            x1, y1 = int(rng.integers(0, w - 2)), int(rng.integers(0, h - 2))
            raw.append((x1, y1, min(x1 + bw, w - 1), min(y1 + bh, h - 1)))
        binomial = bool(rep % 2)
        kmin, kmax = ((0, 3), (0, 0), (2, 2), (1, 4))[rep % 4]
        position = None if rep % 3 else (int(rng.integers(0, gh)), int(rng.integers(0, gw)))
        seed = 1000 + idx
        u8 = synth_u8(1, 3, h, w, salt=idx)[0]
        s, env = run_simple(u8, raw, P, T, kmin, kmax, binomial, seed, position)
        name = f"s{idx:02d}"
        fx[f"{name}/u8"] = u8
        fx[f"{name}/raw_boxes"] = np.array(raw, dtype=np.int64).reshape(-1, 4)
        fx[f"{name}/cfg"] = np.array([P, T, kmin, kmax, int(binomial), seed, -1 if position is None else position[0],
                                      -1 if position is None else position[1]], dtype=np.int64)
        fx[f"{name}/bbox_patches"] = np.array(sorted(env.bbox_patches), dtype=np.int64).reshape(-1, 2)
        for k in SAMPLE_KEYS:
            fx[f"{name}/{k}"] = s[k].numpy()
        meta.append(name)
    fx["names"] = np.array(meta)
    # one collated batch (different box counts -> padding), cfg-1 batch size
    samples = []
    for j, idx in enumerate((1, 2, 3, 5)):
        name = f"s{idx:02d}"
        cfg = fx[f"{name}/cfg"]
        raw = [tuple(r) for r in fx[f"{name}/raw_boxes"].tolist()]
        pos = None if cfg[6] < 0 else (int(cfg[6]), int(cfg[7]))
        s, _ = run_simple(fx[f"{name}/u8"], raw, int(cfg[0]), int(cfg[1]), int(cfg[2]), int(cfg[3]), bool(cfg[4]), int(cfg[5]), pos)
        s["class_id"] = torch.tensor(j, dtype=torch.long)
        samples.append(s)
    batch = se.NeedleSimpleEnv.collate_fn(samples)
    for k, v in batch.items():
        fx[f"collate/{k}"] = v.numpy()
    fx["collate/members"] = np.array([1, 2, 3, 5])
    return fx


def returns_fixtures():
    g = torch.Generator().manual_seed(2024)
    fx = {}
    for name, (b, t) in (("r20", (16, 20)), ("r7", (5, 7)), ("r1", (3, 1))):
        rewards = torch.randn((b, t), generator=g) * 3
        rewards[::2] = (torch.randint(0, 2, (rewards[::2].shape), generator=g).float() - 1 / 20)  # env-like values
        stop_at = torch.randint(0, t + 2, (b,), generator=g)
        terminated = torch.arange(t)[None, :] >= stop_at[:, None]
        masks = torch.cat([torch.ones((b, 1), dtype=torch.bool), ~terminated], dim=1)
        # verbatim data flow of reinforce.py:191-202
        logit_masks = torch.roll(masks[:, 1:], shifts=1, dims=(1,))
        logit_masks[:, 0] = True
        br = torch.flip(rewards, dims=(1,)); bm = torch.flip(logit_masks, dims=(1,))
        ret = torch.flip(torch.cumsum(br * bm, dim=1), dims=(1,))
        fx.update({f"{name}/rewards": rewards.numpy(), f"{name}/masks": masks.numpy(),
                   f"{name}/logit_masks": logit_masks.numpy(), f"{name}/returns": ret.numpy()})
    return fx


def norm_fixture():
    # every uint8 value through ToTensor's `/255` in fp32 (dataset.py:240)
    return {"u8_over_255": (torch.arange(256, dtype=torch.uint8).float() / 255).numpy()}


# ------------------------------------------------------------------------------------------
# randomized reference-vs-oracle cross-check (not stored)
# ------------------------------------------------------------------------------------------
def fuzz(n_rounds):
    from oracle.gaze_oracle import GazeOracle, bbox_patch_mask_raster
    from oracle.traj_oracle import TrajectoryOracle

    rng = np.random.default_rng(7)
    bad = 0
    for it in range(n_rounds):
        gh, gw, P = int(rng.integers(2, 9)), int(rng.integers(2, 9)), int(rng.choice([4, 8, 16]))  # reflect-pad in the reference needs grid >= 2
        h, w, b, n = gh * P, gw * P, int(rng.integers(1, 5)), int(rng.integers(1, 4))
        T, stop = int(rng.integers(1, 12)), bool(rng.integers(0, 2))
        u8 = synth_u8(b, 3, h, w, salt=it)
        boxes = random_boxes(rng, b, n, h, w, 3 * P + 2)
        ref = ge.NeedleGeneralEnv(to_f32(u8), torch.from_numpy(boxes), P, T, 1, stop)
        orc = GazeOracle(to_f32(u8), boxes, P, T, 1, stop)
        ok = np.array_equal(ref.bbox_masks.numpy(), orc.bbox_masks)
        ok &= np.array_equal(orc.bbox_masks, bbox_patch_mask_raster(boxes, h, w, P))
        torch.manual_seed(it); p_ref, i_ref = ref.reset()
        torch.manual_seed(it); p_orc, i_orc = orc.reset()
        ok &= torch.equal(p_ref, p_orc) and np.array_equal(i_ref["positions"].numpy(), i_orc["positions"])
        for t in range(T):
            a = rng.integers(0, 9 if stop else 8, size=b).astype(np.int64)
            o_ref = ref.step(torch.from_numpy(a)); o_orc = orc.step(a)
            ok &= torch.equal(o_ref[0], o_orc[0])
            ok &= np.array_equal(o_ref[1].numpy(), o_orc[1]) and o_orc[1].dtype == np.float32
            ok &= np.array_equal(o_ref[2].numpy(), o_orc[2]) and np.array_equal(o_ref[3].numpy(), o_orc[3])
            ok &= np.array_equal(o_ref[4]["positions"].numpy(), o_orc[4]["positions"])
            ok &= np.array_equal(ref.prop_patches_found.numpy(), orc.prop_patches_found())
        tg_ref = ref.get_detection_targets() if boxes.any(axis=2).any(axis=1).all() else None
        if tg_ref is not None:
            tg_orc = orc.detection_targets()
            ok &= all(np.array_equal(a.numpy(), b_) for a, b_ in zip(tg_ref, tg_orc))
        torch.manual_seed(it + 1); db_ref = ref.get_detection_batch(sample_neg=1)
        torch.manual_seed(it + 1); db_orc = orc.detection_batch(sample_neg=1)
        ok &= torch.equal(db_ref[0], db_orc[0]) and np.array_equal(db_ref[1].numpy(), db_orc[1])
        # simple env
        raw = [tuple(int(v) for v in boxes[0, j]) for j in range(n) if boxes[0, j].any()]
        binomial = bool(rng.integers(0, 2)); kmin = int(rng.integers(0, 3)); kmax = kmin + int(rng.integers(0, 3))
        pos = None if rng.integers(0, 2) else (int(rng.integers(0, gh)), int(rng.integers(0, gw)))
        s_ref, e_ref = run_simple(u8[0], raw, P, T, kmin, kmax, binomial, it, pos)
        random.seed(it * 7 + 1)
        e_orc = TrajectoryOracle(to_f32(u8[0]), P, [((y1, x1), (y2, x2)) for (x1, y1, x2, y2) in raw], seed=it)
        s_orc = e_orc.generate_sample(T, kmin, kmax, binomial, pos)
        for k in SAMPLE_KEYS:
            ok &= torch.equal(s_ref[k], s_orc[k])
        ok &= sorted(e_ref.bbox_patches) == sorted(e_orc.bbox_patches)
        if not ok:
            bad += 1
            print("MISMATCH in fuzz round", it)
    print(f"fuzz: {n_rounds} rounds, {bad} mismatching")
    return bad


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--fuzz", type=int, default=0)
    ap.add_argument("--no-write", action="store_true")
    args = ap.parse_args()
    if not args.no_write:
        np.savez_compressed(os.path.join(HERE, "general_env.npz"), **general_fixtures())
        np.savez_compressed(os.path.join(HERE, "detection.npz"), **detection_fixtures())
        np.savez_compressed(os.path.join(HERE, "simple_env.npz"), **simple_fixtures())
        np.savez_compressed(os.path.join(HERE, "returns.npz"), **returns_fixtures())
        np.savez_compressed(os.path.join(HERE, "norm.npz"), **norm_fixture())
        for f in sorted(os.listdir(HERE)):
            if f.endswith(".npz"):
                print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")
    if args.fuzz:
        sys.exit(1 if fuzz(args.fuzz) else 0)


if __name__ == "__main__":
    main()
