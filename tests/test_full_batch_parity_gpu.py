"""Oracle parity at the FULL batch sizes of the BASELINE configs (the benchmark's own workloads):

  cfg 3  reinforce, 1024 episodes, 2240x2688, P=448, T=20, STOP          (all 20 steps)
  cfg 2  supervised, 256 images, 2240x2688, P=448, T=8, binomial 0-3
  cfg 4  aerial, 8192x8192, P=256 (32x32 grid), T=32, augment-translate  (32 episodes)

Every integer / float output of every episode is compared with ``oracle/`` (which runs without pixels for
that: 1024 LARD images are 74 GB as float32 on the host); the crops are compared bit for bit on a random 5 %
of the episodes, for which the oracle gets the real pixels downloaded from the GPU.
"""
import random

import numpy as np
import pytest
import torch

import bench
from oracle.gaze_oracle import GazeOracle, returns_oracle, translate_oracle
from oracle.traj_oracle import TrajectoryOracle, collate_oracle

pytestmark = pytest.mark.gpu


def _subset(n, frac, seed):
    k = max(2, int(round(n * frac)))
    return np.sort(np.random.default_rng(seed).choice(n, size=k, replace=False))


def _rl_full_batch(workload_cls, batch, sample_frac, seed):
    from jolineedle_b200.env.general_env import NeedleGeneralEnv
    from jolineedle_b200.reinforce import rollout_tail

    dev = torch.device("cuda", 0)
    wl = workload_cls(batch, 0, dev, "u8")
    wl.to_device()
    b, P, T = batch, wl.PATCH, wl.T
    boxes = wl.boxes.numpy()
    env = NeedleGeneralEnv(wl.images, wl.boxes_dev, P, T, 1, stop_enabled=True, normalize=True, history=True,
                           translate=wl.translate_dev)
    full = GazeOracle((b, 3, wl.h, wl.w), boxes, P, T, 1, True)  # no pixels: integer / float outputs only
    idx = _subset(b, sample_frac, seed)
    sub_images = wl.images[torch.from_numpy(idx).to(dev)].cpu().float() / 255  # ToTensor values of the sampled images
    if wl.translate is not None:
        sub_images = translate_oracle(sub_images, wl.translate.numpy()[idx])
    sub = GazeOracle(sub_images, boxes[idx], P, T, 1, True)
    assert np.array_equal(env.bbox_masks.cpu().numpy(), full.bbox_masks)

    torch.manual_seed(seed)
    _, i_o = full.reset()
    torch.manual_seed(seed)
    _, i_e = env.reset()
    assert np.array_equal(i_e["positions"].cpu().numpy(), i_o["positions"])
    crops_sub = [sub.reset(i_o["positions"][idx])[0]]
    rng = np.random.default_rng(seed)
    rew, term = [], []
    for t in range(T):
        a = rng.integers(0, 9, size=b).astype(np.int64)
        o = full.step(a)
        e = env.step(torch.from_numpy(a).to(dev))
        assert np.array_equal(e[4]["positions"].cpu().numpy(), o[4]["positions"]), t
        assert np.array_equal(e[1].cpu().numpy(), o[1]), t          # rewards, bit-exact float32
        assert np.array_equal(e[2].cpu().numpy(), o[2]) and np.array_equal(e[3].cpu().numpy(), o[3]), t
        crops_sub.append(sub.step(a[idx])[0])
        rew.append(torch.from_numpy(o[1])); term.append(torch.from_numpy(o[2]))
    assert np.array_equal(env.visited_patches.cpu().numpy(), full.visited)
    assert np.array_equal(env.prop_patches_found.cpu().numpy(), full.prop_patches_found())
    assert np.array_equal(env.prop_bboxes_found.cpu().numpy(), full.prop_bboxes_found())
    assert np.array_equal(env.terminated.cpu().numpy(), full.terminated())
    # returns tail on the env's own step-major rings
    r_tn, t_tn, _ = env.rollout_buffers()
    tail = rollout_tail(r_tn, t_tn)
    masks = torch.cat([torch.ones((b, 1), dtype=torch.bool), ~torch.stack(term, 1)], dim=1)
    want, lm = returns_oracle(torch.stack(rew, 1), masks)
    assert torch.equal(tail["returns"].cpu(), want) and torch.equal(tail["logit_masks"].cpu(), lm)
    assert torch.equal(tail["masks"].cpu(), masks) and torch.equal(tail["rewards"].cpu(), torch.stack(rew, 1))
    # crops of the sampled episodes, every step, bit for bit
    hist = env.patch_history()[torch.from_numpy(idx).to(dev)].cpu()  # [k, T + 1, C, P, P]
    want_hist = torch.cat(crops_sub, dim=1)
    assert hist.shape == want_hist.shape and torch.equal(hist, want_hist)
    env.check_status()


def test_cfg3_full_batch_1024_episodes_all_steps():
    _rl_full_batch(bench.ReinforceWorkload, 1024, 0.05, seed=3)


def test_cfg4_translate_32_episodes_all_steps():
    _rl_full_batch(bench.AerialWorkload, 32, 0.07, seed=4)


def test_cfg2_full_batch_256_images():
    from jolineedle_b200.env.simple_env import generate_trajectories

    dev = torch.device("cuda", 0)
    b = 256
    wl = bench.SupervisedWorkload(b, 0, dev, "u8")
    wl.to_device()
    P, T = bench.P, wl.T
    idx = set(_subset(b, 0.05, 9).tolist())
    blank = torch.zeros((), dtype=torch.float32).expand(3, wl.h, wl.w)  # shape only: no pixels, no memory
    seeds = wl.seeds(0)
    random.seed(77)
    samples, det_rows = [], []
    for i in range(b):
        img = (wl.images[i].cpu().float() / 255) if i in idx else blank
        boxes = [((y1, x1), (y2, x2)) for (x1, y1, x2, y2) in wl.raw_boxes[i]]
        s = TrajectoryOracle(img, P, boxes, seeds[i]).generate_sample(T, wl.KMIN, wl.KMAX, wl.BINOMIAL, None)
        s["class_id"] = torch.tensor(0, dtype=torch.long)
        det_rows.append(s["patches_yolox"].shape[0])
        samples.append(s)
    want = collate_oracle(samples)
    random.seed(77)
    got = generate_trajectories({"image": wl.images, "bboxes": wl.bboxes, "class_id": wl.class_ids}, P, T, wl.KMIN,
                                wl.KMAX, binomial_keypoints=wl.BINOMIAL, seeds=seeds, normalize=True, check=True)
    assert set(got) == set(want)
    for k in want:
        assert got[k].dtype == want[k].dtype and tuple(got[k].shape) == tuple(want[k].shape), k
        if k not in ("patches", "patches_yolox"):
            assert torch.equal(got[k].cpu(), want[k]), k
    starts = np.concatenate([[0], np.cumsum(det_rows)])
    for i in sorted(idx):
        assert torch.equal(got["patches"][i].cpu(), want["patches"][i]), i
        lo, hi = int(starts[i]), int(starts[i + 1])
        assert torch.equal(got["patches_yolox"][lo:hi].cpu(), want["patches_yolox"][lo:hi]), i
