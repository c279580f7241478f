"""The reference's own CALLER code, unmodified, driving the drop-in env.

``ReinforceTrainer.rollout`` / ``sample_from_logits`` (src/reinforce.py:73-90,108-215, including the
``do_detection`` branch :141-146,162-167) and ``SupervisedTrainer.create_env`` / ``generate_trajectories``
(src/supervised.py:85-136) are read out of the reference's source files (the modules themselves cannot be
imported: they pull yolox, visdom, torchmetrics ...), compiled as they are into a stand-in class with a stub
policy / stub detector, and run twice: on the reference's CPU env and on the CUDA drop-in.  Everything the env
contributes to the results must agree bit for bit.

Needs the reference copy in ``baseline/_ref`` (made by ``__graft_entry__.build()``; it travels to the GPU box).
"""
import random
from types import SimpleNamespace
from typing import Dict, List, Optional, Tuple  # noqa: F401  (names the reference's annotations use)

import numpy as np
import pytest
import torch
from torch.distributions import Categorical  # noqa: F401

from baseline import ref_env
from helpers import random_boxes, synth_u8, to_f32

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_env.available(), reason="no reference copy (baseline/_ref)")]


def compile_methods(rel_path, class_name, names, extra_globals):
    src = ref_env.function_source(rel_path, class_name, names)
    assert all(f"def {n}(" in src for n in names), f"{names} not found in {rel_path}"
    scope = {"torch": torch, "Dict": Dict, "List": List, "Optional": Optional, "Tuple": Tuple,
             "Categorical": Categorical, **extra_globals}
    exec(compile(src, f"<reference {rel_path}>", "exec"), scope)  # the reference's text, unmodified
    return {n: scope[n] for n in names}


def policy(patches, actions, classes, positions, embeddings):
    """Stub for the GPT: logits that depend on the last position, the step and the episode -- integer
    arithmetic, so the CPU and the GPU runs take the same actions; no ties for argmax."""
    b, t = actions.shape
    y, x = positions[:, -1, 0], positions[:, -1, 1]
    a = torch.arange(9, device=actions.device)
    ep = torch.arange(b, device=actions.device)
    score = ((y[:, None] * 7 + x[:, None] * 13 + t * 5 + ep[:, None] * 3 + a[None, :] * 4) % 11).float()
    score = score + a[None, :].float() * 0.01
    assert patches.shape[:2] == (b, t) and positions.shape[:2] == (b, t)  # the history the model consumes
    return score[:, None, :].expand(b, t, 9), embeddings


def bit_checksum(x):
    """Order-independent, exact checksum of float32 data (the same on CPU and GPU)."""
    return int(x.contiguous().view(torch.int32).long().sum().item())


class StubDetector:
    """Stands for the YOLOX head of the do_detection branch: returns one 'prediction' per episode that
    fingerprints the crops it was shown."""

    def __init__(self, batch_size):
        self.batch_size, self.seen = batch_size, []

    def __call__(self, patches, targets):
        self.seen.append(tuple(patches.shape))
        return [bit_checksum(patches)] * self.batch_size, None, {"total_loss": 0.0}


@pytest.mark.parametrize("do_detection", [False, True])
@pytest.mark.parametrize("stop_enabled", [True, False])
def test_reference_rollout_source_runs_on_the_drop_in(do_detection, stop_enabled):
    from jolineedle_b200.env.general_env import NeedleGeneralEnv as Ours

    ge = ref_env.load()[0]
    b, P, gh, gw, T = 6, 32, 4, 5, 9
    rng = np.random.default_rng(12)
    images = to_f32(synth_u8(b, 3, gh * P, gw * P, salt=3))
    boxes = torch.from_numpy(random_boxes(rng, b, 2, gh * P, gw * P, 2 * P))
    results = {}
    for name, env_cls, dev in (("reference", ge.NeedleGeneralEnv, torch.device("cpu")),
                               ("ours", Ours, torch.device("cuda", 0))):
        methods = compile_methods("src/reinforce.py", "ReinforceTrainer", ["sample_from_logits", "rollout"],
                                  {"NeedleGeneralEnv": env_cls})
        Trainer = type("Trainer", (), methods)
        trainer = Trainer()
        trainer.device, trainer.model = dev, policy
        detector = StubDetector(b)
        trainer.yolox_model = lambda d=detector: d
        env = env_cls(images.to(dev), boxes.to(dev), P, T, 1, stop_enabled)
        torch.manual_seed(5)  # start positions: CPU generator in both envs
        out = trainer.rollout(env, do_detection=do_detection, sample_actions=False)
        out["prop_patches_found"], out["terminated"] = env.prop_patches_found, env.terminated
        results[name] = (out, detector)
    ref, det_ref = results["reference"]
    got, det_got = results["ours"]
    assert set(ref) == set(got)
    for k in ("rewards", "masks", "logit_masks", "positions", "patches", "prop_patches_found", "terminated"):
        assert ref[k].dtype == got[k].dtype and tuple(ref[k].shape) == tuple(got[k].shape), k
        assert torch.equal(ref[k], got[k].cpu()), k
    # trainer-side float math (the reference's own torch.cumsum / Categorical, here on CUDA instead of the CPU):
    # within the 1e-6 the north star allows for returns
    for k in ("returns", "logprobs", "entropies"):
        assert tuple(ref[k].shape) == tuple(got[k].shape), k
        assert torch.allclose(ref[k], got[k].cpu(), rtol=1e-5, atol=1e-6), k
    # ... while the package's own returns tail (K3, on the env's step-major rings) reproduces the CPU cumsum bit for bit
    from jolineedle_b200.reinforce import rollout_tail

    r_tn, t_tn, _ = env.rollout_buffers()
    assert torch.equal(rollout_tail(r_tn, t_tn)["returns"].cpu(), ref["returns"])
    assert ref["bboxes"] == got["bboxes"] and det_ref.seen == det_got.seen  # detector saw the same crops
    assert len(ref["bboxes"][0]) == ((ref["rewards"].shape[1] + 1) if do_detection else 0)


class CountingRng:
    """np.random.default_rng stand-in for the unseeded envs the trainer builds (supervised.py:88-92)."""

    def __init__(self):
        self.real, self.count = np.random.default_rng, 0

    def __call__(self, seed=None):
        if seed is None:
            seed, self.count = 90_000 + self.count, self.count + 1
        return self.real(seed)


@pytest.mark.parametrize("binomial", [False, True])
def test_reference_generate_trajectories_source_runs_on_the_drop_in(binomial, monkeypatch):
    from jolineedle_b200.env import simple_env as ours
    from jolineedle_b200.utils import BBox as OurBBox, Position as OurPosition

    _, se, _, ut = ref_env.load()
    b, P, T = 5, 32, 8
    rng = np.random.default_rng(31)
    images, raws = [], []
    for i in range(b):
        gh, gw = 4 + i % 2, 5 - i % 2
        images.append(to_f32(synth_u8(1, 3, gh * P, gw * P, salt=20 + i)[0]))
        raw = []
        for _ in range(int(rng.integers(0, 4))):
            bw, bh = (int(v) for v in rng.integers(4, 2 * P, size=2))
            x1, y1 = int(rng.integers(0, gw * P - 4)), int(rng.integers(0, gh * P - 4))
            raw.append((x1, y1, min(x1 + bw, gw * P - 1), min(y1 + bh, gh * P - 1)))
        raws.append(raw)
    config = SimpleNamespace(patch_size=P, max_seq_len=T, min_keypoints=0, max_keypoints=3, binomial_keypoints=binomial)
    outs = {}
    for name, mod, Box, Pos, dev in (("reference", se, ut.BBox, ut.Position, "cpu"),
                                     ("ours", ours, OurBBox, OurPosition, "cuda")):
        methods = compile_methods("src/supervised.py", "SupervisedTrainer", ["create_env", "generate_trajectories"],
                                  {"NeedleSimpleEnv": mod.NeedleSimpleEnv, "Position": Pos, "BBox": Box})
        trainer = type("Trainer", (), methods)()
        trainer.config = config
        batch = {"image": [im.to(dev) for im in images],
                 "bboxes": [[Box(Pos(y1, x1), Pos(y2, x2)) for (x1, y1, x2, y2) in r] for r in raws],
                 "class_id": list(range(b))}
        monkeypatch.setattr(np.random, "default_rng", CountingRng())
        random.seed(3)
        outs[name] = trainer.generate_trajectories(batch)
        monkeypatch.undo()
    ref, got = outs["reference"], outs["ours"]
    assert set(ref) == set(got)
    for k in ref:
        assert ref[k].dtype == got[k].dtype and tuple(ref[k].shape) == tuple(got[k].shape), k
        assert torch.equal(ref[k], got[k].cpu()), k


def test_trajectory_mixin_is_the_batched_path():
    """``TrajectoryMixin`` (what a SupervisedTrainer inherits): same keys, dtypes and shapes as the
    reference's method, consistent with the per-env path it replaces."""
    from jolineedle_b200.env.simple_env import NeedleSimpleEnv
    from jolineedle_b200.supervised import TrajectoryMixin
    from jolineedle_b200.utils import BBox, Position

    P, T, b = 32, 8, 4

    class Trainer(TrajectoryMixin):
        config = SimpleNamespace(patch_size=P, max_seq_len=T, min_keypoints=0, max_keypoints=2, binomial_keypoints=True,
                                 normalize_on_gather=True)
        device = torch.device("cuda", 0)

    images = [torch.from_numpy(synth_u8(1, 3, 4 * P, 5 * P, salt=i)[0]) for i in range(b)]  # uint8, on the HOST
    boxes = [[BBox(Position(10, 12), Position(60, 70))], [], [BBox(Position(0, 0), Position(20, 20))] * 2,
             [BBox(Position(90, 100), Position(127, 159))]]
    trainer = Trainer()
    out = trainer.generate_trajectories({"image": images, "bboxes": boxes, "class_id": [3, 2, 1, 0]})
    assert set(out) == {"patches", "current_actions", "next_actions", "positions", "masks", "labels", "local_bboxes",
                        "class_id", "patches_yolox", "bboxes_yolox"}
    assert tuple(out["patches"].shape) == (b, T, 3, P, P) and out["patches"].dtype == torch.float32
    assert tuple(out["local_bboxes"].shape) == (b, T, 2, 6) and out["class_id"].tolist() == [3, 2, 1, 0]
    assert out["patches"].is_cuda and float(out["masks"][:, 0].min()) == 1.0
    # every recorded slot holds the crop at its recorded position (ToTensor values of the uint8 image)
    for i in range(b):
        for t in range(int(out["masks"][i].sum())):
            y, x = out["positions"][i, t].tolist()
            want = images[i][:, y * P:(y + 1) * P, x * P:(x + 1) * P].float() / 255
            assert torch.equal(out["patches"][i, t].cpu(), want)
    env, class_id = trainer.create_env({"image": images[0], "bboxes": boxes[0], "class_id": 7})
    assert isinstance(env, NeedleSimpleEnv) and class_id == 7 and env.image.is_cuda
