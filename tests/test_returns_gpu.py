"""K3b parity: returns tail against the reference fixtures and the torch-CPU expression."""
import numpy as np
import pytest
import torch

from helpers import load_golden, scenario
from oracle.gaze_oracle import GazeOracle, returns_oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["r20", "r7", "r1"])
def test_returns_match_reference_fixtures(name):
    from jolineedle_b200.reinforce import compute_returns, rollout_tail

    c = scenario(load_golden("returns.npz"), name)
    rewards, masks = torch.from_numpy(c["rewards"]), torch.from_numpy(c["masks"])
    terminated_tn = (~masks[:, 1:]).t().contiguous()
    out = rollout_tail(rewards.t().contiguous().cuda(), terminated_tn.cuda())
    assert np.array_equal(out["returns"].cpu().numpy(), c["returns"])  # bit-exact (tolerance allowed: 1e-6 rel)
    assert np.array_equal(out["logit_masks"].cpu().numpy(), c["logit_masks"])
    assert np.array_equal(out["masks"].cpu().numpy(), c["masks"])
    assert np.array_equal(out["rewards"].cpu().numpy(), c["rewards"])
    got = compute_returns(rewards.cuda(), torch.from_numpy(c["logit_masks"]).cuda())
    assert np.array_equal(got.cpu().numpy(), c["returns"])


def test_returns_random_wide_dynamic_range():
    from jolineedle_b200.reinforce import compute_returns, rollout_tail

    g = torch.Generator().manual_seed(3)
    for b, t in ((1024, 20), (333, 32), (7, 1), (4096, 8)):
        rewards = torch.randn((b, t), generator=g) * torch.exp(torch.randn((b, t), generator=g) * 4)
        stop_at = torch.randint(0, t + 2, (b,), generator=g)
        terminated = torch.arange(t)[None, :] >= stop_at[:, None]
        masks = torch.cat([torch.ones((b, 1), dtype=torch.bool), ~terminated], dim=1)
        want, lm = returns_oracle(rewards, masks)
        out = rollout_tail(rewards.t().contiguous().cuda(), terminated.t().contiguous().cuda())
        assert torch.equal(out["returns"].cpu(), want) and torch.equal(out["logit_masks"].cpu(), lm)
        assert torch.equal(compute_returns(rewards.cuda(), lm.cuda()).cpu(), want)
        # non-contiguous rows (a slice of a wider buffer)
        wide = torch.zeros((b, t + 3)).cuda()
        wide[:, :t] = rewards.cuda()
        assert torch.equal(compute_returns(wide[:, :t], lm.cuda()).cpu(), want)


def test_rollout_driver_matches_oracle_loop():
    """reinforce.py:108-215 with a deterministic stand-in policy: same actions fed to the oracle
    env must give the same rewards / masks / returns / positions / patch history."""
    from helpers import random_boxes, synth_u8, to_f32
    from jolineedle_b200.env.general_env import NeedleGeneralEnv
    from jolineedle_b200.reinforce import rollout

    b, P, gh, gw, T = 6, 64, 4, 5, 12
    rng = np.random.default_rng(0)
    u8 = synth_u8(b, 3, gh * P, gw * P, salt=1)
    boxes = random_boxes(rng, b, 2, gh * P, gw * P, 100)
    images = to_f32(u8)

    def policy(patches, actions, classes, positions, emb):
        # logits depend on the newest crop and position only: deterministic argmax policy
        t = patches.shape[1]
        feat = patches[:, -1].mean(dim=(1, 2, 3)) * 1000 + positions[:, -1].sum(dim=1) * 3 + t
        logits = torch.sin(feat[:, None] * torch.arange(1, 10, device=patches.device)[None, :])
        return logits[:, None, :].expand(-1, t, -1), None

    for history in (True, False):
        env = NeedleGeneralEnv(images.cuda(), torch.from_numpy(boxes), P, T, 1, True, history=history)
        torch.manual_seed(4)
        out = rollout(env, policy, sample_actions=False)
        steps = out["rewards"].shape[1]
        orc = GazeOracle(images, boxes, P, T, 1, True)
        torch.manual_seed(4)
        p, infos = orc.reset()
        hist, rew, term = [p], [], []
        assert np.array_equal(out["positions"][:, 0].cpu().numpy(), infos["positions"])
        for t in range(steps):
            a = out["actions"][:, t + 1].cpu().numpy()
            p, r, te, tr, infos = orc.step(a)
            hist.append(p); rew.append(torch.from_numpy(r)); term.append(torch.from_numpy(te))
            assert np.array_equal(out["positions"][:, t + 1].cpu().numpy(), infos["positions"])
        rewards = torch.stack(rew, dim=1)
        masks = torch.cat([torch.ones((b, 1), dtype=torch.bool), ~torch.stack(term, dim=1)], dim=1)
        want, lm = returns_oracle(rewards, masks)
        assert torch.equal(out["rewards"].cpu(), rewards) and torch.equal(out["masks"].cpu(), masks)
        assert torch.equal(out["returns"].cpu(), want) and torch.equal(out["logit_masks"].cpu(), lm)
        assert torch.equal(out["patches"].cpu(), torch.cat(hist, dim=1))
        assert steps == T or bool(torch.all(torch.stack(term, dim=1)[:, -1]))
