"""Rollout driver and returns tail for the REINFORCE path (reference:
``ReinforceTrainer.rollout``, ``src/reinforce.py:108-215``).

Only the env-side work lives here: stepping the batched env, keeping the per-step rewards /
termination flags in step-major device buffers, and the returns tail (K3,
``jn_returns``).  The policy (GPT + YOLOX in the reference) is any callable with the
reference model's signature; it stays in PyTorch and is outside this package.
"""
from typing import Callable, Dict, Optional, Tuple

import torch
from torch import Tensor
from torch.distributions import Categorical

from . import _cabi
from .env.general_env import NeedleGeneralEnv


def compute_returns(rewards: Tensor, logit_masks: Tensor) -> Tensor:
    """``returns[:, t] = sum_{s>=t} rewards[:, s] * logit_masks[:, s]`` for episode-major
    ``[B, T]`` inputs (reinforce.py:196-202), float64 accumulation from the last step."""
    _cabi.require_cuda(rewards, "rewards")
    assert rewards.dim() == 2 and rewards.shape == logit_masks.shape
    rewards = rewards.float()
    if rewards.stride(1) != 1:
        rewards = rewards.contiguous()
    lm = logit_masks.to(torch.bool)
    if lm.stride(1) != 1:
        lm = lm.contiguous()
    b, t = rewards.shape
    out = torch.empty((b, t), dtype=torch.float32, device=rewards.device)
    with _cabi.on_device(rewards.device):
        _cabi.check(_cabi.lib().jn_returns_rows(rewards.data_ptr(), rewards.stride(0), lm.data_ptr(), lm.stride(0),
                                                t, b, out.data_ptr(), _cabi.stream_ptr(rewards.device)))
    return out


def rollout_tail(rewards_tn: Tensor, terminated_tn: Tensor) -> Dict[str, Tensor]:
    """Step-major ``[T, B]`` rewards / terminated flags -> the reference's episode-major
    ``rewards [B,T]``, ``masks [B,T+1]``, ``logit_masks [B,T]``, ``returns [B,T]``
    (reinforce.py:186-202) in one kernel."""
    _cabi.require_cuda(rewards_tn, "rewards")
    t, b = rewards_tn.shape
    assert terminated_tn.shape == (t, b) and rewards_tn.dtype == torch.float32 and terminated_tn.dtype == torch.bool
    rewards_tn, terminated_tn = rewards_tn.contiguous(), terminated_tn.contiguous()
    dev = rewards_tn.device
    rewards = torch.empty((b, t), dtype=torch.float32, device=dev)
    masks = torch.empty((b, t + 1), dtype=torch.bool, device=dev)
    logit_masks = torch.empty((b, t), dtype=torch.bool, device=dev)
    returns = torch.empty((b, t), dtype=torch.float32, device=dev)
    with _cabi.on_device(dev):
        _cabi.check(_cabi.lib().jn_returns(rewards_tn.data_ptr(), terminated_tn.data_ptr(), t, b, rewards.data_ptr(),
                                           masks.data_ptr(), logit_masks.data_ptr(), returns.data_ptr(),
                                           _cabi.stream_ptr(dev)))
    return {"rewards": rewards, "masks": masks, "logit_masks": logit_masks, "returns": returns}


def sample_from_logits(logits: Tensor, take_best_action: bool = False) -> Tuple[Tensor, Tensor, Tensor]:
    """Last-token categorical sampling, as in reinforce.py:73-90 (stays in PyTorch)."""
    last = logits[:, -1, :]
    dist = Categorical(logits=last)
    actions = last.argmax(dim=1) if take_best_action else dist.sample()
    return actions, dist.log_prob(actions), dist.entropy()


@torch.no_grad()
def rollout(
    env: NeedleGeneralEnv,
    policy: Callable[[Tensor, Tensor, Tensor, Tensor, Optional[Tensor]], Tuple[Tensor, Optional[Tensor]]],
    sample_actions: bool = True,
    early_exit: bool = True,
) -> Dict[str, Tensor]:
    """One rollout on ``env`` (reinforce.py:108-215) without the detection branch.

    ``policy(patches [B,t,C,P,P], actions [B,t], classes [B], positions [B,t,2], embeddings)``
    returns ``(action_logits [B,t,n_actions], embeddings)``.  The history tensors handed to
    the policy are views of pre-allocated buffers (the env writes step t's crops into slot t
    when built with ``history=True``) instead of per-step ``torch.concat`` results; their
    values are identical.  ``early_exit`` keeps the reference's all-done check (one D2H sync
    per step); switch it off to run fixed-length rollouts without syncs.
    """
    b, t_max, dev = env.batch_size, env.max_ep_len, env.device
    actions = torch.zeros((b, t_max + 1), dtype=torch.long, device=dev)
    positions = torch.zeros((b, t_max + 1, 2), dtype=torch.long, device=dev)
    logprobs_tn = torch.empty((t_max, b), dtype=torch.float32, device=dev)
    entropies_tn = torch.empty((t_max, b), dtype=torch.float32, device=dev)
    classes = torch.zeros((b,), dtype=torch.int64, device=dev)
    keep_history = env._history is not None
    patches, infos = env.reset()
    positions[:, 0] = infos["positions"]
    embeddings = None
    steps = 0
    for step_id in range(t_max):
        hist = env.patch_history(step_id) if keep_history else patches
        logits, embeddings = policy(hist, actions[:, : step_id + 1], classes, positions[:, : step_id + 1], embeddings)
        new_actions, logprobs, entropies = sample_from_logits(logits, take_best_action=not sample_actions)
        # (the env keeps each step's rewards / flags in step-major rings: nothing to stack here)
        new_patches, step_rewards, terminated, truncated, infos = env.step(new_actions)
        logprobs_tn[step_id] = logprobs
        entropies_tn[step_id] = entropies
        actions[:, step_id + 1] = new_actions
        positions[:, step_id + 1] = infos["positions"]
        if not keep_history:
            patches = torch.concat((patches, new_patches), dim=1)
        steps = step_id + 1
        if early_exit and bool(torch.all(terminated | truncated)):
            break
    rewards_tn, terminated_tn, _ = env.rollout_buffers()
    tail = rollout_tail(rewards_tn, terminated_tn)
    tail.update(
        logprobs=logprobs_tn[:steps].t().contiguous(),
        entropies=entropies_tn[:steps].t().contiguous(),
        positions=positions[:, : steps + 1],
        actions=actions[:, : steps + 1],
        patches=env.patch_history(steps) if keep_history else patches,
        bboxes=[[] for _ in range(b)],
    )
    return tail
