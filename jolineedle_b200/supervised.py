"""Trainer-side boundary of the supervised pipeline (reference: ``src/supervised.py:85-136``).

``TrajectoryMixin`` carries ``create_env`` / ``generate_trajectories`` with the reference
trainer's signatures: a ``SupervisedTrainer`` that inherits it gets its
batches from the CUDA path without any other change.  It only reads ``self.config`` fields
the reference trainer already has (patch_size, max_seq_len, min/max_keypoints,
binomial_keypoints) and, optionally, ``self.device``.
"""
from typing import Dict, Optional, Tuple

from .env.simple_env import NeedleSimpleEnv, generate_trajectories
from .utils import Position


class TrajectoryMixin:
    def create_env(self, sample: Dict) -> Tuple[NeedleSimpleEnv, int]:  # supervised.py:85-93
        env = NeedleSimpleEnv(sample["image"], self.config.patch_size, sample["bboxes"],
                              device=getattr(self, "device", None))
        return env, sample["class_id"]

    def generate_trajectories(self, batch: Dict, position: Optional[Position] = None) -> Dict:  # supervised.py:95-136
        cfg = self.config
        return generate_trajectories(
            batch, cfg.patch_size, cfg.max_seq_len, cfg.min_keypoints, cfg.max_keypoints,
            binomial_keypoints=cfg.binomial_keypoints, position=position,
            normalize=getattr(cfg, "normalize_on_gather", False), device=getattr(self, "device", None),
            focus=getattr(cfg, "focus_layout", False),
        )
