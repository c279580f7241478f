"""jolineedle_b200 -- B200-native gaze environment for JoliNeedle (one hot path, drop-in).

Public surface (mirrors the reference's names):
  env.general_env.NeedleGeneralEnv    batched RL env            (src/env/general_env.py)
  env.simple_env.NeedleSimpleEnv      per-image supervised env  (src/env/simple_env.py)
  env.simple_env.generate_trajectories batched supervised entry (src/supervised.py:95-136)
  reinforce.rollout / rollout_tail    rollout driver + returns  (src/reinforce.py:108-215)
  env.common.Action ...               action vocabulary         (src/env/common.py)
  utils.Position / BBox               value types               (src/utils.py:10-12)

Everything that touches pixels or per-episode state runs in the CUDA library
``libjolineedle_b200.so`` (C ABI: include/jolineedle_b200.h).  There is no CPU fallback.
"""
from . import _cabi  # noqa: F401
from .utils import BBox, Position, bboxes_to_tensor  # noqa: F401

__version__ = "0.1.0"
