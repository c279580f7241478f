"""Import the UNMODIFIED reference env modules from /root/reference (build container only).

The reference is pure Python but needs three third-party modules that are absent from this
image: ``matplotlib`` (imported by src/utils.py:4 and never used on the env path),
``gymnasium`` (only the Env base class and three space constructors, general_env.py:14,61-72)
and ``kornia.geometry.boxes.Boxes`` (general_env.py:9,373-374).  They are replaced by the
minimal stand-ins below.  The kornia stand-in encodes the *documented* ``xyxy_plus`` /
``to_mask`` semantics (inclusive xmax/ymax, clamp to the image, fill 1) -- this is the one
place where parity is anchored on documentation rather than on code we can run
("parity unpinned at the kornia boundary").

Nothing under tests/ imports this module at test time: it is used by make_golden.py to
produce the committed fixtures, and /root/reference does not exist on the GPU box.
"""
import sys
import types

import torch

REFERENCE_ROOT = "/root/reference"


class _StubBoxes:
    def __init__(self, data):
        self._data = data

    @classmethod
    def from_tensor(cls, boxes, mode="xyxy", validate_boxes=True):
        assert mode == "xyxy_plus"
        w = boxes[..., 2] - boxes[..., 0] + 1
        h = boxes[..., 3] - boxes[..., 1] + 1
        if validate_boxes and ((w <= 0).any() or (h <= 0).any()):
            raise ValueError("Some boxes have negative widths/heights or 0.")
        return cls(boxes)

    def to_mask(self, height, width):
        b = self._data
        mask = torch.zeros((b.shape[0], b.shape[1], height, width), dtype=torch.float32)
        for i in range(b.shape[0]):
            for j in range(b.shape[1]):
                x1, y1, x2, y2 = (int(v) for v in b[i, j])
                x1c, x2c = min(max(x1, 0), width), min(max(x2 + 1, 0), width)
                y1c, y2c = min(max(y1, 0), height), min(max(y2 + 1, 0), height)
                mask[i, j, y1c:y2c, x1c:x2c] = 1
        return mask


def _install_stubs():
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        plt.Axes = plt.Figure = object  # only used in type annotations of plotting helpers
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
    if "gymnasium" not in sys.modules:
        gym = types.ModuleType("gymnasium")

        class Env:  # noqa: D401 - stand-in for gymnasium.Env
            pass

        class _Space:
            def __init__(self, *a, **k):
                pass

        spaces = types.ModuleType("gymnasium.spaces")
        spaces.Box = spaces.Tuple = spaces.Discrete = _Space
        gym.Env = Env
        gym.spaces = spaces
        sys.modules["gymnasium"] = gym
        sys.modules["gymnasium.spaces"] = spaces
    if "kornia" not in sys.modules:
        kornia = types.ModuleType("kornia")
        geometry = types.ModuleType("kornia.geometry")
        boxes = types.ModuleType("kornia.geometry.boxes")
        boxes.Boxes = _StubBoxes
        kornia.geometry = geometry
        geometry.boxes = boxes
        sys.modules["kornia"] = kornia
        sys.modules["kornia.geometry"] = geometry
        sys.modules["kornia.geometry.boxes"] = boxes


def load_reference():
    """Returns (general_env_module, simple_env_module, common_module, utils_module)."""
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # `src/env/__init__.py` only pulls common.py, so importing the package is safe.
    import src.env.common as common
    import src.utils as utils
    import src.env.simple_env as simple_env
    import src.env.general_env as general_env

    return general_env, simple_env, common, utils
