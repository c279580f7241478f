"""Loader of the unmodified reference env (``src/env/*.py``, ``src/utils.py``) from ``baseline/_ref`` or, in the
build container, from ``/root/reference`` itself.

The reference is pure Python and has no ``setup.py`` / ``pyproject.toml``, so there is nothing for ``pip install
--target baseline/_ref`` to install; :func:`install` copies the handful of files of the gaze-environment path
instead (verbatim, with the licence).  The env modules import three third-party packages that are absent from this
image and irrelevant to the path: ``matplotlib`` (``src/utils.py:4``, plotting helpers), ``gymnasium`` (the ``Env``
base class and three space constructors, ``general_env.py:14,61-72``) and ``kornia.geometry.boxes.Boxes``
(``general_env.py:9,373-374``).  The stand-ins below replace them; the kornia one encodes the *documented*
``xyxy_plus`` / ``to_mask`` semantics (inclusive xmax / ymax, clamp to the image, fill 1) -- the one place where
parity is anchored on documentation rather than on code that can be run here ("parity unpinned at the
kornia boundary").
"""
import os
import shutil
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LOCAL_ROOT = os.path.join(HERE, "_ref")
REFERENCE_ROOT = "/root/reference"
# the gaze-environment path and its two callers (the callers cannot be imported -- they pull yolox, visdom,
# torchmetrics ... -- tests read single functions out of their source)
FILES = ["LICENSE", "src/utils.py", "src/env/__init__.py", "src/env/common.py", "src/env/general_env.py",
         "src/env/simple_env.py", "src/reinforce.py", "src/supervised.py"]


def install(reference_root: str = REFERENCE_ROOT) -> bool:
    """Copy the reference files of the path into ``baseline/_ref`` (verbatim).  Returns False when the
    reference is not there (GPU box: the prebuilt copy travels with the snapshot)."""
    if not os.path.isdir(reference_root):
        return False
    for rel in FILES:
        src, dst = os.path.join(reference_root, rel), os.path.join(LOCAL_ROOT, rel)
        if not os.path.exists(src):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
    with open(os.path.join(LOCAL_ROOT, "README"), "w") as f:
        f.write("Verbatim copies of jolibrain/jolineedle files (see LICENSE), made by baseline/ref_env.py:install();\n"
                "git-ignored, used only as the CPU reference arm of bench.py and as a test oracle.\n")
    return True


def root() -> str:
    """Directory the reference is imported from: the local copy, else the read-only original."""
    if os.path.exists(os.path.join(LOCAL_ROOT, "src", "env", "general_env.py")):
        return LOCAL_ROOT
    if os.path.exists(os.path.join(REFERENCE_ROOT, "src", "env", "general_env.py")):
        return REFERENCE_ROOT
    return ""


def available() -> bool:
    return bool(root())


class _StubBoxes:
    """``kornia.geometry.boxes.Boxes`` as far as general_env.py:373-374 uses it (documented semantics)."""

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


def install_stubs():
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib.pyplot  # noqa: F401
        except Exception:
            mpl = types.ModuleType("matplotlib")
            plt = types.ModuleType("matplotlib.pyplot")
            plt.Axes = plt.Figure = object  # only used in type annotations of plotting helpers
            mpl.pyplot = plt
            sys.modules["matplotlib"] = mpl
            sys.modules["matplotlib.pyplot"] = plt
    if "gymnasium" not in sys.modules:
        try:
            import gymnasium  # noqa: F401
        except Exception:
            gym = types.ModuleType("gymnasium")

            class Env:  # stand-in for gymnasium.Env
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
        try:
            import kornia.geometry.boxes  # noqa: F401
        except Exception:
            kornia = types.ModuleType("kornia")
            geometry = types.ModuleType("kornia.geometry")
            boxes = types.ModuleType("kornia.geometry.boxes")
            boxes.Boxes = _StubBoxes
            kornia.geometry = geometry
            geometry.boxes = boxes
            sys.modules["kornia"] = kornia
            sys.modules["kornia.geometry"] = geometry
            sys.modules["kornia.geometry.boxes"] = boxes


def kornia_is_real() -> bool:
    install_stubs()
    return sys.modules["kornia.geometry.boxes"].Boxes is not _StubBoxes


def load(reference_root: str = ""):
    """Returns (general_env_module, simple_env_module, common_module, utils_module) of the unmodified reference."""
    reference_root = reference_root or root()
    if not reference_root:
        raise FileNotFoundError("the reference is neither in baseline/_ref nor in /root/reference")
    install_stubs()
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    # `src/env/__init__.py` only pulls common.py, so importing the package is safe.
    import src.env.common as common
    import src.utils as utils
    import src.env.simple_env as simple_env
    import src.env.general_env as general_env

    return general_env, simple_env, common, utils


def function_source(rel_path: str, class_name: str, names) -> str:
    """Source text of the named methods of ``class_name`` in a reference file that cannot be imported as a whole
    (``src/reinforce.py`` / ``src/supervised.py`` pull yolox, visdom, torchmetrics ...): tests compile those
    methods, unmodified, into a stand-in class and run them against the drop-in env."""
    import ast
    import textwrap

    path = os.path.join(root(), rel_path)
    text = open(path).read()
    tree = ast.parse(text)
    lines = text.splitlines()
    out = []
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name == class_name:
            for item in node.body:
                if isinstance(item, ast.FunctionDef) and item.name in names:
                    first = min([item.lineno] + [d.lineno for d in item.decorator_list])
                    out.append(textwrap.dedent("\n".join(lines[first - 1:item.end_lineno])))
    return "\n\n".join(out)
