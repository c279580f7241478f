"""Action vocabulary of the gaze environment.

Mirrors the public names of the reference's ``src/env/common.py:4-56`` (``Action``,
``ACTION_DELTAS``, ``MOVES``, ``ActionInfo``, ``get_actions_info``) so that callers written
against the reference import them unchanged.  The numeric codes are part of the contract
with the CUDA kernels (``csrc/jn_device.cuh: kActionDy/kActionDx``):

    code  name        (dy, dx)
    0     LEFT        ( 0, -1)
    1     RIGHT       ( 0, +1)
    2     UP          (-1,  0)
    3     DOWN        (+1,  0)
    4     LEFT_UP     (-1, -1)
    5     RIGHT_UP    (-1, +1)
    6     LEFT_DOWN   (+1, -1)
    7     RIGHT_DOWN  (+1, +1)
    8     STOP        ( 0,  0)
"""
from dataclasses import dataclass
from enum import Enum


class Action(Enum):
    LEFT = 0
    RIGHT = 1
    UP = 2
    DOWN = 3
    LEFT_UP = 4
    RIGHT_UP = 5
    LEFT_DOWN = 6
    RIGHT_DOWN = 7
    STOP = 8


# (dy, dx) per action code, in code order.  Kept as a flat tuple for the kernels and the
# host planner; ACTION_DELTAS is the dict view the reference exposes.
DELTA_TABLE = (
    (0, -1),
    (0, 1),
    (-1, 0),
    (1, 0),
    (-1, -1),
    (-1, 1),
    (1, -1),
    (1, 1),
    (0, 0),
)

ACTION_DELTAS = {action: DELTA_TABLE[action.value] for action in Action}

MOVES = [action for action in Action if action is not Action.STOP]

N_ACTIONS = len(DELTA_TABLE)
STOP_CODE = Action.STOP.value


@dataclass
class ActionInfo:
    action_type: str  # "categorical" or "scalar"
    nclasses: int


def get_actions_info(train_config):
    """One categorical head: 9 classes when STOP is enabled, else the 8 moves."""
    nclasses = N_ACTIONS if train_config.stop_enabled else N_ACTIONS - 1
    return [ActionInfo(action_type="categorical", nclasses=nclasses)]


def direction_code(dy: int, dx: int) -> int:
    """Action code of the greedy 8-direction step along the gradient (dy, dx).

    Same decision table as the reference's ``move_towards`` (simple_env.py:84-125):
    axis-aligned when one component is zero, diagonal while both are non-zero, STOP when
    both are zero.  Exposed as an integer function because the host planner and the
    trajectory kernel share it.
    """
    sy = (dy > 0) - (dy < 0)
    sx = (dx > 0) - (dx < 0)
    return _DIRECTION_LUT[(sy + 1) * 3 + (sx + 1)]


# index = (sign(dy)+1)*3 + (sign(dx)+1)
_DIRECTION_LUT = (
    4,  # dy<0 dx<0  LEFT_UP
    2,  # dy<0 dx=0  UP
    5,  # dy<0 dx>0  RIGHT_UP
    0,  # dy=0 dx<0  LEFT
    8,  # dy=0 dx=0  STOP
    1,  # dy=0 dx>0  RIGHT
    6,  # dy>0 dx<0  LEFT_DOWN
    3,  # dy>0 dx=0  DOWN
    7,  # dy>0 dx>0  RIGHT_DOWN
)
