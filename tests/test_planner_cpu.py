"""Host half of the supervised pipeline (no GPU): the product's planner makes the reference's
RNG calls in the reference's order, so expanding its plan sequentially must reproduce the
reference trajectories of the golden fixtures."""
import numpy as np
import torch

from helpers import expand_plan_host, load_golden, simple_case, seed_python_random, to_f32
from jolineedle_b200.env.common import Action, MOVES, direction_code, get_actions_info, ACTION_DELTAS
from jolineedle_b200.env.simple_env import NeedleSimpleEnv, move_towards, pixel_pos_to_patch_pos
from jolineedle_b200.utils import BBox, Position, bboxes_to_tensor


def make_env(c, cfg):
    boxes = [BBox(Position(y1, x1), Position(y2, x2)) for (x1, y1, x2, y2) in c["raw_boxes"].tolist()]
    return NeedleSimpleEnv(to_f32(c["u8"]), cfg["P"], boxes, seed=cfg["seed"])


def test_planner_reproduces_reference_trajectories():
    fx = load_golden("simple_env.npz")
    for name in fx["names"]:
        c, cfg = simple_case(fx, str(name))
        seed_python_random(cfg["seed"])
        env = make_env(c, cfg)
        assert sorted(env.bbox_patches) == [tuple(r) for r in c["bbox_patches"].tolist()]
        pos = None if cfg["position"] is None else Position(*cfg["position"])
        plan = env.plan_sample(cfg["kmin"], cfg["kmax"], cfg["binomial"], pos)
        got = expand_plan_host(plan, cfg["T"], lambda y, x: Position(y, x) in env.bbox_patches)
        for k in ("positions", "current_actions", "next_actions", "labels", "masks"):
            assert np.array_equal(got[k], c[k]), (name, k)
        # detection patches: same patches in the same (set-iteration) order
        P = cfg["P"]
        img = to_f32(c["u8"])
        tiles = torch.stack([img[:, y * P:(y + 1) * P, x * P:(x + 1) * P] for (y, x) in plan.det_positions])
        assert np.array_equal(tiles.numpy(), c["patches_yolox"]), name
        boxes = torch.stack([env.local_bboxes(p) for p in plan.det_positions])
        assert np.array_equal(boxes.numpy(), c["bboxes_yolox"]), name
        # per-step local boxes
        ep = int(c["masks"].sum())
        for t in range(ep):
            lb = env.local_bboxes(Position(*c["positions"][t].tolist()))
            assert np.array_equal(lb.numpy(), c["local_bboxes"][t]), (name, t)


def test_action_vocabulary():
    assert [a.name for a in Action] == ["LEFT", "RIGHT", "UP", "DOWN", "LEFT_UP", "RIGHT_UP", "LEFT_DOWN",
                                        "RIGHT_DOWN", "STOP"]
    assert [a.value for a in Action] == list(range(9))
    assert ACTION_DELTAS[Action.LEFT] == (0, -1) and ACTION_DELTAS[Action.DOWN] == (1, 0)
    assert ACTION_DELTAS[Action.RIGHT_UP] == (-1, 1) and ACTION_DELTAS[Action.STOP] == (0, 0)
    assert MOVES == [a for a in Action if a != Action.STOP] and len(MOVES) == 8

    class Cfg:
        stop_enabled = True

    assert get_actions_info(Cfg)[0].nclasses == 9
    Cfg.stop_enabled = False
    assert get_actions_info(Cfg)[0].nclasses == 8 and get_actions_info(Cfg)[0].action_type == "categorical"
    # moving along the returned direction reduces the Chebyshev distance by one
    for dy in range(-3, 4):
        for dx in range(-3, 4):
            a = move_towards(Position(0, 0), Position(dy, dx))
            if dy == 0 and dx == 0:
                assert a == Action.STOP
                continue
            ddy, ddx = ACTION_DELTAS[a]
            assert max(abs(dy - ddy), abs(dx - ddx)) == max(abs(dy), abs(dx)) - 1
            assert direction_code(dy, dx) == a.value


def test_value_types():
    b = [BBox(Position(2, 1), Position(8, 5)), BBox(Position(0, 0), Position(3, 3))]
    t = bboxes_to_tensor(b)
    assert t.tolist() == [[1, 2, 5, 8], [0, 0, 3, 3]]  # x1, y1, x2, y2
    assert pixel_pos_to_patch_pos(Position(447, 448), 448) == Position(0, 1)
