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


# ---------------------------------------------------------------------------------------------------
# native (C++) planner == python planner (real numpy / random / set), no GPU needed
# ---------------------------------------------------------------------------------------------------
def _random_case(rng, n):
    P = int(rng.choice([16, 32, 64, 448]))
    heights, widths, bboxes = [], [], []
    for _ in range(n):
        gh, gw = int(rng.integers(1, 12)), int(rng.integers(1, 12))
        h, w = gh * P, gw * P
        boxes = []
        for _ in range(int(rng.integers(0, 5))):
            bw, bh = (int(v) for v in rng.integers(1, 3 * P, size=2))
            x1 = int(rng.integers(-P // 2, w))  # may start left of / above the image or end outside of it
            y1 = int(rng.integers(-P // 2, h))
            boxes.append(BBox(Position(y1, x1), Position(y1 + bh, x1 + bw)))
        heights.append(h); widths.append(w); bboxes.append(boxes)
    return P, heights, widths, bboxes


def _assert_same_plans(a, b):
    for f in ("start", "seg_begin", "seg_to", "seg_tgt", "seg_flags", "draw_begin", "draws", "det_begin", "det_yx",
              "rows", "cols", "n_boxes"):
        x, y = getattr(a, f), getattr(b, f)
        assert x.dtype == y.dtype and np.array_equal(x, y), f
    assert np.array_equal(a.boxes[:, :max(a.n_max, 1)], b.boxes[:, :max(b.n_max, 1)]) and a.n_max == b.n_max


def test_native_planner_reproduces_the_python_planner():
    import random

    from jolineedle_b200.env.trajectories import plan_batch

    rng = np.random.default_rng(2024)
    n_episodes = 0
    for it in range(120):
        n = int(rng.integers(1, 24))
        P, heights, widths, bboxes = _random_case(rng, n)
        binomial = bool(rng.integers(0, 2))
        kmin = int(rng.integers(0, 3)); kmax = kmin + int(rng.integers(0, 4))
        seeds = [int(s) for s in rng.integers(0, 2**63, size=n)]
        if it % 5 == 0:
            seeds[0] = 0
            seeds[-1] = 2**64 - 1
        position = None
        if it % 3 == 0:
            position = Position(int(rng.integers(0, min(heights) // P)), int(rng.integers(0, min(widths) // P)))
        random.seed(it)
        py = plan_batch(bboxes, heights, widths, P, kmin, kmax, binomial, position, seeds, planner="python")
        state_py = random.getstate()
        random.seed(it)
        nat = plan_batch(bboxes, heights, widths, P, kmin, kmax, binomial, position, seeds, planner="native")
        assert random.getstate() == state_py, "python's global random stream must advance identically"
        _assert_same_plans(py, nat)
        n_episodes += n
    assert n_episodes > 1000


def test_native_planner_reproduces_reference_fixtures():
    """Golden trajectories of the unmodified reference, through native plan + sequential expansion."""
    import random
    from types import SimpleNamespace

    from jolineedle_b200.env.trajectories import plan_batch

    fx = load_golden("simple_env.npz")
    for name in fx["names"]:
        c, cfg = simple_case(fx, str(name))
        seed_python_random(cfg["seed"])
        boxes = [BBox(Position(y1, x1), Position(y2, x2)) for (x1, y1, x2, y2) in c["raw_boxes"].tolist()]
        pos = None if cfg["position"] is None else Position(*cfg["position"])
        _, h, w = c["u8"].shape
        p = plan_batch([boxes], [h], [w], cfg["P"], cfg["kmin"], cfg["kmax"], cfg["binomial"], pos, [cfg["seed"]],
                       planner="native")
        plan = SimpleNamespace(start=tuple(p.start[0].tolist()), seg_to=[tuple(v) for v in p.seg_to.tolist()],
                               seg_tgt=[tuple(v) for v in p.seg_tgt.tolist()], seg_first=p.seg_flags.tolist(),
                               draws=p.draws.tolist())
        inside = {tuple(r) for r in c["bbox_patches"].tolist()}
        got = expand_plan_host(plan, cfg["T"], lambda y, x: (y, x) in inside)
        for k in ("positions", "current_actions", "next_actions", "labels", "masks"):
            assert np.array_equal(got[k], c[k]), (name, k)
        P = cfg["P"]
        img = to_f32(c["u8"])
        tiles = torch.stack([img[:, y * P:(y + 1) * P, x * P:(x + 1) * P] for (y, x) in p.det_yx.tolist()])
        assert np.array_equal(tiles.numpy(), c["patches_yolox"]), name


def test_native_planner_unseeded_and_unsupported_inputs():
    from jolineedle_b200.env.trajectories import plan_batch

    boxes = [[BBox(Position(10, 10), Position(60, 90))]]
    a = plan_batch(boxes, [128], [160], 32, 0, 3, True, None, None, planner="native")  # OS entropy: just runs
    assert a.n == 1 and a.seg_begin[-1] >= 1
    # float boxes -> python planner; forcing native is an error
    fboxes = [[BBox(Position(10.5, 10.25), Position(60.0, 90.0))]]
    b = plan_batch(fboxes, [128], [160], 32, 0, 0, False, Position(0, 0), [3], planner="auto")
    assert b.n == 1
    try:
        plan_batch(fboxes, [128], [160], 32, 0, 0, False, None, [3], planner="native")
        raise RuntimeError("expected ValueError")
    except ValueError:
        pass
    # start position outside the grid -> AssertionError like get_patch (simple_env.py:73-74)
    try:
        plan_batch(boxes, [128], [160], 32, 0, 0, False, Position(9, 0), [1], planner="native")
        raise RuntimeError("expected AssertionError")
    except AssertionError:
        pass


def test_best_next_action_equals_reference_eval_oracle():
    """supervised.py:301-309: the expert action = next_actions[0] of a fresh 50-step sample generated from the
    current position / visited set.  The host-only shortcut must give the same action and leave both random
    streams (python's global one and the env's numpy generator) where the reference call leaves them."""
    import copy
    import random

    from oracle.traj_oracle import Pos, TrajectoryOracle

    fx = load_golden("simple_env.npz")
    rng = np.random.default_rng(0)
    checked = 0
    for name in fx["names"]:
        c, cfg = simple_case(fx, str(name))
        boxes = c["raw_boxes"].tolist()
        env = make_env(c, cfg)
        orc = TrajectoryOracle(to_f32(c["u8"]), cfg["P"], [((y1, x1), (y2, x2)) for (x1, y1, x2, y2) in boxes],
                               seed=cfg["seed"])
        gh, gw = env.patch_height, env.patch_width
        for _ in range(4):
            pos = (int(rng.integers(0, gh)), int(rng.integers(0, gw)))
            visited = {p for p in sorted(env.bbox_patches) if rng.random() < 0.3}
            random.seed(checked)
            want = int(orc.generate_sample(50, 0, 0, False, Pos(*pos), {Pos(*v) for v in visited})["next_actions"][0])
            state_ref, np_ref = random.getstate(), orc.rng.bit_generator.state
            random.seed(checked)
            got = env.best_next_action(Position(*pos), {Position(*v) for v in visited})
            assert got.value == want, (name, pos)
            assert random.getstate() == state_ref
            assert env.rng.bit_generator.state == np_ref
            checked += 1
    assert checked >= 100


def test_native_planner_on_its_own_thread_matches_the_blocking_call():
    """jn_plan_start / jn_plan_wait: same plans, same advanced `random` state; one run at a time; errors surface
    at the join."""
    import random

    import pytest

    from jolineedle_b200 import _cabi
    from jolineedle_b200.env import trajectories as tj

    P, heights, widths, bboxes = _random_case(np.random.default_rng(31), 40)
    seeds = list(range(500, 540))
    random.seed(9)
    want = tj.plan_batch(bboxes, heights, widths, P, 0, 3, True, None, seeds, planner="native")
    state_after = random.getstate()
    random.seed(9)
    finish = tj.plan_batch(bboxes, heights, widths, P, 0, 3, True, None, seeds, planner="native", deferred=True)
    # a second start while the first is in flight is refused, not queued
    rows = np.array([h // P for h in heights], dtype=np.int32)
    lib, handle = _cabi.lib(), tj._native_plan.handle
    mt = np.zeros(625, dtype=np.uint32)
    rc = lib.jn_plan_start(handle, 0, None, None, 0, rows.ctypes.data, rows.ctypes.data, P, None, None, 0, 0, 0, None,
                           mt.ctypes.data)
    assert rc == _cabi.JN_ERR_INVALID
    got = finish()
    assert random.getstate() == state_after
    _assert_same_plans(want, got)
    # a start position outside the grid is reported by the join, as the reference's assert (simple_env.py:73-74)
    finish = tj.plan_batch(bboxes, heights, widths, P, 0, 3, True, Position(99, 0), seeds, planner="native", deferred=True)
    with pytest.raises(AssertionError):
        finish()
    again = tj.plan_batch(bboxes, heights, widths, P, 0, 3, True, None, seeds, planner="native")  # usable again
    assert again.n == want.n


def test_native_planner_wide_grids_follow_numpy_btpe():
    """Grids wider than 60 patches: numpy's binomial switches from inversion to BTPE at n*p > 30; the native
    planner restates both.  Around the switch (55..70) and far above it, against the planner that calls numpy."""
    import random

    from jolineedle_b200.env.trajectories import plan_batch

    rng = np.random.default_rng(77)
    segments = 0
    for it in range(24):
        n, P = int(rng.integers(1, 5)), 16
        heights, widths, bboxes = [], [], []
        for _ in range(n):
            lo, hi = ((55, 70) if it % 2 else (61, 300))
            gh, gw = int(rng.integers(lo, hi)), int(rng.integers(lo, hi))
            boxes = []
            for _ in range(int(rng.integers(0, 4))):
                bw, bh = (int(v) for v in rng.integers(1, 3 * P, size=2))
                x1, y1 = int(rng.integers(0, gw * P)), int(rng.integers(0, gh * P))
                boxes.append(BBox(Position(y1, x1), Position(y1 + bh, x1 + bw)))
            heights.append(gh * P); widths.append(gw * P); bboxes.append(boxes)
        seeds = [int(s) for s in rng.integers(0, 2**63, size=n)]
        random.seed(it)
        py = plan_batch(bboxes, heights, widths, P, 2, 6, True, None, seeds, planner="python")
        state_py = random.getstate()
        random.seed(it)
        nat = plan_batch(bboxes, heights, widths, P, 2, 6, True, None, seeds, planner="native")
        assert random.getstate() == state_py
        _assert_same_plans(py, nat)
        segments += len(py.seg_flags)
    assert segments > 500


def test_native_planner_self_check_and_fallback(monkeypatch):
    """First use compares the native planner with the python one on seeded episodes; a mismatch (a CPython /
    numpy bump) switches to the python planner with a warning instead of silently changing seeded trajectories."""
    import warnings

    from jolineedle_b200.env import trajectories as tr

    monkeypatch.setattr(tr, "_native_verdict", None)
    import random

    random.seed(1234)
    before = random.getstate()
    assert tr.native_planner_verified() is True and tr._native_verdict is True
    assert random.getstate() == before  # the check leaves python's global stream where it was

    # a planner that disagrees: perturb one exported array
    real = tr.plan_native

    def broken(*a, **k):
        p = real(*a, **k)
        if len(p.seg_to):
            p.seg_to = p.seg_to.copy()
            p.seg_to[0, 0] += 1
        return p

    monkeypatch.setattr(tr, "_native_verdict", None)
    monkeypatch.setattr(tr, "plan_native", broken)
    with warnings.catch_warnings(record=True) as seen:
        warnings.simplefilter("always")
        assert tr.native_planner_verified() is False
    assert any("falling back to the python planner" in str(w.message) for w in seen)
    boxes = [[BBox(Position(10, 12), Position(60, 70))]]
    random.seed(8)
    a = tr.plan_batch(boxes, [128], [160], 32, 0, 3, True, None, [5], planner="auto")   # python planner now
    monkeypatch.setattr(tr, "plan_native", real)
    monkeypatch.setattr(tr, "_native_verdict", None)
    random.seed(8)
    b = tr.plan_batch(boxes, [128], [160], 32, 0, 3, True, None, [5], planner="native")
    for k in ("start", "seg_begin", "seg_to", "seg_tgt", "seg_flags", "draw_begin", "draws", "det_begin", "det_yx"):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k


def test_boxes_array_flattening_and_exactness_flag():
    """Host packing of the trainer's lists of BBox: x1, y1, x2, y2 order, zero-padded rows, per-image counts, and
    the flag that sends boxes with fractional coordinates to the float64 kernels / the python planner."""
    from jolineedle_b200.env.trajectories import boxes_array

    boxes = [[BBox(Position(2, 1), Position(8, 5)), BBox(Position(0, 0), Position(3, 3))], [],
             [BBox(Position(7, 9), Position(20, 30))]]
    arr, counts, n_max, exact, arr_f = boxes_array(boxes, want_float=True)
    assert arr.dtype == np.int64 and arr.shape == (3, 2, 4) and counts.tolist() == [2, 0, 1] and n_max == 2
    assert arr[0].tolist() == [[1, 2, 5, 8], [0, 0, 3, 3]] and arr[1].tolist() == [[0] * 4] * 2
    assert arr[2].tolist() == [[9, 7, 30, 20], [0, 0, 0, 0]] and exact and arr_f is None
    # whole-pixel floats are still exact; one fractional coordinate is not
    whole = [[BBox(Position(2.0, 1.0), Position(8.0, 5.0))]]
    assert boxes_array(whole, want_float=True)[3:] == (True, None)
    frac = [[BBox(Position(2.5, 1), Position(8, 5))], [BBox(Position(1, 1), Position(4, 4))]]
    arr, counts, n_max, exact, arr_f = boxes_array(frac, want_float=True)
    assert not exact and arr_f.dtype == np.float64 and arr_f[0, 0].tolist() == [1.0, 2.5, 5.0, 8.0]
    assert arr_f[1, 0].tolist() == [1.0, 1.0, 4.0, 4.0] and arr[0, 0].tolist() == [1, 2, 5, 8]  # (truncated copy)
    # no boxes at all
    arr, counts, n_max, exact = boxes_array([[], []])
    assert arr.shape == (2, 1, 4) and counts.tolist() == [0, 0] and n_max == 0 and exact
    # fractional boxes are planned by the python planner, whatever was asked for with "auto"
    from jolineedle_b200.env.trajectories import plan_batch

    p = plan_batch(frac, [64, 64], [64, 64], 16, 0, 2, True, None, [1, 2], planner="auto")
    assert p.boxes_f64 is not None and p.boxes_f64.shape == (2, 1, 4)
