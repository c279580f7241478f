"""The C-ABI library loads without a GPU and exports every symbol the header declares."""
import ctypes
import os
import re

import numpy as np
import pytest

from helpers import load_golden
from jolineedle_b200 import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "jolineedle_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set(re.findall(r"\b(jn_[a-z0-9_]+)\s*\(", text))
    names -= {"jn_bitmap_words", "jn_images_table_bytes"}  # static inline helpers
    return sorted(names)


def test_library_exports_every_declared_symbol():
    lib = _cabi.lib()
    names = declared_symbols()
    assert len(names) >= 18
    for name in names:
        assert hasattr(lib, name), f"{name} is declared in the header but missing from the library"
        assert name in _cabi.SIGNATURES, f"{name} has no ctypes signature in _cabi.py"
    assert sorted(_cabi.SIGNATURES) == names
    assert lib.jn_abi_version() == 1


def test_host_selftest_matches_torch_and_reference_table():
    lib = _cabi.lib()
    unit = (ctypes.c_float * 256)()
    direction = (ctypes.c_int * 9)()
    assert lib.jn_selftest_host(unit, direction) == 0
    assert np.array_equal(np.array(unit[:], dtype=np.float32), load_golden("norm.npz")["u8_over_255"])
    # (sign(dy)+1)*3 + sign(dx)+1 -> action code, move_towards decision table
    assert list(direction) == [4, 2, 5, 0, 8, 1, 6, 3, 7]


def test_gather_ticket_schedule_tiles_the_launch():
    """The converting gather hands its chunks out as tickets; whatever the launch size and grid, the tickets must
    cover every chunk exactly once, in order, with batch sizes that never grow towards the end."""
    lib = _cabi.lib()
    rng = np.random.default_rng(0)
    cases = [(1, 1), (3, 3), (6144, 296), (43008, 148), (49152, 296), (31, 148), (296, 296), (297, 296), (1 << 20, 444)]
    cases += [(int(rng.integers(1, 200000)), int(rng.integers(1, 600))) for _ in range(200)]
    for i, (total, grid) in enumerate(cases):
        grid = min(grid, total)
        batch = (1, 2, 4, 8, 16, 32)[i % 6]
        sizes, tickets, chunks = (ctypes.c_int32 * 6)(), (ctypes.c_int32 * 7)(), (ctypes.c_int32 * 7)()
        n = lib.jn_claim_schedule_host(total, grid, batch, sizes, tickets, chunks)
        assert 1 <= n <= 6 and tickets[0] == 0 and chunks[0] == 0 and chunks[n] == total, (total, grid)
        covered, last_size = 0, 33
        for j in range(n):
            n_batches = tickets[j + 1] - tickets[j]
            assert n_batches > 0 and 1 <= sizes[j] <= batch and sizes[j] < last_size
            last_size = sizes[j]
            span = chunks[j + 1] - chunks[j]
            assert (n_batches - 1) * sizes[j] < span <= n_batches * sizes[j]  # only the last batch may be partial
            assert chunks[j] == covered
            covered += span
        assert covered == total
        # the tail is fine-grained: a CTA never holds more than a few chunks when the tickets run out
        if total >= 70 * grid and batch > 1:
            assert sizes[0] == batch and sizes[n - 1] == 1 and tickets[n] - tickets[n - 1] == 6 * grid


def test_invalid_arguments_are_reported_not_crashed():
    lib = _cabi.lib()
    handle = ctypes.c_void_p()
    ptrs = (ctypes.c_void_p * 1)(16)
    one = (ctypes.c_int32 * 1)(1)
    h = (ctypes.c_int32 * 1)(100)
    w = (ctypes.c_int32 * 1)(96)
    rc = lib.jn_images_create(ctypes.byref(handle), 1, ptrs, one, h, w, 3, _cabi.JN_F32, 16, None, None, None)
    assert rc == _cabi.JN_ERR_INVALID  # 100 is not a multiple of 16
    assert b"multiple of patch_size" in lib.jn_last_error()
    with pytest.raises(AssertionError):
        _cabi.check(rc, invalid_exc=AssertionError)


def test_product_refuses_cpu_tensors():
    import torch
    from jolineedle_b200.env.general_env import NeedleGeneralEnv

    with pytest.raises(_cabi.NativeLibraryError):
        NeedleGeneralEnv(torch.zeros(1, 3, 32, 32), torch.zeros(1, 1, 4, dtype=torch.long), 16, 4, 1)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "jolineedle_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
