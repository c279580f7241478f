"""ctypes binding of ``libjolineedle_b200.so`` (the C ABI in ``include/jolineedle_b200.h``).

The library is the product: there is no Python or CPU fallback behind these calls.  If the
shared object is missing or a call fails, an exception is raised.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_int64, c_longlong, c_uint32, c_void_p
from typing import Optional

import torch

_LIB_NAME = "libjolineedle_b200.so"
# JN_LIB_PATH: development override (A/B of two builds under the same harness)
_LIB_PATH = os.environ.get("JN_LIB_PATH") or os.path.join(os.path.dirname(os.path.abspath(__file__)), _LIB_NAME)

JN_OK, JN_ERR_INVALID, JN_ERR_CUDA, JN_ERR_UNSUPPORTED, JN_ERR_NO_DEVICE = range(5)
JN_U8, JN_F32 = 0, 1
GATHER_NORMALIZE, GATHER_FOCUS, GATHER_SHIFT_ALIGNED, GATHER_SKIP_NEGATIVE = 1, 2, 4, 8
ENGINE_AUTO, ENGINE_TENSOR, ENGINE_BULK, ENGINE_LDG = 0, 1, 2, 3
ENGINES = {"auto": ENGINE_AUTO, "tensor": ENGINE_TENSOR, "bulk": ENGINE_BULK, "ldg": ENGINE_LDG}
RULE_ANY_PIXEL, RULE_AREA5 = 0, 1

# status bits written by the kernels
STATUS_BAD_POSITION, STATUS_BAD_ACTION, STATUS_BAD_BOX, STATUS_BAD_PLAN = 1, 2, 4, 8


class NativeLibraryError(RuntimeError):
    """The CUDA library is missing or a call into it failed."""


# name -> (restype, argtypes); one entry per symbol declared in include/jolineedle_b200.h
_P = c_void_p


class EnvStepArgs(ctypes.Structure):
    """``jn_env_step_args`` of the header: borrowed device pointers of one env + the gather of its new glimpses.
    An env fills it once and rewrites the handful of per-step fields before every call."""

    _fields_ = [
        ("pos_in", _P), ("actions", _P), ("pos_out", _P), ("visited", _P), ("bbox", _P), ("steps", _P),
        ("has_stopped", _P), ("rewards", _P), ("terminated", _P), ("truncated", _P), ("first_slot", _P),
        ("host_src", _P), ("history_src", _P), ("host_tiles", _P), ("status", _P),
        ("n", c_int32), ("rows", c_int32), ("cols", c_int32), ("max_ep_len", c_int32), ("stop_enabled", c_int32),
        ("slots", c_int32), ("t", c_int32), ("cost", c_float),
        ("shifts", _P), ("out", _P), ("out_item_stride_bytes", c_int64), ("flags", c_uint32), ("engine", c_int32),
    ]


SIGNATURES = {
    "jn_abi_version": (c_int, []),
    "jn_last_error": (c_char_p, []),
    "jn_launch_count": (c_longlong, []),
    "jn_source_hash": (c_char_p, []),
    "jn_claim_schedule_host": (c_int, [c_int, c_int, c_int, POINTER(c_int32), POINTER(c_int32), POINTER(c_int32)]),
    "jn_env_step_gather": (c_int, [_P, _P, POINTER(EnvStepArgs), _P]),
    "jn_env_reset_gather": (c_int, [_P, _P, POINTER(EnvStepArgs), _P]),
    "jn_device_info": (c_int, [POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "jn_selftest_host": (c_int, [POINTER(c_float), POINTER(c_int)]),
    "jn_images_create": (c_int, [POINTER(_P), c_int, _P, _P, _P, _P, c_int, c_int, c_int, _P, _P, _P]),
    "jn_images_create_padded": (c_int, [POINTER(_P), c_int, _P, _P, _P, _P, c_int, c_int, c_int, _P, _P, _P]),
    "jn_images_destroy": (None, [_P]),
    "jn_images_tma_ok": (c_int, [_P, c_int]),
    "jn_gather": (c_int, [_P, _P, _P, _P, c_int, _P, c_int64, c_uint32, c_int, _P, _P]),
    "jn_patch_bitmaps": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_int, _P, _P, c_int, _P, c_int, _P]),
    "jn_patch_bitmaps_f64": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_int, _P, _P, c_int, _P, c_int, _P]),
    "jn_local_boxes_f64": (c_int, [_P, _P, c_int, c_int, _P, _P, c_int, _P, _P]),
    "jn_resize_aa_reflect": (c_int, [_P, c_int64, _P, _P, c_int64, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P, _P,
                                     c_int, _P, _P, _P, c_int, _P]),
    "jn_bitmap_unpack": (c_int, [_P, c_int, c_int, c_int, _P, _P]),
    "jn_split_boxes": (c_int, [_P, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P]),
    "jn_local_boxes": (c_int, [_P, _P, c_int, c_int, _P, _P, c_int, _P, _P]),
    "jn_env_reset": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, _P, _P]),
    "jn_env_step": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_float, c_int,
                            _P, _P]),
    "jn_env_props": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, _P, _P, _P]),
    "jn_env_rewards": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_float, c_int, _P, _P]),
    "jn_returns": (c_int, [_P, _P, c_int, c_int, _P, _P, _P, _P, _P]),
    "jn_returns_rows": (c_int, [_P, c_int64, _P, c_int64, c_int, c_int, _P, _P]),
    "jn_traj_expand": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, c_int, _P, c_int, c_int, _P, _P, _P, _P, _P, _P,
                               _P, _P, _P]),
    "jn_tile_lookup": (c_int, [_P, _P, c_int, _P, _P, c_int, c_int, _P, _P, _P]),
    "jn_tile_dedupe": (c_int, [_P, _P, c_int, c_int, _P, _P, _P]),
    "jn_visit_sources": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P]),
    "jn_plan_create": (c_int, [POINTER(_P)]),
    "jn_plan_destroy": (None, [_P]),
    "jn_plan_error": (c_char_p, [_P]),
    "jn_plan_run": (c_int, [_P, c_int, _P, _P, c_int, _P, _P, c_int, _P, _P, c_int, c_int, c_int, _P, _P]),
    "jn_plan_start": (c_int, [_P, c_int, _P, _P, c_int, _P, _P, c_int, _P, _P, c_int, c_int, c_int, _P, _P]),
    "jn_plan_wait": (c_int, [_P]),
    "jn_plan_sizes": (c_int, [_P, POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "jn_plan_export": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
}

_lib: Optional[ctypes.CDLL] = None


def library_path() -> str:
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    """Load the shared object (once).  Raises ``NativeLibraryError`` when it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise NativeLibraryError(
                f"{_LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C jolineedle_b200/csrc` (there is no CPU fallback)"
            )
        handle = ctypes.CDLL(_LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError here = header and library out of sync
            fn.restype = restype
            fn.argtypes = argtypes
        if handle.jn_abi_version() != 1:
            raise NativeLibraryError(f"ABI version mismatch: library reports {handle.jn_abi_version()}")
        _lib = handle
    return _lib


def launch_count() -> int:
    """Kernels launched by the library since it was loaded (counted at the launch sites, in C)."""
    return int(lib().jn_launch_count())


def check(rc: int, invalid_exc=ValueError):
    if rc == JN_OK:
        return
    msg = lib().jn_last_error().decode("utf-8", "replace")
    if rc == JN_ERR_INVALID:
        raise invalid_exc(msg)
    raise NativeLibraryError(f"jolineedle_b200 native call failed (status {rc}): {msg}")


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    """Raw device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr(device) -> int:
    """cudaStream_t of torch's current stream on ``device`` -- every launch goes there."""
    index = device.index if isinstance(device, torch.device) else torch.device(device).index
    if index is None:
        index = torch.cuda.current_device()
    return _raw_stream(index)


# the raw-handle accessor skips the construction of a torch.cuda.Stream object (~8 us per call, a few
# dozen calls per batch); fall back to the public API if a torch release moves it
_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None) or (
    lambda index: torch.cuda.current_stream(index).cuda_stream)


class on_device:
    """``with on_device(dev):`` -- make ``dev`` the current CUDA device for the library calls inside.
    Unlike ``torch.cuda.device`` it does nothing at all (no context object, no driver call) when ``dev``
    already is the current device, which is the case in every one-process-per-GPU program."""

    __slots__ = ("index", "guard")

    def __init__(self, device):
        self.index = device.index
        self.guard = None

    def __enter__(self):
        if self.index is not None and torch.cuda.current_device() != self.index:
            self.guard = torch.cuda.device(self.index)
            self.guard.__enter__()

    def __exit__(self, *exc):
        if self.guard is not None:
            self.guard.__exit__(*exc)
            self.guard = None
        return False


def require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise NativeLibraryError(
            f"{what} must live on a CUDA device (got {t.device}); jolineedle_b200 has no CPU path"
        )


def dtype_code(dtype: torch.dtype) -> int:
    if dtype == torch.uint8:
        return JN_U8
    if dtype == torch.float32:
        return JN_F32
    raise ValueError(f"images must be uint8 or float32, got {dtype}")


def device_info():
    sm, major, minor = c_int(), c_int(), c_int()
    check(lib().jn_device_info(ctypes.byref(sm), ctypes.byref(major), ctypes.byref(minor)))
    return sm.value, major.value, minor.value
