"""Which sources is the loaded library built from?

The Makefile hashes the native sources (``csrc/*`` + ``include/*.h``, in the order below) into the library
(``jn_source_hash()``); :func:`source_hash` computes the same digest from the working tree.  Profiles under
``profiles/`` are stamped with it, so that a number copied from an ncu capture is only quoted for the code that
was actually profiled (``bench.py`` prints ``traffic: null`` otherwise).
"""
import hashlib
import os

from . import _cabi

_HERE = os.path.dirname(os.path.abspath(__file__))
# same list, same order as SRC + HDR of csrc/Makefile
SOURCES = ["csrc/jn_api.cu", "csrc/jn_planner.cpp", "csrc/jn_device.cuh", "csrc/jn_gather.cuh", "csrc/jn_env.cuh",
           "csrc/jn_scan.cuh", "csrc/jn_pyramid.cuh", "../include/jolineedle_b200.h"]


def source_hash() -> str:
    """First 16 hex digits of the SHA-256 of the concatenated native sources in the working tree."""
    h = hashlib.sha256()
    for rel in SOURCES:
        with open(os.path.join(_HERE, rel), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def library_source_hash() -> str:
    """The digest compiled into the loaded ``libjolineedle_b200.so``."""
    return _cabi.lib().jn_source_hash().decode("ascii")
