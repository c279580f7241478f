"""Import the UNMODIFIED reference env modules (build container: straight from /root/reference).

The stand-ins for the three absent third-party packages and the import itself live in
``baseline/ref_env.py`` (shared with the benchmark's reference arm); this module keeps the name the
fixture generators use.  Fixtures are always generated from the read-only original, never from the
``baseline/_ref`` copy.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from baseline import ref_env  # noqa: E402

REFERENCE_ROOT = ref_env.REFERENCE_ROOT


def load_reference():
    """Returns (general_env_module, simple_env_module, common_module, utils_module)."""
    return ref_env.load(REFERENCE_ROOT if os.path.isdir(REFERENCE_ROOT) else "")
