"""CPU oracle for the gaze-environment hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

This package restates, on the CPU with numpy / torch-CPU, the algorithm of the reference's
gaze environment (jolibrain/jolineedle ``src/env/*.py`` and the returns tail of
``src/reinforce.py``).  It exists to *check* the CUDA path.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs
may import it; nothing under ``jolineedle_b200/`` does, and the product raises if its CUDA
library is missing instead of falling back to this code.

Pinning: the restatement is validated against the UNMODIFIED reference modules imported in
the build container (``tests/golden/make_golden.py``), and the resulting input/output vectors
are committed under ``tests/golden/`` -- the CPU test-suite re-checks the oracle against
them on every run, and additionally against the two known-answer tests the reference ships
(``tests/test_env.py:10-31``, ``tests/test_map.py:9-34``).

One boundary is pinned on documentation only: ``bbox_masks`` in the reference comes from
``kornia.geometry.boxes.Boxes.from_tensor(.., "xyxy_plus").to_mask`` (general_env.py:373-374;
``requirements.txt:9`` ``kornia>=0.6.12``, un-pinned, not vendored, not installed here) and no
reference test asserts a mask or a reward -> **parity unpinned at the kornia boundary**.  The
oracle uses the documented inclusive-xmax/ymax, clamp-to-image semantics.
"""
