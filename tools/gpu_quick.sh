#!/bin/bash
# Quick GPU session: smoke + parity suite + bench lines.   Usage: bash tools/gpu_quick.sh [tag]
D=gpurun_out/${1:-q}
mkdir -p $D
python -c "import __graft_entry__ as g; g.smoke()" > $D/smoke.log 2>&1; echo "smoke rc=$?"
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > $D/pytest_gpu.log 2>&1; echo "pytest rc=$? $(tail -1 $D/pytest_gpu.log)"
timeout 900 python bench.py --workload aerial --steps 20 --no-cpu-baseline > $D/aerial.json 2> $D/aerial.err; echo "aerial rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > $D/bench_default.json 2> $D/bench_default.err; echo "default rc=$?"
tail -5 $D/pytest_gpu.log
