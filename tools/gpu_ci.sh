#!/bin/bash
# Runs on the GPU box (under gpurun): smoke, the GPU parity suites one process per file (a CUDA
# fault in one file must not poison the others), then the quick gather microbench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/summary.txt
for f in tests/test_gather_gpu.py tests/test_general_env_gpu.py tests/test_simple_env_gpu.py tests/test_returns_gpu.py; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -m gpu -q -x --timeout 600 > gpurun_out/$n.log 2>&1
  echo "$n rc=$? $(tail -1 gpurun_out/$n.log)" | tee -a gpurun_out/summary.txt
done
timeout 600 python tools/microbench_gather.py --quick --engines tensor,bulk,ldg --out gpurun_out/micro_quick.jsonl > gpurun_out/micro_quick.log 2>&1
echo "micro rc=$?" | tee -a gpurun_out/summary.txt
cat gpurun_out/summary.txt
