#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest_gpu rc=$? $(tail -1 gpurun_out/pytest_gpu.log)" | tee -a gpurun_out/summary.txt
timeout 900 python bench.py --workload aerial > gpurun_out/bench_aerial.json 2> gpurun_out/bench_aerial.err
echo "bench aerial rc=$?" | tee -a gpurun_out/summary.txt
cat gpurun_out/summary.txt
