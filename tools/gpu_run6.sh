#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest_gpu rc=$? $(tail -1 gpurun_out/pytest_gpu.log)" | tee -a gpurun_out/summary.txt
python tools/profile_step.py f32 > gpurun_out/profile_step_f32.log 2>&1
timeout 900 python bench.py > gpurun_out/bench_supervised.json 2> gpurun_out/bench_supervised.err
echo "bench supervised rc=$?" | tee -a gpurun_out/summary.txt
timeout 900 python bench.py --src u8 > gpurun_out/bench_supervised_u8.json 2> gpurun_out/bench_supervised_u8.err
echo "bench supervised u8 rc=$?" | tee -a gpurun_out/summary.txt
timeout 600 python tools/microbench_zero_copy.py > gpurun_out/zero_copy.jsonl 2> gpurun_out/zero_copy.err
echo "zero copy micro rc=$?" | tee -a gpurun_out/summary.txt
cat gpurun_out/summary.txt
