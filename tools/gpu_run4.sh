#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 -k "not zero_copy" > gpurun_out/pytest_gpu.log 2>&1
echo "pytest_gpu rc=$? $(tail -1 gpurun_out/pytest_gpu.log)" | tee -a gpurun_out/summary.txt
for k in "zero_copy_gather and ldg" "zero_copy_gather and bulk" "zero_copy_gather and tensor" "zero_copy_batch"; do
  timeout 300 python -m pytest tests -m gpu -q -x --timeout 200 -k "$k" > "gpurun_out/pytest_zc_${k// /_}.log" 2>&1
  echo "zc[$k] rc=$? $(tail -1 "gpurun_out/pytest_zc_${k// /_}.log")" | tee -a gpurun_out/summary.txt
done
nvidia-smi --query-gpu=name,clocks.sm --format=csv,noheader | tee -a gpurun_out/summary.txt
timeout 900 python bench.py > gpurun_out/bench_supervised.json 2> gpurun_out/bench_supervised.err
echo "bench supervised rc=$?" | tee -a gpurun_out/summary.txt
timeout 900 python bench.py --src u8 > gpurun_out/bench_supervised_u8.json 2> gpurun_out/bench_supervised_u8.err
echo "bench supervised u8 rc=$?" | tee -a gpurun_out/summary.txt
timeout 900 python bench.py --workload reinforce --steps 5 > gpurun_out/bench_reinforce.json 2> gpurun_out/bench_reinforce.err
echo "bench reinforce rc=$?" | tee -a gpurun_out/summary.txt
timeout 600 python tools/microbench_gather.py --quick --engines tensor,bulk --out gpurun_out/micro_tuned.jsonl > gpurun_out/micro_tuned.log 2>&1
echo "micro rc=$?" | tee -a gpurun_out/summary.txt
cat gpurun_out/summary.txt
