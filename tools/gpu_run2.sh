#!/bin/bash
# Second GPU session: full parity suite, bench lines, gather tuning sweep, ncu evidence.
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt gpurun_out/tune.jsonl
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest_gpu rc=$? $(tail -1 gpurun_out/pytest_gpu.log)" | tee -a gpurun_out/summary.txt
timeout 900 python bench.py > gpurun_out/bench_supervised.json 2> gpurun_out/bench_supervised.err
echo "bench supervised rc=$?" | tee -a gpurun_out/summary.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
echo "bench reference rc=$?" | tee -a gpurun_out/summary.txt
timeout 900 python bench.py --workload reinforce --steps 5 > gpurun_out/bench_reinforce.json 2> gpurun_out/bench_reinforce.err
echo "bench reinforce rc=$?" | tee -a gpurun_out/summary.txt
# tuning sweep of the copy kernel (stages, lookahead, chunk bytes, CTAs/SM)
for t in "6,3,32768,0" "4,2,32768,0" "3,2,57344,0" "6,4,32768,0" "8,4,24576,0" "8,6,24576,0" "12,6,16384,0" "12,9,16384,0" "4,2,16384,2" "6,3,14336,2" "3,2,28672,2" "4,2,14336,3" "6,4,14336,2"; do
  JN_GATHER_TUNE=$t timeout 300 python tools/microbench_gather.py --patches 448,256 --batches 2048 --modes f32 --layouts plain \
     --engines tensor,bulk --label "copy:$t" --out gpurun_out/tune.jsonl >> gpurun_out/tune.log 2>&1
done
for t in "0,0,0,0,16384,0" "0,0,0,0,28672,0" "0,0,0,0,8192,0" "0,0,0,0,16384,2" "0,0,0,0,49152,0"; do
  JN_GATHER_TUNE=$t timeout 300 python tools/microbench_gather.py --patches 448,256 --batches 2048 --modes f32,u8 --layouts plain,focus \
     --engines tensor,bulk --label "xform:$t" --out gpurun_out/tune.jsonl >> gpurun_out/tune.log 2>&1
done
echo "tune rc=$?" | tee -a gpurun_out/summary.txt
# ncu: launch list of the bench command, then one full capture of the gather kernels
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?" | tee -a gpurun_out/summary.txt
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gather -s 6 -c 4 -o gpurun_out/prof_gather $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?" | tee -a gpurun_out/summary.txt
cat gpurun_out/summary.txt
