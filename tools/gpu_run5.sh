#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest_gpu rc=$? $(tail -1 gpurun_out/pytest_gpu.log)" | tee -a gpurun_out/summary.txt
python tools/profile_step.py f32 > gpurun_out/profile_step_f32.log 2>&1
timeout 900 python bench.py > gpurun_out/bench_supervised.json 2> gpurun_out/bench_supervised.err
echo "bench supervised rc=$?" | tee -a gpurun_out/summary.txt
timeout 900 python bench.py --workload reinforce --steps 5 > gpurun_out/bench_reinforce.json 2> gpurun_out/bench_reinforce.err
echo "bench reinforce rc=$?" | tee -a gpurun_out/summary.txt
# ncu: step kernel + gather in the reinforce workload (small batch keeps the replay short)
CMD="python bench.py --workload reinforce --steps 1 --warmup 3 --batch 256 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain_rl.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"env_step|gather_xform" -s 40 -c 4 -o gpurun_out/prof_rl $CMD > gpurun_out/ncu_rl.log 2>&1
echo "ncu rl rc=$?" | tee -a gpurun_out/summary.txt
CMD2="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD2 > gpurun_out/plain_sup.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"gather_copy|traj_expand" -s 6 -c 3 -o gpurun_out/prof_sup $CMD2 > gpurun_out/ncu_sup.log 2>&1
echo "ncu sup rc=$?" | tee -a gpurun_out/summary.txt
$CMD2 > gpurun_out/plain_sup2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_sup.csv $CMD2 > gpurun_out/ncu_launches_sup.log 2>&1
echo "ncu launches rc=$?" | tee -a gpurun_out/summary.txt
cat gpurun_out/summary.txt
