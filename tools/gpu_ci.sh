#!/bin/bash
# One GPU session (run under gpurun from the repo root):
#   smoke -> GPU parity suite -> bench lines (cfg 2 uint8 / fp32, cfg 3, cfg 4, reference arm)
#   -> quick gather microbench.   Usage: bash tools/gpu_ci.sh [ncu]
# With "ncu" it also captures the launch list and full profiles of the gather / step kernels, each
# right after the same command has exited 0 without ncu (profiles/README.md says how they are read;
# tools/ncu_summarize.py reduces the reports to profiles/r01/ncu_summary.json).
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
note() { echo "$@" | tee -a gpurun_out/summary.txt; }

nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv,noheader > gpurun_out/gpu.txt 2>&1
{ nvidia-smi topo -m; lscpu | grep -iE "model name|socket|numa|^cpu\(s\)"; python -c "import os; print('affinity', sorted(os.sched_getaffinity(0)))"; free -g | head -2; } > gpurun_out/topo.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; note "smoke rc=$?"
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu.log 2>&1
note "pytest_gpu rc=$? $(tail -1 gpurun_out/pytest_gpu.log)"

timeout 900 python bench.py > gpurun_out/bench_supervised.json 2> gpurun_out/bench_supervised.err; note "bench cfg2 u8 (default) rc=$?"
timeout 900 python bench.py --src f32 > gpurun_out/bench_supervised_f32.json 2> gpurun_out/bench_supervised_f32.err; note "bench cfg2 f32 rc=$?"
timeout 900 python bench.py --workload reinforce > gpurun_out/bench_reinforce.json 2> gpurun_out/bench_reinforce.err; note "bench cfg3 rc=$?"
timeout 900 python bench.py --workload aerial > gpurun_out/bench_aerial.json 2> gpurun_out/bench_aerial.err; note "bench cfg4 rc=$?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; note "bench reference rc=$?"
rm -f gpurun_out/micro_quick.jsonl gpurun_out/micro_translate.jsonl
timeout 600 python tools/microbench_gather.py --quick --engines tensor,bulk,auto --out gpurun_out/micro_quick.jsonl > gpurun_out/micro_quick.log 2>&1; note "micro rc=$?"
timeout 600 python tools/microbench_gather.py --quick --batches 2048 --translate --engines auto,ldg --out gpurun_out/micro_translate.jsonl > gpurun_out/micro_translate.log 2>&1; note "micro translate rc=$?"

if [ "$1" = "ncu" ]; then
  CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
  $CMD > gpurun_out/plain_sup.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_sup.csv $CMD > gpurun_out/ncu_launches_sup.log 2>&1
  note "ncu launches rc=$?"
  $CMD > gpurun_out/plain_sup2.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:"gather_xform|traj_expand" -s 6 -c 3 -o gpurun_out/prof_sup_u8 $CMD > gpurun_out/ncu_sup.log 2>&1
  note "ncu supervised u8 rc=$?"
  CMD="python bench.py --src f32 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
  $CMD > gpurun_out/plain_sup_f32.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:"gather_copy" -s 6 -c 2 -o gpurun_out/prof_sup_f32 $CMD > gpurun_out/ncu_sup_f32.log 2>&1
  note "ncu supervised f32 rc=$?"
  CMD="python bench.py --workload reinforce --steps 1 --warmup 3 --batch 256 --no-e2e --no-cpu-baseline"
  $CMD > gpurun_out/plain_rl.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:"env_step|gather_xform" -s 130 -c 4 -o gpurun_out/prof_rl $CMD > gpurun_out/ncu_rl.log 2>&1
  note "ncu reinforce rc=$?"
  CMD="python bench.py --workload aerial --steps 1 --warmup 3 --batch 64 --no-e2e --no-cpu-baseline"
  $CMD > gpurun_out/plain_aerial.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:"env_step|gather_xform" -s 200 -c 4 -o gpurun_out/prof_aerial $CMD > gpurun_out/ncu_aerial.log 2>&1
  note "ncu aerial rc=$?"
fi
cat gpurun_out/summary.txt
