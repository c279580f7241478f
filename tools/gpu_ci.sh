#!/bin/bash
# One GPU session (run under gpurun from the repo root).   Usage: bash tools/gpu_ci.sh <tag> [ncu] [micro]
#   smoke -> GPU parity suite -> default bench line (cfg 3 + cfg 2 / cfg 4 nested) -> reference arm
#   ncu:   launch lists + `--set full` captures of the gather / step kernels at the CONFIG batch sizes, each right
#          after the same command has exited 0 without ncu (profiles/README.md says how they are read;
#          tools/ncu_summarize.py reduces the reports to profiles/<round>/ncu_summary.json, stamped with the
#          source hash written here)
#   micro: cfg-5 microbenches (gather sweep quick, step kernel, fused step call)
D=gpurun_out/${1:-ci}
mkdir -p $D
note() { echo "$@" | tee -a $D/summary.txt; }
rm -f $D/summary.txt
python -c "from jolineedle_b200 import buildinfo; print(buildinfo.library_source_hash())" > $D/source_hash.txt 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv,noheader > $D/gpu.txt 2>&1
{ nvidia-smi topo -m; lscpu | grep -iE "model name|socket|numa|^cpu\(s\)"; free -g | head -2; } > $D/topo.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $D/smoke.log 2>&1; note "smoke rc=$?"
timeout 2400 python -m pytest tests -m gpu -q --timeout 1200 > $D/pytest_gpu.log 2>&1
note "pytest_gpu rc=$? $(tail -1 $D/pytest_gpu.log)"
SECONDS=0
timeout 1200 python bench.py --steps 20 --warmup 5 > $D/bench_default.json 2> $D/bench_default.err; note "bench default rc=$? (${SECONDS}s)"
SECONDS=0
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $D/bench_reference.json 2> $D/bench_reference.err; note "bench reference rc=$? (${SECONDS}s)"

for arg in "$@"; do
if [ "$arg" = "micro" ]; then
  rm -f $D/micro_step.jsonl $D/micro_fused.jsonl $D/micro_quick.jsonl $D/micro_translate.jsonl
  timeout 600 python tools/microbench_step.py --out $D/micro_step.jsonl > $D/micro_step.log 2>&1; note "micro step rc=$?"
  timeout 900 python tools/microbench_step.py --fused --out $D/micro_fused.jsonl > $D/micro_fused.log 2>&1; note "micro fused rc=$?"
  timeout 600 python tools/microbench_gather.py --quick --engines auto --out $D/micro_quick.jsonl > $D/micro_quick.log 2>&1; note "micro gather rc=$?"
  timeout 600 python tools/microbench_gather.py --patches 128,256 --batches 64,256,512 --modes u8 --engines auto --out $D/micro_small.jsonl > $D/micro_small.log 2>&1; note "micro small rc=$?"
  timeout 600 python tools/microbench_gather.py --patches 256 --batches 256,2048 --modes u8 --translate --engines auto --out $D/micro_translate.jsonl > $D/micro_translate.log 2>&1; note "micro translate rc=$?"
fi
if [ "$arg" = "ncu" ]; then
  CMD="python bench.py --steps 2 --warmup 3 --also none --no-e2e --no-cpu-baseline"
  $CMD > $D/plain_rl.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $D/launches_reinforce.csv $CMD > $D/ncu_launches_rl.log 2>&1
  note "ncu launches cfg3 rc=$?"
  $CMD > $D/plain_rl2.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:"env_step|gather_xform" -s 130 -c 4 -o $D/prof_reinforce $CMD > $D/ncu_rl.log 2>&1
  note "ncu cfg3 (B=1024) rc=$?"
  CMD="python bench.py --workload aerial --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
  $CMD > $D/plain_aerial.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $D/launches_aerial.csv $CMD > $D/ncu_launches_aerial.log 2>&1
  note "ncu launches cfg4 rc=$?"
  $CMD > $D/plain_aerial2.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:"env_step|gather_xform" -s 200 -c 4 -o $D/prof_aerial $CMD > $D/ncu_aerial.log 2>&1
  note "ncu cfg4 (B=256) rc=$?"
  CMD="python bench.py --workload supervised --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
  $CMD > $D/plain_sup.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $D/launches_supervised.csv $CMD > $D/ncu_launches_sup.log 2>&1
  note "ncu launches cfg2 rc=$?"
  $CMD > $D/plain_sup2.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:"gather_xform|traj_expand" -s 6 -c 3 -o $D/prof_supervised $CMD > $D/ncu_sup.log 2>&1
  note "ncu cfg2 rc=$?"
  # single-launch captures of the microbenches: Focus gather, float32 pass-through, step kernel on a 45-word grid
  CMD="python tools/microbench_gather.py --patches 448 --batches 2048 --modes u8 --layouts focus --engines auto"
  $CMD > $D/plain_focus.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:gather_xform -s 6 -c 1 -o $D/prof_focus $CMD > $D/ncu_focus.log 2>&1
  note "ncu micro focus rc=$?"
  CMD="python tools/microbench_gather.py --patches 448 --batches 2048 --modes f32 --layouts plain --engines auto"
  $CMD > $D/plain_f32.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:gather_xform -s 6 -c 1 -o $D/prof_f32plain $CMD > $D/ncu_f32.log 2>&1
  note "ncu micro f32 rc=$?"
  CMD="python tools/microbench_step.py --grids 40x36,32x32,8x8 --batches 8192"
  $CMD > $D/plain_step.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:env_step -s 20 -c 1 -o $D/prof_step4036 $CMD > $D/ncu_step.log 2>&1
  note "ncu micro step rc=$?"
fi
done
cat $D/summary.txt
