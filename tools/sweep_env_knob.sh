#!/bin/bash
# A/B of one environment knob of the library on the gather microbench and the cfg 3 / cfg 4 steps.
#   Usage (on a GPU box): bash tools/sweep_env_knob.sh <VAR> "<v1 v2 ...>" <out.jsonl>
#   e.g. JN_TMA_L2_PROMOTION "256 128 64 none" (profiles/r02/sweep_l2_promotion.jsonl), JN_L2_HINT "0 1 2 3"
VAR=$1; VALUES=$2; OUT=${3:-gpurun_out/sweep_$1.jsonl}; rm -f $OUT
for v in $VALUES; do
  export $VAR=$v
  T=$(mktemp)
  python tools/microbench_gather.py --patches 448 --batches 1024,2048 --modes u8,f32 --layouts plain,focus --engines auto --out $T > /dev/null 2>&1
  python tools/microbench_gather.py --patches 256 --batches 256,2048 --modes u8 --layouts plain --engines auto --out $T > /dev/null 2>&1
  python tools/microbench_gather.py --patches 256 --batches 256,2048 --modes u8 --translate --engines auto --out $T > /dev/null 2>&1
  python tools/microbench_gather.py --patches 128 --batches 2048 --modes u8 --layouts plain --engines auto --out $T > /dev/null 2>&1
  sed "s/^{/{\"$VAR\": \"$v\", /" $T >> $OUT
  for wl in reinforce aerial; do
    python bench.py --workload $wl --also none --no-e2e --no-cpu-baseline --steps 10 --warmup 3 2>/dev/null |
      python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print(json.dumps({'$VAR': '$v', 'bench': '$wl', 'value': round(d['value']), 'frac': d['roofline']['frac'], 'avg_launch_ms': d['roofline']['avg_launch_ms']}))" >> $OUT
  done
  rm -f $T
done
