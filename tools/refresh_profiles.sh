#!/bin/bash
# Reduce one `tools/gpu_ci.sh <tag> ncu micro` session (gpurun_out/<tag>) to the tracked evidence of a round.
#   Usage: bash tools/refresh_profiles.sh <tag> [round]      (run in the build container, after gpurun merged the files)
T=gpurun_out/${1:?tag}; R=profiles/${2:-r02}
H=$(cat $T/source_hash.txt)
[ "$H" = "$(python -c 'from jolineedle_b200 import buildinfo; print(buildinfo.source_hash())')" ] || { echo "session hash $H is not the tree's"; exit 1; }
mkdir -p $R
rm -f $R/ncu_summary.json
python tools/ncu_summarize.py reinforce_u8=$T/prof_reinforce.ncu-rep aerial_u8=$T/prof_aerial.ncu-rep \
  supervised_u8=$T/prof_supervised.ncu-rep --out $R/ncu_summary.json --source-hash $H \
  --note "tools/gpu_ci.sh ncu: python bench.py [--workload ...] --steps 2 --warmup 3 --no-e2e --no-cpu-baseline under ncu --set full at the config batch" | tail -12 | cut -c1-140
python tools/ncu_summarize.py micro_u8_focus_p448_n2048=$T/prof_focus.ncu-rep micro_step_n8192=$T/prof_step4036.ncu-rep \
  micro_f32_plain_p448_n2048=$T/prof_f32plain.ncu-rep --out $R/ncu_summary.json --source-hash $H \
  --note "tools/microbench_*.py under ncu --set full (one launch each)" | grep micro_ | cut -c1-140
python tools/sass_summary.py > $R/sass_summary.txt
cp $T/bench_default.json $R/bench_default_1gpu.json
cp $T/bench_reference.json $R/bench_reference.json
for w in reinforce aerial supervised; do cp $T/launches_$w.csv $R/launches_$w.csv; done
for f in micro_step micro_fused micro_quick micro_small micro_translate; do [ -f $T/$f.jsonl ] && cp $T/$f.jsonl $R/$f.jsonl; done
echo "refreshed $R from $T (source hash $H)"
