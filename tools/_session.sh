D=gpurun_out/${1:-s5}; mkdir -p $D
timeout 2400 python -m pytest tests -m gpu -q -x --timeout 1200 > $D/pytest_gpu.log 2>&1; echo "pytest rc=$? $(tail -1 $D/pytest_gpu.log)"
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $D/bench_default.json 2> $D/bench_default.err; echo "default rc=$?"
rm -f $D/*.jsonl
timeout 600 python tools/microbench_gather.py --quick --engines auto --out $D/micro_quick.jsonl > $D/micro_quick.log 2>&1; echo "micro rc=$?"
timeout 600 python tools/microbench_gather.py --patches 128,256 --batches 64,256,512 --modes u8 --engines auto --out $D/micro_small.jsonl > $D/micro_small.log 2>&1
timeout 600 python tools/microbench_gather.py --patches 256 --batches 256,2048 --modes u8 --translate --engines auto --out $D/micro_translate.jsonl > $D/micro_translate.log 2>&1
timeout 600 python tools/microbench_gather.py --patches 448 --batches 1024 --modes u8 --layouts plain --engines auto,tensor,bulk --out $D/micro_1024.jsonl > $D/micro_1024.log 2>&1
echo done
