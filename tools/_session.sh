D=gpurun_out/${1:-n2}; mkdir -p $D
N=${2:-2}
SECONDS=0
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > $D/bench_n$N.json 2> $D/bench_n$N.err; echo "bench N=$N rc=$? (${SECONDS}s)"
SECONDS=0
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > $D/ref_n$N.json 2> $D/ref_n$N.err; echo "reference N=$N rc=$? (${SECONDS}s)"
{ nvidia-smi topo -m; lscpu | grep -iE "model name|socket|numa|^cpu\(s\)"; free -g | head -2; } > $D/topo.txt 2>&1
tail -2 $D/bench_n$N.err
