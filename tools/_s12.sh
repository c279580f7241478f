mkdir -p gpurun_out/s12
timeout 600 python bench.py --steps 20 --no-cpu-baseline --e2e-timing > gpurun_out/s12/u8.json 2>gpurun_out/s12/err.log; echo rc=$?
timeout 600 python bench.py --steps 20 --no-cpu-baseline > gpurun_out/s12/u8_b.json 2>gpurun_out/s12/err.log; echo rc=$?
timeout 600 python bench.py --workload aerial --steps 10 --no-e2e --no-cpu-baseline > gpurun_out/s12/aerial.json 2>>gpurun_out/s12/err.log; echo rc=$?
