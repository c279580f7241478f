mkdir -p gpurun_out/s7
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/s7/pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/s7/pytest.log)"
timeout 300 python tools/profile_step.py u8 > gpurun_out/s7/profile_u8.log 2>&1
timeout 300 python bench.py --src u8 --steps 20 --no-e2e --no-cpu-baseline > gpurun_out/s7/u8_20.json 2>gpurun_out/s7/err.log; echo rc=$?
timeout 300 python bench.py --src u8 --steps 100 --no-e2e --no-cpu-baseline > gpurun_out/s7/u8_100.json 2>gpurun_out/s7/err.log; echo rc=$?
timeout 300 python bench.py --steps 20 --no-e2e --no-cpu-baseline > gpurun_out/s7/f32_20.json 2>>gpurun_out/s7/err.log; echo rc=$?
