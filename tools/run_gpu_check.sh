mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_b200.json 2> gpurun_out/bench_b200.err; echo "bench rc=$?"; cat gpurun_out/bench_b200.json | head -c 3000
