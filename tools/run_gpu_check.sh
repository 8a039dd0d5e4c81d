# Round-end style check on a B200 box: GPU tests, the bench line, ncu launch list + full capture of the two
# tile kernels.  Usage: gpurun --timeout 1500 -- 'bash tools/run_gpu_check.sh [tag]'
tag=${1:-latest}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_b200_$tag.json 2> gpurun_out/bench_b200.err; echo "bench rc=$?"; head -c 600 gpurun_out/bench_b200_$tag.json; echo
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv python tools/profile_step.py --steps 3 > gpurun_out/ncu_launches.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_render -s 4 -c 2 -f -o gpurun_out/prof_render_$tag python tools/profile_step.py --steps 4 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | grep "$tag"
