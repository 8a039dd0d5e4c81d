# Round-end style check on a B200 box: GPU tests, the bench line (both arms), the 3 M parity report, ncu launch
# list + full capture of the two tile kernels.  Usage: gpurun --timeout 2000 -- 'bash tools/run_gpu_check.sh [tag]'
tag=${1:-latest}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_b200_$tag.json 2> gpurun_out/bench_b200.err; echo "bench rc=$?"; head -c 400 gpurun_out/bench_b200_$tag.json; echo
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_reference_$tag.json 2> gpurun_out/bench_reference.err; echo "bench ref rc=$?"; head -c 300 gpurun_out/bench_reference_$tag.json; echo
timeout 300 python tools/parity_report.py --n 3000000 --res 1080p --out gpurun_out/parity_3M_1080p_$tag.json > gpurun_out/parity_3M.log 2>&1; echo "parity rc=$?"; tail -4 gpurun_out/parity_3M.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv python tools/profile_step.py --steps 3 > gpurun_out/ncu_launches.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_render -s 4 -c 2 -f -o gpurun_out/prof_render_$tag python tools/profile_step.py --steps 4 > gpurun_out/ncu_full.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
ls -la gpurun_out | grep "$tag"
