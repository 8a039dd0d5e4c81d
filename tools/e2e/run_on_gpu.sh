set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt
timeout 200 python tools/e2e/make_colmap_scene.py --out gpurun_out/e2e_scene > gpurun_out/e2e_scene.log 2>&1
rm -rf gpurun_out/e2e_scene_keep; 
for impl in b200-fused b200 ref; do
  timeout 600 python tools/e2e/run_trainer.py --impl $impl --scene gpurun_out/e2e_scene > gpurun_out/e2e_$impl.log 2>&1
  echo "exit $impl $?"
  tail -3 gpurun_out/e2e_$impl.log
done
# do not carry the images back (64 MiB cap)
rm -rf gpurun_out/e2e_scene/images_1 gpurun_out/e2e_*/*.png gpurun_out/e2e_*/*.pt
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v4.csv python tools/profile_step.py --steps 3 > gpurun_out/ncu_launches_v4.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_render -s 4 -c 2 -f -o gpurun_out/prof_render_v4 python tools/profile_step.py --steps 4 > gpurun_out/ncu_full_v4.log 2>&1
ls -la gpurun_out | tail -20
