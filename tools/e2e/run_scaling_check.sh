# bench.py and the sharded view-parallel step on N GPUs of one box.
# Usage: gpurun --gpus N --timeout 900 -- 'bash tools/e2e/run_scaling_check.sh N'
n=${1:-4}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29571 \
    bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/bench_${n}gpu.json 2> gpurun_out/bench_${n}gpu.err
echo "bench rc=$?"; cut -c1-330 gpurun_out/bench_${n}gpu.json
for o in sharded flat; do
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29572 \
    tools/e2e/view_parallel_step.py --steps 12 --optimizer $o > gpurun_out/vp${n}_$o.json 2> gpurun_out/vp${n}_$o.err
echo "vp $o rc=$?"; cut -c100-520 gpurun_out/vp${n}_$o.json
done
