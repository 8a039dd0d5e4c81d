# Three optimizer modes of the view-parallel step, back to back on N GPUs of one box.
# Usage: gpurun --gpus 2 --timeout 1000 -- 'bash tools/e2e/run_view_parallel_modes.sh 2'
n=${1:-2}
mkdir -p gpurun_out
port=29550
for o in torch flat sharded torch; do
  port=$((port+1))
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
      tools/e2e/view_parallel_step.py --steps 16 --optimizer $o --stages > gpurun_out/vp${n}_$o.json 2> gpurun_out/vp${n}_$o.err
  echo "rc=$? $o"; cat gpurun_out/vp${n}_$o.json | cut -c100-700
done
