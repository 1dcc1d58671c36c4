set -x
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "GPUs: $N"
if [ "$N" -le 2 ]; then
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -q --timeout=900 -x -p no:cacheprovider 2>&1 | tail -5
fi
for n in 2 4 8; do
  if [ "$n" -le "$N" ]; then
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 200 --warmup 20 --no-cpu > gpurun_out/bench_r1_n$n.json 2> gpurun_out/bench_r1_n$n.err
    tail -c 1700 gpurun_out/bench_r1_n$n.json; grep -v "OMP_NUM_THREADS\|\*\*\*\*" gpurun_out/bench_r1_n$n.err | tail -5
  fi
done
