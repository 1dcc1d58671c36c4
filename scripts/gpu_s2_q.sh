# session 2, call Q (8 GPUs): C2 at N=8 after the parallel epilogue of the sharded factor apply
set -x
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 200 --warmup 20 --no-cpu > gpurun_out/s2q_bench_c2_n8.json 2> gpurun_out/s2q_bench_c2_n8.err
grep '^{' gpurun_out/s2q_bench_c2_n8.json | tail -c 2600; grep -v "OMP_NUM_THREADS\|\*\*\*\*" gpurun_out/s2q_bench_c2_n8.err | tail -8
