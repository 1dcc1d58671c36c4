# session 2, call O (8 GPUs): row-block bench at N=8 (strong scaling of C2) and C4 (BASELINE's 8-GPU config)
set -x
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29508 bench.py --gpus 8 --steps 200 --warmup 20 --no-cpu > gpurun_out/s2o_bench_c2_n8.json 2> gpurun_out/s2o_bench_c2_n8.err
grep '^{' gpurun_out/s2o_bench_c2_n8.json | tail -c 2600; grep -v "OMP_NUM_THREADS\|\*\*\*\*" gpurun_out/s2o_bench_c2_n8.err | tail -8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29509 bench.py --gpus 8 --config c4 --steps 200 --warmup 20 --no-cpu --no-e2e > gpurun_out/s2o_bench_c4_n8.json 2> gpurun_out/s2o_bench_c4_n8.err
grep '^{' gpurun_out/s2o_bench_c4_n8.json | tail -c 1800; grep -v "OMP_NUM_THREADS\|\*\*\*\*" gpurun_out/s2o_bench_c4_n8.err | tail -8
