# session 2, call M (4 GPUs): row-block bench at N=4 (strong scaling of C2), short
set -x
mkdir -p gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29504 bench.py --gpus 4 --steps 200 --warmup 20 --no-cpu > gpurun_out/s2m_bench_c2_n4.json 2> gpurun_out/s2m_bench_c2_n4.err
grep '^{' gpurun_out/s2m_bench_c2_n4.json | tail -c 2600; grep -v "OMP_NUM_THREADS\|\*\*\*\*" gpurun_out/s2m_bench_c2_n4.err | tail -8
