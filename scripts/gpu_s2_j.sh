# session 2, call J (2 GPUs): row-block tests + 2-GPU bench with the tensor-core Gram, fused setup
# passes, pooled allocator; single-GPU e2e timeline with the fp32 factor
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -q --timeout=900 -x -p no:cacheprovider 2>&1 | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 bench.py --gpus 2 --steps 200 --warmup 20 --no-cpu > gpurun_out/s2j_bench_c2_n2.json 2> gpurun_out/s2j_bench_c2_n2.err
grep '^{' gpurun_out/s2j_bench_c2_n2.json | tail -c 2500; grep -v "OMP_NUM_THREADS\|\*\*\*\*" gpurun_out/s2j_bench_c2_n2.err | tail -5
POGS_B200_TRACE=1 python bench.py --steps 200 --warmup 20 --no-cpu > gpurun_out/s2j_bench_c2_trace.json 2> gpurun_out/s2j_trace_c2.txt; grep "trace:" gpurun_out/s2j_trace_c2.txt | tail -17
python bench.py --steps 200 --warmup 20 --no-cpu > gpurun_out/s2j_bench_c2.json 2> gpurun_out/s2j_bench_c2.err; tail -c 1200 gpurun_out/s2j_bench_c2.json; tail -3 gpurun_out/s2j_bench_c2.err
timeout 600 python -m pytest tests/test_gpu_units.py tests/test_gpu_solve.py -m gpu -q --timeout=300 -p no:cacheprovider 2>&1 | tail -5
