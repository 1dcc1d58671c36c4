set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
nvidia-smi topo -m | head -12
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -q --timeout=900 -x -p no:cacheprovider -k "world0 or 2" 2>&1 | tail -30
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 10 --no-cpu > gpurun_out/bench_r1_n2.json 2> gpurun_out/bench_r1_n2.err; tail -c 2500 gpurun_out/bench_r1_n2.json; tail -5 gpurun_out/bench_r1_n2.err
timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu > gpurun_out/bench_r1_n1b.json 2> gpurun_out/bench_r1_n1b.err; tail -c 2500 gpurun_out/bench_r1_n1b.json; tail -5 gpurun_out/bench_r1_n1b.err
