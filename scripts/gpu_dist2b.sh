set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py tests/test_gpu_solve.py -m gpu -q --timeout=900 -x -p no:cacheprovider 2>&1 | tail -5
for v in a b; do
  if [ $v = b ]; then export POGS_B200_NO_SHARD=1; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 bench.py --gpus 2 --steps 200 --warmup 20 --no-cpu --no-e2e > gpurun_out/bench_r1_n2$v.json 2> gpurun_out/bench_r1_n2$v.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_r1_n2$v.json')); print('N2$v', d['value'], d['ms_per_step'], d['roofline']['phases_ms'], d['gpu_launches'], d['setup_parts_ms'])"
  grep -v "OMP_NUM_THREADS\|\*\*\*\*" gpurun_out/bench_r1_n2$v.err | tail -5
done
unset POGS_B200_NO_SHARD
python bench.py --steps 200 --warmup 20 --no-cpu > gpurun_out/bench_r1_n1c.json 2> gpurun_out/bench_r1_n1c.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r1_n1c.json')); print('N1', d['value'], d['ms_per_step'], d['roofline']['phases_ms'], d['gpu_launches'], d['e2e'])"
tail -3 gpurun_out/bench_r1_n1c.err
python bench.py --config c3 --steps 500 --warmup 20 --no-cpu --no-e2e > gpurun_out/bench_r1_c3b.json 2> gpurun_out/bench_r1_c3b.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r1_c3b.json')); print('C3', d['value'], d['ms_per_step'], d['roofline']['phases_ms'], d['gpu_launches'])"
POGS_B200_NO_COND=1 python bench.py --config c3 --steps 500 --warmup 20 --no-cpu --no-e2e > gpurun_out/bench_r1_c3c.json 2> gpurun_out/bench_r1_c3c.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r1_c3c.json')); print('C3 nocond', d['value'], d['ms_per_step'], d['gpu_launches'])"
tail -3 gpurun_out/bench_r1_c3b.err
