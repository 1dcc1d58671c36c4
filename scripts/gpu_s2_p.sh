# session 2, call P (2 GPUs): parallel epilogue of the sharded factor apply: parity subset + bench
set -x
mkdir -p gpurun_out
POGS_DIST_CASES=c2s_lasso_10000x1000,svm_600x200 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_worker.py 2>&1 | grep "RESULT\|Error\|error" | cut -c1-1500
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 200 --warmup 20 --no-cpu --no-e2e > gpurun_out/s2p_bench_c2_n2.json 2> gpurun_out/s2p_bench_c2_n2.err
python - <<'P'
import json
d=json.loads([l for l in open("gpurun_out/s2p_bench_c2_n2.json") if l.startswith("{")][0]); print("N2", round(d["value"],1), d["ms_per_step"], d["roofline"]["phases_ms"], d["sanity"])
P
grep -v "OMP_NUM_THREADS\|\*\*\*\*" gpurun_out/s2p_bench_c2_n2.err | tail -5
