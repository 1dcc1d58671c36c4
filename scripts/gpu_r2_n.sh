#!/bin/bash
# round 2, call N (4 GPUs): row-block parity worker on a subset of cases + the bench-scale case, one-shot entry on 4
# GPUs, 4-GPU bench line
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
N=${1:-4}
POGS_DIST_CASES=c2s_lasso_10000x1000,c4s_logistic_20000x500 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tests/dist_gpu_worker.py 2>&1 | grep RESULT | cut -c1-4000 > gpurun_out/r2n_dist_worker_n$N.log
cat gpurun_out/r2n_dist_worker_n$N.log | cut -c1-2500
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q -k one_shot 2>&1 | tail -3
POGS_B200_PASS_TIMING=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 200 --warmup 20 > gpurun_out/r2n_bench_c2_n$N.json 2> gpurun_out/r2n_bench_c2_n$N.err
tail -c 300 gpurun_out/r2n_bench_c2_n$N.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2n_bench_c2_n$N.json").read().strip().splitlines()[-1])
c=d.get("converged") or {}
print(round(d["value"],1), round(d["ms_per_step"]*1e3,1), "e2e", d["e2e"]["value"], "conv", c.get("value"), [round(v,1) for v in d["roofline"].get("pass_phase_us") or []], d["sanity"]["k_then_k"])
PY
