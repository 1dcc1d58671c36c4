#!/bin/bash
# round 2, call E (N GPUs): row-block tests and the N-GPU bench line
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
N=${1:-2}
timeout 1200 python -m pytest tests/test_gpu_dist.py -m gpu -x -q -s 2>&1 | tail -15 | cut -c1-3000 > gpurun_out/r2e_pytest_n$N.log
tail -c 1500 gpurun_out/r2e_pytest_n$N.log
POGS_B200_PASS_TIMING=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 200 --warmup 20 > gpurun_out/r2e_bench_c2_n$N.json 2> gpurun_out/r2e_bench_c2_n$N.err
tail -c 400 gpurun_out/r2e_bench_c2_n$N.err
python - <<PY
import json
for f in ("r2e_bench_c2_n$N",):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        c=d.get("converged") or {}
        print(f, round(d["value"],1), round(d["ms_per_step"]*1e3,1), "e2e", d["e2e"]["value"], "conv", c.get("value"), c.get("iterations"), [round(v,1) for v in d["roofline"].get("pass_phase_us") or []], d["sanity"]["k_then_k"])
    except Exception as e:
        print(f, "ERR", e)
PY
