#!/bin/bash
# round 2, call D (2 GPUs): whole GPU suite incl. the row-block tests, then the 2-GPU bench line
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi -L
timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -40 > gpurun_out/r2d_pytest.log
tail -c 2500 gpurun_out/r2d_pytest.log
POGS_B200_PASS_TIMING=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 20 > gpurun_out/r2d_bench_c2_n2.json 2> gpurun_out/r2d_bench_c2_n2.err
tail -c 600 gpurun_out/r2d_bench_c2_n2.err
python - <<'PY'
import json
for f in ("r2d_bench_c2_n2",):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        c=d.get("converged") or {}
        print(f, round(d["value"],1), round(d["ms_per_step"]*1e3,1), "e2e", d["e2e"]["value"], "conv", c.get("value"), c.get("iterations"), [round(v,1) for v in d["roofline"].get("pass_phase_us") or []], d["sanity"]["k_then_k"])
    except Exception as e:
        print(f, "ERR", e)
PY
