#!/bin/bash
# round 2, call F (N GPUs): bench lines only (C2 and C4)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
N=${1:-8}
for CFG in c2 c4; do
POGS_B200_PASS_TIMING=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 200 --warmup 20 --config $CFG > gpurun_out/r2f_bench_${CFG}_n$N.json 2> gpurun_out/r2f_bench_${CFG}_n$N.err
tail -c 300 gpurun_out/r2f_bench_${CFG}_n$N.err
done
python - <<PY
import json
for f in ("r2f_bench_c2_n$N","r2f_bench_c4_n$N"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        c=d.get("converged") or {}
        print(f, round(d["value"],1), round(d["ms_per_step"]*1e3,1), "e2e", d["e2e"]["value"], "conv", c.get("value"), c.get("iterations"), [round(v,1) for v in d["roofline"].get("pass_phase_us") or []], d["sanity"]["k_then_k"])
    except Exception as e:
        print(f, "ERR", e)
PY
