#!/bin/bash
# round 2, call AN (8 GPUs): the final build on the row-block path
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 200 --warmup 20 --no-cpu --no-e2e > gpurun_out/r2an_bench_c2_n8.json 2> gpurun_out/r2an_bench_c2_n8.err
tail -c 300 gpurun_out/r2an_bench_c2_n8.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2an_bench_c2_n8.json").read().strip().splitlines()[-1])
print(round(d["value"],1), round(d["ms_per_step"]*1e3,1), "conv", (d.get("converged") or {}).get("value"), d["sanity"].get("k_then_k"))
PY
