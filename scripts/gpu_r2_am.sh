#!/bin/bash
# round 2, call AM (4 GPUs): the final build on the row-block path
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 200 --warmup 20 --no-cpu > gpurun_out/r2am_bench_c2_n4.json 2> gpurun_out/r2am_bench_c2_n4.err
tail -c 300 gpurun_out/r2am_bench_c2_n4.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2am_bench_c2_n4.json").read().strip().splitlines()[-1])
print(round(d["value"],1), round(d["ms_per_step"]*1e3,1), "e2e", d["e2e"]["value"], "conv", (d.get("converged") or {}).get("value"), d["sanity"].get("k_then_k"))
PY
