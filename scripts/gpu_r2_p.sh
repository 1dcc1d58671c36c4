#!/bin/bash
# round 2, call P: suite + C3 lines with the one-launch kernel enabled from 8 KB rows
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6
POGS_B200_PASS_TIMING=1 timeout 600 python bench.py --config c3 --steps 400 --warmup 20 > gpurun_out/r2p_bench_c3.json 2>/dev/null
timeout 600 python bench.py --config c3path > gpurun_out/r2p_bench_c3path.json 2>/dev/null
python - <<'PY'
import json
for f in ("r2p_bench_c3","r2p_bench_c3path"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        c=d.get("converged") or {}
        print(f, round(d["value"],1), round(d["ms_per_step"]*1e3,1), "e2e", (d.get("e2e") or {}).get("value"), "conv", c.get("value"), d.get("path"), d["sanity"].get("parity"), (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(f, "ERR", e)
PY
