#!/bin/bash
# round 2, call K: full GPU suite, C5 / C5s / C2 bench lines after the sparse and exact-branch changes
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 900 python bench.py --config c5 --steps 100 --warmup 10 > gpurun_out/r2k_bench_c5.json 2> gpurun_out/r2k_bench_c5.err
timeout 600 python bench.py --config c5s --steps 100 --warmup 10 > gpurun_out/r2k_bench_c5s.json 2> gpurun_out/r2k_bench_c5s.err
POGS_B200_PASS_TIMING=1 timeout 600 python bench.py --steps 200 --warmup 20 --no-cpu --no-e2e > gpurun_out/r2k_bench_c2.json 2> gpurun_out/r2k_bench_c2.err
tail -c 300 gpurun_out/r2k_bench_c5.err gpurun_out/r2k_bench_c5s.err gpurun_out/r2k_bench_c2.err
python - <<'PY'
import json
for f in ("r2k_bench_c5","r2k_bench_c5s","r2k_bench_c2"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        c=d.get("converged") or {}
        print(f, round(d["value"],1), round(d["ms_per_step"]*1e3,1), "e2e", (d.get("e2e") or {}).get("value"), "conv", c.get("value"), c.get("iterations"), "cpu", (d.get("cpu_baseline") or {}).get("value"), d["roofline"].get("frac"), d["sanity"].get("parity"))
    except Exception as e:
        print(f, "ERR", e)
PY
