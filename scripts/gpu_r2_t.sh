#!/bin/bash
# round 2, call T: 2-D tiled sparse product (sparse_tiled.cuh) -- sparse tests, C5 lines tiled vs blocked
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sparse.py -m gpu -q -x 2>&1 | tail -15
POGS_B200_SPMV=tiled timeout 900 python bench.py --config c5 --steps 100 --warmup 10 > gpurun_out/r2t_bench_c5_tiled.json 2> gpurun_out/r2t_bench_c5_tiled.err
timeout 900 python bench.py --config c5 --steps 100 --warmup 10 --no-cpu > gpurun_out/r2t_bench_c5_blocked.json 2> gpurun_out/r2t_bench_c5_blocked.err
POGS_B200_SPMV=tiled timeout 600 python bench.py --config c5s --steps 100 --warmup 10 --no-cpu > gpurun_out/r2t_bench_c5s_tiled.json 2> gpurun_out/r2t_bench_c5s_tiled.err
tail -c 600 gpurun_out/r2t_bench_c5_tiled.err
python - <<'PY'
import json
for f in ("r2t_bench_c5_tiled","r2t_bench_c5_blocked","r2t_bench_c5s_tiled"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        c=d.get("converged") or {}
        print(f, round(d["value"],1), round(d["ms_per_step"]*1e3,1), "e2e", (d.get("e2e") or {}).get("value"), "conv", c.get("value"), c.get("iterations"), "cpu", (d.get("cpu_baseline") or {}).get("value"), d["roofline"].get("frac"), d["sanity"].get("parity"))
    except Exception as e:
        print(f, "ERR", e)
PY
