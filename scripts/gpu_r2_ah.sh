#!/bin/bash
# round 2, call AH: y = t_y - r from the CGLS residual recurrence (one product less per ADMM iteration) -- sparse / indirect
# tests, C5 loop time with and without it, C5 / C5s lines
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sparse.py tests/test_gpu_scale.py tests/test_gpu_cone.py -m gpu -q -x > gpurun_out/r2ah_pytest.log 2>&1
tail -4 gpurun_out/r2ah_pytest.log
timeout 600 python scripts/dev/c5_probe.py c5 tiled 2>&1 | tail -3
POGS_B200_Y_REC=0 timeout 600 python scripts/dev/c5_probe.py c5 tiled 2>&1 | tail -2
timeout 900 python bench.py --config c5 --steps 100 --warmup 10 > gpurun_out/r2ah_bench_c5.json 2> gpurun_out/r2ah_bench_c5.err
timeout 600 python bench.py --config c5s --steps 100 --warmup 10 > gpurun_out/r2ah_bench_c5s.json 2> gpurun_out/r2ah_bench_c5s.err
tail -c 300 gpurun_out/r2ah_bench_c5.err
python - <<'PY'
import json
for f in ("r2ah_bench_c5","r2ah_bench_c5s"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        c=d.get("converged") or {}
        print(f, round(d["value"],1), round(d["ms_per_step"]*1e3,1), "e2e", (d.get("e2e") or {}).get("value"), "conv", c.get("value"), c.get("iterations"), "cpu", (d.get("cpu_baseline") or {}).get("value"), d["roofline"].get("frac"), d["sanity"].get("parity"), d["gpu_launches"])
    except Exception as e:
        print(f, "ERR", e)
PY
