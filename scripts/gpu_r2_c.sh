#!/bin/bash
# round 2, call C: quick timing of the one-launch kernel (C2, C4) after a change
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
POGS_B200_PASS_TIMING=1 timeout 600 python bench.py --steps 200 --warmup 20 --no-cpu --no-e2e > gpurun_out/r2c_bench_c2.json 2> gpurun_out/r2c_bench_c2.err
POGS_B200_PASS_TIMING=1 timeout 600 python bench.py --config c4 --steps 200 --warmup 20 --no-cpu --no-e2e > gpurun_out/r2c_bench_c4.json 2> gpurun_out/r2c_bench_c4.err
timeout 600 python -m pytest tests/test_gpu_scale.py tests/test_gpu_solve.py -m gpu -x -q 2>&1 | tail -5
tail -c 400 gpurun_out/r2c_bench_c2.err gpurun_out/r2c_bench_c4.err
python - <<'PY'
import json
for f in ("r2c_bench_c2","r2c_bench_c4"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        c=d.get("converged") or {}
        print(f, round(d["value"],1), round(d["ms_per_step"]*1e3,1), "conv", c.get("value"), c.get("iterations"), [round(v,1) for v in d["roofline"].get("pass_phase_us") or []], d["sanity"]["k_then_k"])
    except Exception as e:
        print(f, "ERR", e)
PY
