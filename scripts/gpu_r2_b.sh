#!/bin/bash
# round 2, call B: one-launch iteration kernel -- parity suite, then bench lines with / without it
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -60 > gpurun_out/r2b_pytest.log
tail -c 3000 gpurun_out/r2b_pytest.log
POGS_B200_PASS_TIMING=1 timeout 600 python bench.py --steps 200 --warmup 20 --no-cpu > gpurun_out/r2b_bench_c2_k200.json 2> gpurun_out/r2b_bench_c2_k200.err
POGS_B200_MEGA=0 timeout 600 python bench.py --steps 200 --warmup 20 --no-cpu --no-e2e > gpurun_out/r2b_bench_c2_k200_nomega.json 2> gpurun_out/r2b_bench_c2_k200_nomega.err
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench_c2_k20.json 2> gpurun_out/r2b_bench_c2_k20.err
tail -c 600 gpurun_out/r2b_bench_c2_k200.err gpurun_out/r2b_bench_c2_k20.err
python - <<'PY'
import json
for f in ("r2b_bench_c2_k200","r2b_bench_c2_k200_nomega","r2b_bench_c2_k20"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d.get("converged",{}), d["roofline"].get("pass_phase_us"), d["sanity"])
    except Exception as e:
        print(f, "ERR", e)
PY
