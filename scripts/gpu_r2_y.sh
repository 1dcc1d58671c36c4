#!/bin/bash
# round 2, call Y: rho-action prediction in the one-launch kernel -- dense parity suites, C2 / C4 lines
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_solve.py tests/test_gpu_scale.py tests/test_gpu_units.py -m gpu -q -x > gpurun_out/r2y_pytest.log 2>&1
tail -4 gpurun_out/r2y_pytest.log
timeout 600 python bench.py --steps 200 --warmup 20 --no-cpu --no-e2e > gpurun_out/r2y_bench_c2.json 2> gpurun_out/r2y_bench_c2.err
timeout 600 python bench.py --config c4 --steps 200 --warmup 20 --no-cpu --no-e2e > gpurun_out/r2y_bench_c4.json 2> gpurun_out/r2y_bench_c4.err
tail -c 300 gpurun_out/r2y_bench_c2.err
python - <<'PY'
import json
for f in ("r2y_bench_c2","r2y_bench_c4"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"],1), round(d["ms_per_step"]*1e3,1), "conv", d.get("converged"), d["sanity"].get("k_then_k"))
    except Exception as e:
        print(f, "ERR", e)
PY
