#!/bin/bash
# round 2, call AA: vectorised double-buffered fp32 GEMM of the one-time factorisation -- factor / scale / sparse tests,
# warm timeline of one PogsS call (C2, K = 20), driver-style bench line
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_units.py tests/test_gpu_scale.py tests/test_gpu_sparse.py -m gpu -q -x > gpurun_out/r2aa_pytest.log 2>&1
tail -4 gpurun_out/r2aa_pytest.log
POGS_B200_TRACE=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-converged > gpurun_out/r2aa_bench_c2_k20_trace.json 2> gpurun_out/r2aa_trace.txt
grep trace gpurun_out/r2aa_trace.txt | tail -22
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2aa_bench_c2_k20.json 2> gpurun_out/r2aa_bench_c2_k20.err
python - <<'PY'
import json
for f in ("r2aa_bench_c2_k20",):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"],1), round(d["ms_per_step"]*1e3,1), "e2e", d.get("e2e"), "setup", d.get("setup_ms"), d.get("setup_parts_ms"), d["sanity"])
    except Exception as e:
        print(f, "ERR", e)
PY
