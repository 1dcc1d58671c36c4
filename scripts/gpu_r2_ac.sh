#!/bin/bash
# round 2, call AC: diagonal-block kernel with the inverse in the same sweep -- factor / scale tests, warm timeline of one
# PogsS call (C2, K = 20), driver-style bench line
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_units.py tests/test_gpu_scale.py -m gpu -q -x > gpurun_out/r2ac_pytest.log 2>&1
tail -3 gpurun_out/r2ac_pytest.log
POGS_B200_TRACE=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-converged > gpurun_out/r2ac_bench_c2_k20_trace.json 2> gpurun_out/r2ac_trace.txt
grep trace gpurun_out/r2ac_trace.txt | tail -17 | head -12
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2ac_bench_c2_k20.json 2> gpurun_out/r2ac_bench_c2_k20.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2ac_bench_c2_k20.json").read().strip().splitlines()[-1])
print(round(d["value"],1), "e2e", d.get("e2e"), "conv", (d.get("converged") or {}).get("value"), d["sanity"].get("parity"))
PY
