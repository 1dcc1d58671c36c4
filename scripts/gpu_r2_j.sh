#!/bin/bash
# round 2, call J: factorisation variants (unit tests), warm setup timeline, driver-style bench line
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_units.py tests/test_gpu_solve.py -m gpu -x -q 2>&1 | tail -4
POGS_B200_TRACE=1 timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2j_bench_c2_k20.json 2> gpurun_out/r2j_trace_c2.txt
grep -n "trace" gpurun_out/r2j_trace_c2.txt | sed -n 57,73p
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2j_bench_c2_k20.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["e2e"]["call_s"], d["sanity"]["parity"], d["cpu_baseline"]["value"])
PY
