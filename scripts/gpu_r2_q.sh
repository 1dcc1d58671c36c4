#!/bin/bash
# round 2, call Q: two-level Cholesky -- unit/scale tests, warm setup timeline, driver-style lines (product + reference arm), smoke
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_units.py tests/test_gpu_scale.py tests/test_gpu_solve.py -m gpu -q -s 2>&1 | grep -E "fp64 6000|passed|failed|FAILED" | tail -8
POGS_B200_TRACE=1 POGS_B200_PASS_TIMING=1 timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2q_bench_c2_k20.json 2> gpurun_out/r2q_trace_c2.txt
grep -n "trace" gpurun_out/r2q_trace_c2.txt | sed -n 57,73p
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2q_bench_ref_k20.json 2>/dev/null
python __graft_entry__.py smoke 2>&1 | tail -3
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2q_bench_c2_k20.json").read().strip().splitlines()[-1])
r=json.loads(open("gpurun_out/r2q_bench_ref_k20.json").read().strip().splitlines()[-1])
print("ours", d["value"], "e2e", d["e2e"]["value"], d["e2e"]["call_s"], "roofline", d["roofline"]["frac"], d["roofline"]["achieved"], "conv", d["converged"]["value"], "parity", d["sanity"]["parity"]["ok"], d["sanity"]["k_then_k"])
print("ref", r["value"], "e2e", r["e2e"]["value"], r["sanity"]["k_then_k"])
print("ratio", d["value"]/r["value"], "e2e ratio", d["e2e"]["value"]/r["e2e"]["value"])
PY
