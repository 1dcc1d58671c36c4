#!/bin/bash
# round 2, call I: warm timeline of one PogsS call (setup with the library's own factorisation), driver-style bench
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
POGS_B200_TRACE=1 timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2i_bench_c2_k20.json 2> gpurun_out/r2i_trace_c2.txt
grep -n "trace" gpurun_out/r2i_trace_c2.txt | tail -45
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2i_bench_c2_k20.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"], d["setup_parts_ms"], d["sanity"]["parity"], d["cpu_baseline"]["value"])
PY
