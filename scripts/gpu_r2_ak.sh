#!/bin/bash
# round 2, call AK: one-shot sparse call (C5) from pinned vs pageable CSR arrays
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for pin in 1 0 1; do
BENCH_SPARSE_PIN=$pin timeout 900 python bench.py --config c5 --steps 100 --warmup 10 --no-cpu --no-converged > gpurun_out/r2ak_bench_c5_pin$pin.json 2> gpurun_out/r2ak_pin$pin.err
python - $pin <<'PY'
import json, sys
d=json.loads(open(f"gpurun_out/r2ak_bench_c5_pin{sys.argv[1]}.json").read().strip().splitlines()[-1])
print("pin", sys.argv[1], round(d["value"],1), d["e2e"]["value"], d["e2e"]["call_s"])
PY
done
