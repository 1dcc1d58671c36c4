#!/bin/bash
# round 2, call O: does the one-launch kernel pay for 8 KB rows (C3)?  PDL on / off.
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for MODE in default force; do
  if [ $MODE = force ]; then export POGS_B200_FORCE_FUSE=1; else unset POGS_B200_FORCE_FUSE; fi
  POGS_B200_PASS_TIMING=1 timeout 600 python bench.py --config c3 --steps 400 --warmup 20 --no-cpu --no-e2e > gpurun_out/r2o_bench_c3_$MODE.json 2>/dev/null
  timeout 600 python bench.py --config c3path --no-cpu > gpurun_out/r2o_bench_c3path_$MODE.json 2>/dev/null
done
unset POGS_B200_FORCE_FUSE
POGS_B200_PDL=0 POGS_B200_PASS_TIMING=1 timeout 600 python bench.py --steps 200 --warmup 20 --no-cpu --no-e2e --no-converged > gpurun_out/r2o_bench_c2_nopdl.json 2>/dev/null
POGS_B200_PASS_TIMING=1 timeout 600 python bench.py --steps 200 --warmup 20 --no-cpu --no-e2e --no-converged > gpurun_out/r2o_bench_c2_pdl.json 2>/dev/null
python - <<'PY'
import json
for f in ("r2o_bench_c3_default","r2o_bench_c3_force","r2o_bench_c3path_default","r2o_bench_c3path_force","r2o_bench_c2_nopdl","r2o_bench_c2_pdl"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        c=d.get("converged") or {}
        print(f, round(d["value"],1), round(d["ms_per_step"]*1e3,1), "conv", c.get("value"), c.get("iterations"), d.get("path"), [round(v,1) for v in (d["roofline"].get("pass_phase_us") or [])])
    except Exception as e:
        print(f, "ERR", e)
PY
