#!/bin/bash
# round 2, call U: tiled sparse product (TMA ring, two kernels) -- tests, loop time, per-kernel launch list
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sparse.py -m gpu -q -x > gpurun_out/r2u_pytest.log 2>&1
tail -5 gpurun_out/r2u_pytest.log
timeout 600 python scripts/dev/c5_probe.py c5 tiled blocked > gpurun_out/r2u_probe.log 2>&1
tail -20 gpurun_out/r2u_probe.log
POGS_B200_SPMV=tiled POGS_B200_NO_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 200 --csv --log-file gpurun_out/r2u_launches_c5_tiled.csv python bench.py --config c5 --steps 12 --warmup 3 --no-cpu --no-e2e --no-converged > gpurun_out/r2u_ncu_list.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r2u_launches_c5_tiled.csv")) if len(r) > 5 and r[0].isdigit()]
agg = collections.defaultdict(list)
for r in rows:
    agg[r[4][:60]].append(float(r[-1]))
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:62s} n={len(v):4d} mean={sum(v)/len(v)/1e3:9.1f} us min={min(v)/1e3:9.1f} max={max(v)/1e3:9.1f}")
PY
