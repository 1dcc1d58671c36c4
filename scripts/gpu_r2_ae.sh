#!/bin/bash
# round 2, call AE: same shared-memory carve-out for every kernel of the factorisation -- timeline + launch list of the setup
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
POGS_B200_TRACE=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-converged > gpurun_out/r2ae_bench.json 2> gpurun_out/r2ae_trace.txt
grep trace gpurun_out/r2ae_trace.txt | tail -19 | head -13
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1300 --csv --log-file gpurun_out/r02_launches_setup_c2.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-converged > gpurun_out/r2ae_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r02_launches_setup_c2.csv")) if len(r) > 5 and r[0].isdigit()]
agg = collections.defaultdict(list)
for r in rows:
    agg[r[4][:70]].append(float(r[-1]))
print("kernels", len(rows))
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]))[:12]:
    print(f"{k:72s} n={len(v):4d} sum={sum(v)/1e6:8.2f} ms mean={sum(v)/len(v)/1e3:8.1f} us max={max(v)/1e3:8.1f}")
PY
