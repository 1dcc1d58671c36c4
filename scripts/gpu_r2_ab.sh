#!/bin/bash
# round 2, call AB: launch list of the one-time setup (C2) -- where the factorisation spends its time; sparse tests
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sparse.py -m gpu -q -x > gpurun_out/r2ab_pytest.log 2>&1
tail -2 gpurun_out/r2ab_pytest.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r02_launches_setup_c2.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-converged > gpurun_out/r2ab_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r02_launches_setup_c2.csv")) if len(r) > 5 and r[0].isdigit()]
agg = collections.defaultdict(list)
for r in rows:
    agg[r[4][:70]].append(float(r[-1]))
tot = sum(sum(v) for v in agg.values())
print("kernels", len(rows), "total ms", tot / 1e6)
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]))[:22]:
    print(f"{k:72s} n={len(v):4d} sum={sum(v)/1e6:8.2f} ms mean={sum(v)/len(v)/1e3:8.1f} us max={max(v)/1e3:8.1f}")
PY
