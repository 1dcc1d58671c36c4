#!/bin/bash
# round 2, call X: tiled layout as the default, own CSR->CSC transpose (no cuSPARSE): sparse + unit tests, C5 / C5s lines,
# launch list and ncu --set full of the final product kernel, build timeline of a first solver
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out /tmp/prof
timeout 900 python -m pytest tests/test_gpu_sparse.py tests/test_gpu_units.py tests/test_capi.py -m gpu -q -x > gpurun_out/r2x_pytest.log 2>&1
tail -3 gpurun_out/r2x_pytest.log
POGS_B200_TRACE=1 timeout 600 python scripts/dev/c5_probe.py c5 tiled > gpurun_out/r2x_probe.log 2>&1
grep -v trace gpurun_out/r2x_probe.log | tail -3; grep trace gpurun_out/r2x_probe.log | head -12
timeout 900 python bench.py --config c5 --steps 100 --warmup 10 > gpurun_out/r2x_bench_c5.json 2> gpurun_out/r2x_bench_c5.err
timeout 600 python bench.py --config c5s --steps 100 --warmup 10 > gpurun_out/r2x_bench_c5s.json 2> gpurun_out/r2x_bench_c5s.err
tail -c 300 gpurun_out/r2x_bench_c5.err
python - <<'PY'
import json
for f in ("r2x_bench_c5","r2x_bench_c5s"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        c=d.get("converged") or {}
        print(f, round(d["value"],1), round(d["ms_per_step"]*1e3,1), "e2e", (d.get("e2e") or {}).get("value"), "conv", c.get("value"), c.get("iterations"), "cpu", (d.get("cpu_baseline") or {}).get("value"), d["roofline"].get("frac"), d["sanity"].get("parity"), d["gpu_launches"])
    except Exception as e:
        print(f, "ERR", e)
PY
POGS_B200_NO_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 300 --csv --log-file gpurun_out/r02_launches_c5_tiled.csv python bench.py --config c5 --steps 12 --warmup 3 --no-cpu --no-e2e --no-converged > gpurun_out/r2x_ncu_list.log 2>&1
POGS_B200_NO_GRAPH=1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_spmv_tiled -s 120 -c 2 -o /tmp/prof/spmvt -f python bench.py --config c5 --steps 6 --warmup 3 --no-cpu --no-e2e --no-converged > gpurun_out/r2x_ncu_full.log 2>&1
ncu -i /tmp/prof/spmvt.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_spmv_tiled_c5.csv 2>/dev/null
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r02_launches_c5_tiled.csv")) if len(r) > 5 and r[0].isdigit()]
agg = collections.defaultdict(list)
for r in rows:
    agg[r[4][:60]].append(float(r[-1]))
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:62s} n={len(v):4d} mean={sum(v)/len(v)/1e3:9.1f} us min={min(v)/1e3:9.1f} max={max(v)/1e3:9.1f}")
PY
