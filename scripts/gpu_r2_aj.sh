#!/bin/bash
# round 2, call AJ: ncu evidence of the final build -- launch list of the captured C2 loop, --set full of k_admm_pass on C2 and
# C4 (after the rho-action prediction), launch list of the C5 loop (three products per iteration); corrected C5 line
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out /tmp/prof
timeout 600 python -m pytest tests/test_gpu_sparse.py tests/test_gpu_scale.py -m gpu -q -x 2>&1 | tail -2
timeout 900 python bench.py --config c5 --steps 100 --warmup 10 > gpurun_out/r2aj_bench_c5.json 2> gpurun_out/r2aj_bench_c5.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_c2.csv python bench.py --steps 40 --warmup 5 --no-cpu --no-e2e --no-converged > gpurun_out/r2aj_ncu_list.log 2>&1
POGS_B200_NO_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 300 --csv --log-file gpurun_out/r02_launches_c5_tiled.csv python bench.py --config c5 --steps 12 --warmup 3 --no-cpu --no-e2e --no-converged > gpurun_out/r2aj_ncu_list_c5.log 2>&1
for CFG in c2 c4; do
POGS_B200_NO_GRAPH=1 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_admm_pass -s 9 -c 2 -o /tmp/prof/admm_$CFG -f python bench.py --config $CFG --steps 8 --warmup 3 --no-cpu --no-e2e --no-converged > gpurun_out/r2aj_ncu_full_$CFG.log 2>&1
ncu -i /tmp/prof/admm_$CFG.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_admm_pass_$CFG.csv 2>/dev/null
done
rm -rf /tmp/prof
python - <<'PY'
import csv, json
d=json.loads(open("gpurun_out/r2aj_bench_c5.json").read().strip().splitlines()[-1])
print("c5", round(d["value"],1), d["roofline"]["frac"], d["roofline"].get("products_per_iteration"), d["sanity"]["parity"]["ok"], (d.get("converged") or {}).get("value"))
for cfg in ("c2","c4"):
    rows=list(csv.reader(open(f"gpurun_out/r02_ncu_full_admm_pass_{cfg}.csv")))
    hdr=rows[0]
    for w in ("gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum","launch__registers_per_thread"):
        i=hdr.index(w); print(cfg, w, [r[i] for r in rows[2:]])
PY
