#!/bin/bash
# round 2, call G: timing after the (W, B) planner change, full GPU suite, ncu evidence for the one-launch kernel
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
bash scripts/gpu_r2_c.sh 2>&1 | tail -6
POGS_B200_PASS_WB=4,1 POGS_B200_PASS_TIMING=1 timeout 600 python bench.py --config c4 --steps 200 --warmup 20 --no-cpu --no-e2e --no-converged > gpurun_out/r2g_bench_c4_w4b1.json 2>/dev/null
POGS_B200_PASS_WB=2,1 POGS_B200_PASS_TIMING=1 timeout 600 python bench.py --config c2 --steps 200 --warmup 20 --no-cpu --no-e2e --no-converged > gpurun_out/r2g_bench_c2_w2b1.json 2>/dev/null
python - <<'PY'
import json
for f in ("r2g_bench_c4_w4b1","r2g_bench_c2_w2b1"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"],1), [round(v,1) for v in d["roofline"].get("pass_phase_us") or []])
    except Exception as e:
        print(f, "ERR", e)
PY
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
# launch list of the captured loop (shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_c2.csv python bench.py --steps 40 --warmup 5 --no-cpu --no-e2e --no-converged > gpurun_out/r2g_ncu_list.log 2>&1
# full capture of one real launch of the one-launch iteration kernel (host-driven loop: launches alternate
# mode 1 [returns at its gate] / mode 0 [real])
for CFG in c2 c4; do
POGS_B200_NO_GRAPH=1 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_admm_pass -s 9 -c 2 -o gpurun_out/r02_prof_admm_$CFG -f python bench.py --config $CFG --steps 8 --warmup 3 --no-cpu --no-e2e --no-converged > gpurun_out/r2g_ncu_full_$CFG.log 2>&1
ncu -i gpurun_out/r02_prof_admm_$CFG.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_admm_pass_$CFG.csv 2>/dev/null
ncu -i gpurun_out/r02_prof_admm_$CFG.ncu-rep --page source --csv > gpurun_out/r02_ncu_source_admm_pass_$CFG.csv 2>/dev/null
done
ls -la gpurun_out | grep r02_
