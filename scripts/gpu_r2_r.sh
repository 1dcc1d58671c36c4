#!/bin/bash
# round 2, call R: launch list + full capture of the vectorised blocked SpMV on C5
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out /tmp/prof
POGS_B200_NO_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 300 --csv --log-file gpurun_out/r02_launches_c5.csv python bench.py --config c5 --steps 12 --warmup 3 --no-cpu > gpurun_out/r2r_ncu_list.log 2>&1
POGS_B200_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none -k regex:k_spmv_blocked -s 400 -c 2 -o /tmp/prof/spmv -f python bench.py --config c5 --steps 6 --warmup 3 --no-cpu > gpurun_out/r2r_ncu_full.log 2>&1
ncu -i /tmp/prof/spmv.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_spmv_blocked_c5.csv 2>/dev/null
rm -rf /tmp/prof
ls -la gpurun_out | grep r02_
