#!/bin/bash
# round 2, call V: tiled sparse product -- ring configurations, build timeline of a first solver, ncu --set full
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out /tmp/prof
POGS_B200_SPMV=tiled timeout 900 python -m pytest tests/test_gpu_sparse.py -m gpu -q -x > gpurun_out/r2v_pytest.log 2>&1
tail -3 gpurun_out/r2v_pytest.log
for c in 0 1 2; do
  POGS_B200_TL=$c POGS_B200_TRACE=1 timeout 600 python scripts/dev/c5_probe.py c5 tiled > gpurun_out/r2v_probe_$c.log 2>&1
  grep -v "trace" gpurun_out/r2v_probe_$c.log | tail -4
done
grep "trace" gpurun_out/r2v_probe_0.log | head -40
POGS_B200_SPMV=tiled POGS_B200_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none -k regex:k_spmv_tiled -s 120 -c 2 -o /tmp/prof/spmvt -f python bench.py --config c5 --steps 6 --warmup 3 --no-cpu --no-e2e --no-converged > gpurun_out/r2v_ncu_full.log 2>&1
ncu -i /tmp/prof/spmvt.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_spmv_tiled_c5.csv 2>/dev/null
tail -3 gpurun_out/r2v_ncu_full.log
