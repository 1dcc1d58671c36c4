#!/bin/bash
# round 2, call AD: Cholesky chain replayed from a captured graph vs plain launches -- host / device split of the phase
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for g in 1 0; do
  POGS_B200_CHOL_GRAPH=$g POGS_B200_TRACE=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-converged > gpurun_out/r2ad_bench_g$g.json 2> gpurun_out/r2ad_trace_g$g.txt
  echo "== CHOL_GRAPH=$g"; grep trace gpurun_out/r2ad_trace_g$g.txt | tail -19 | head -13
done
timeout 900 python -m pytest tests/test_gpu_units.py tests/test_gpu_scale.py -m gpu -q -x > gpurun_out/r2ad_pytest.log 2>&1
tail -3 gpurun_out/r2ad_pytest.log
