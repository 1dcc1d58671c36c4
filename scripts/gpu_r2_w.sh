#!/bin/bash
# round 2, call W: tiled sparse product -- (warps, chunk, stages) configurations
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for c in 0 1 2 3 4 5; do
  POGS_B200_TL=$c timeout 600 python scripts/dev/c5_probe.py c5 tiled > gpurun_out/r2w_probe_$c.log 2>&1
  tail -3 gpurun_out/r2w_probe_$c.log
done
POGS_B200_TL=1 POGS_B200_SPMV=tiled timeout 900 python -m pytest tests/test_gpu_sparse.py -m gpu -q -x 2>&1 | tail -3
