#!/bin/bash
# round 2, call AG: the several-tiles-per-CTA path of the tiled sparse product (POGS_B200_TL_GRID)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sparse.py -m gpu -q -x > gpurun_out/r2ag_pytest.log 2>&1
tail -4 gpurun_out/r2ag_pytest.log
