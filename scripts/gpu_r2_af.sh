#!/bin/bash
# round 2, call AF: fold kernel with one row per thread -- sparse tests, C5 loop time
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sparse.py -m gpu -q -x 2>&1 | tail -2
timeout 600 python scripts/dev/c5_probe.py c5 tiled 2>&1 | tail -3
