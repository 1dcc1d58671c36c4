#!/bin/bash
# round 2, call M: cone-form slice (tests), then the whole GPU suite
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cone.py -m gpu -q 2>&1 | tail -40 | cut -c1-400
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6
