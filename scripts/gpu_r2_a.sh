#!/bin/bash
# round 2, call A: parity suite incl. the bench-scale tests, first bench lines with the new contract keys
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -40 > gpurun_out/r2a_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench_c2_k20.json 2> gpurun_out/r2a_bench_c2_k20.err
timeout 600 python bench.py --config c3path > gpurun_out/r2a_bench_c3path.json 2> gpurun_out/r2a_bench_c3path.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2a_bench_ref.json 2> gpurun_out/r2a_bench_ref.err
tail -c 1500 gpurun_out/r2a_pytest.log
tail -c 600 gpurun_out/r2a_bench_c2_k20.err gpurun_out/r2a_bench_c3path.err gpurun_out/r2a_bench_ref.err
nproc
