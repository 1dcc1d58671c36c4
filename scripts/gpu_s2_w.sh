set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 300 python -m pytest tests/test_gpu_units.py -m gpu -q --timeout=300 -p no:cacheprovider -k "symmetric or projection or setup" 2>&1 | tail -3
python bench.py --steps 200 --warmup 20 --no-cpu > gpurun_out/s2w_bench_c2.json 2> gpurun_out/s2w_bench_c2.err; tail -2 gpurun_out/s2w_bench_c2.err; tail -c 2600 gpurun_out/s2w_bench_c2.json
