# session 2, call D: first run of the tcgen05 Gram kernel (bounded by timeout), then the full suite
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_units.py -m gpu -q --timeout=120 -p no:cacheprovider -k "gram" 2>&1 | tail -30
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 --maxfail=15 -p no:cacheprovider -k "not gram" 2>&1 | tail -15
POGS_B200_TRACE=1 python bench.py --steps 200 --warmup 20 --no-cpu > gpurun_out/s2d_bench_c2_trace.json 2> gpurun_out/s2d_trace_c2.txt; grep "trace:" gpurun_out/s2d_trace_c2.txt | tail -22
python bench.py --steps 200 --warmup 20 --no-cpu > gpurun_out/s2d_bench_c2.json 2> gpurun_out/s2d_bench_c2.err; tail -c 2600 gpurun_out/s2d_bench_c2.json; tail -3 gpurun_out/s2d_bench_c2.err
