# session 2, call K: pipelined single-pass kernel + tensor-core inverse: suite, benches, ncu
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 --maxfail=15 -p no:cacheprovider 2>&1 | tail -25
POGS_B200_TRACE=1 python bench.py --steps 200 --warmup 20 --no-cpu > gpurun_out/s2k_bench_c2_trace.json 2> gpurun_out/s2k_trace_c2.txt; grep "trace:" gpurun_out/s2k_trace_c2.txt | tail -18
python bench.py --steps 200 --warmup 20 --no-cpu > gpurun_out/s2k_bench_c2.json 2> gpurun_out/s2k_bench_c2.err; tail -c 2300 gpurun_out/s2k_bench_c2.json; tail -3 gpurun_out/s2k_bench_c2.err
python bench.py --config c4 --steps 200 --warmup 20 --no-cpu > gpurun_out/s2k_bench_c4.json 2> gpurun_out/s2k_bench_c4.err; tail -c 1500 gpurun_out/s2k_bench_c4.json; tail -3 gpurun_out/s2k_bench_c4.err
ncu --set full --clock-control none --import-source on -k regex:k_fused_pass -s 60 -c 1 -o /tmp/prof_fused_c2 -f python bench.py --steps 6 --warmup 4 --no-e2e --no-cpu > gpurun_out/s2k_ncu_fused.log 2>&1; tail -2 gpurun_out/s2k_ncu_fused.log
ncu -i /tmp/prof_fused_c2.ncu-rep --page raw --csv > gpurun_out/s2k_ncu_fused_c2_raw.csv 2>/dev/null
ncu -i /tmp/prof_fused_c2.ncu-rep --page source --csv > gpurun_out/s2k_ncu_fused_c2_source.csv 2>/dev/null
ls -la gpurun_out | tail -8
