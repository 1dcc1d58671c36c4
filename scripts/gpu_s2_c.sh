# session 2, call C: GPU test suite on the pooled allocator + graph-captured CGLS loop, e2e
# timeline, C5 bench, ncu captures exported to CSV on the box (reports are too big to bring back)
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 --maxfail=15 -p no:cacheprovider 2>&1 | tail -25
POGS_B200_TRACE=1 python bench.py --steps 200 --warmup 20 --no-cpu > gpurun_out/s2c_bench_c2.json 2> gpurun_out/s2c_trace_c2.txt; grep "trace:" gpurun_out/s2c_trace_c2.txt | tail -22; tail -c 2600 gpurun_out/s2c_bench_c2.json
timeout 600 python bench.py --config c5 --steps 50 --warmup 5 > gpurun_out/s2c_bench_c5.json 2> gpurun_out/s2c_bench_c5.err; tail -c 1500 gpurun_out/s2c_bench_c5.json; tail -3 gpurun_out/s2c_bench_c5.err
timeout 300 python bench.py --config c5s --steps 50 --warmup 5 > gpurun_out/s2c_bench_c5s.json 2> gpurun_out/s2c_bench_c5s.err; tail -c 1200 gpurun_out/s2c_bench_c5s.json; tail -3 gpurun_out/s2c_bench_c5s.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/s2c_launches_setup_c2.csv python bench.py --steps 6 --warmup 4 --no-e2e --no-cpu > gpurun_out/s2c_ncu_setup.log 2>&1; tail -2 gpurun_out/s2c_ncu_setup.log
ncu --set full --clock-control none --import-source on -k regex:k_fused_pass -s 2 -c 1 -o /tmp/prof_fused_c2 -f python bench.py --steps 6 --warmup 4 --no-e2e --no-cpu > gpurun_out/s2c_ncu_fused.log 2>&1; tail -2 gpurun_out/s2c_ncu_fused.log
ncu -i /tmp/prof_fused_c2.ncu-rep --page raw --csv > gpurun_out/s2c_ncu_fused_c2_raw.csv 2>/dev/null
ncu -i /tmp/prof_fused_c2.ncu-rep --page source --csv > gpurun_out/s2c_ncu_fused_c2_source.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:k_fused_pass -s 2 -c 1 -o /tmp/prof_fused_c4 -f python bench.py --config c4 --steps 6 --warmup 4 --no-e2e --no-cpu > gpurun_out/s2c_ncu_fused_c4.log 2>&1; tail -2 gpurun_out/s2c_ncu_fused_c4.log
ncu -i /tmp/prof_fused_c4.ncu-rep --page raw --csv > gpurun_out/s2c_ncu_fused_c4_raw.csv 2>/dev/null
POGS_B200_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_spmv' -s 220 -c 4 -o /tmp/prof_spmv_c5 -f python bench.py --config c5 --steps 4 --warmup 3 > gpurun_out/s2c_ncu_spmv.log 2>&1; tail -2 gpurun_out/s2c_ncu_spmv.log
ncu -i /tmp/prof_spmv_c5.ncu-rep --page raw --csv > gpurun_out/s2c_ncu_spmv_c5_raw.csv 2>/dev/null
ls -la gpurun_out/ /tmp/*.ncu-rep
