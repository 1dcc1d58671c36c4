# session 2, call B: e2e timeline, kernel time list of the setup phase, full ncu capture of the
# single-pass kernel (a launch that really runs), C3/C4/C5 bench lines, ncu of the sparse products
set -x
mkdir -p gpurun_out
POGS_B200_TRACE=1 python bench.py --steps 200 --warmup 20 --no-cpu > gpurun_out/s2b_bench_c2_trace.json 2> gpurun_out/s2b_trace_c2.txt; tail -40 gpurun_out/s2b_trace_c2.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/s2b_launches_setup_c2.csv python bench.py --steps 6 --warmup 4 --no-e2e --no-cpu > gpurun_out/s2b_ncu_setup.log 2>&1; tail -2 gpurun_out/s2b_ncu_setup.log
ncu --set full --clock-control none --import-source on -k regex:k_fused_pass -s 2 -c 1 -o gpurun_out/s2b_prof_fused_c2 -f python bench.py --steps 6 --warmup 4 --no-e2e --no-cpu > gpurun_out/s2b_ncu_fused.log 2>&1; tail -2 gpurun_out/s2b_ncu_fused.log
python bench.py --config c3 --steps 500 --warmup 20 --no-cpu > gpurun_out/s2b_bench_c3.json 2> gpurun_out/s2b_bench_c3.err; tail -c 1500 gpurun_out/s2b_bench_c3.json
python bench.py --config c4 --steps 200 --warmup 20 --no-cpu > gpurun_out/s2b_bench_c4.json 2> gpurun_out/s2b_bench_c4.err; tail -c 1500 gpurun_out/s2b_bench_c4.json
timeout 600 python bench.py --config c5 --steps 20 --warmup 3 > gpurun_out/s2b_bench_c5.json 2> gpurun_out/s2b_bench_c5.err; tail -c 1500 gpurun_out/s2b_bench_c5.json; tail -3 gpurun_out/s2b_bench_c5.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_spmv' -s 220 -c 4 -o gpurun_out/s2b_prof_spmv_c5 -f python bench.py --config c5 --steps 4 --warmup 3 > gpurun_out/s2b_ncu_spmv.log 2>&1; tail -2 gpurun_out/s2b_ncu_spmv.log
ls -la gpurun_out/
