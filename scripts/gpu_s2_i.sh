set -x
mkdir -p gpurun_out
timeout 300 python scripts/debug_gram.py 2>&1 | grep "gaussian"
timeout 200 python -m pytest tests/test_gpu_units.py -m gpu -q --timeout=120 -p no:cacheprovider -k "gram" 2>&1 | grep -v "^$" | tail -12
POGS_B200_TRACE=1 python bench.py --steps 200 --warmup 20 --no-cpu > gpurun_out/s2i_bench_c2_trace.json 2> gpurun_out/s2i_trace_c2.txt; grep "trace:" gpurun_out/s2i_trace_c2.txt | tail -22
python bench.py --steps 200 --warmup 20 > gpurun_out/s2i_bench_c2.json 2> gpurun_out/s2i_bench_c2.err; tail -c 3000 gpurun_out/s2i_bench_c2.json; tail -3 gpurun_out/s2i_bench_c2.err
python bench.py --config c4 --steps 200 --warmup 20 --no-cpu > gpurun_out/s2i_bench_c4.json 2> gpurun_out/s2i_bench_c4.err; tail -c 1500 gpurun_out/s2i_bench_c4.json
python bench.py --config c3 --steps 500 --warmup 20 --no-cpu > gpurun_out/s2i_bench_c3.json 2> gpurun_out/s2i_bench_c3.err; tail -c 1500 gpurun_out/s2i_bench_c3.json
timeout 600 python bench.py --config c5 --steps 50 --warmup 5 > gpurun_out/s2i_bench_c5.json 2> gpurun_out/s2i_bench_c5.err; tail -c 1500 gpurun_out/s2i_bench_c5.json; tail -3 gpurun_out/s2i_bench_c5.err
ncu --set full --clock-control none --import-source on -k regex:k_gram_tf32x3 -c 1 -o /tmp/prof_gram_c2 -f python bench.py --steps 6 --warmup 4 --no-e2e --no-cpu > gpurun_out/s2i_ncu_gram.log 2>&1; tail -2 gpurun_out/s2i_ncu_gram.log
ncu -i /tmp/prof_gram_c2.ncu-rep --page raw --csv > gpurun_out/s2i_ncu_gram_c2_raw.csv 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_fused_pass|k_gram|k_split|k_colacc|k_rowdot' -c 140 --csv --log-file gpurun_out/s2i_launches_setup_c2.csv python bench.py --steps 6 --warmup 4 --no-e2e --no-cpu > gpurun_out/s2i_ncu_setup.log 2>&1; tail -2 gpurun_out/s2i_ncu_setup.log
ls -la gpurun_out | tail -12
