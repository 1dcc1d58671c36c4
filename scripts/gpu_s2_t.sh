set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sparse.py -m gpu -q --timeout=300 -p no:cacheprovider 2>&1 | tail -4
timeout 400 python bench.py --config c5 --steps 50 --warmup 5 > gpurun_out/s2t_bench_c5.json 2> gpurun_out/s2t_bench_c5.err; tail -2 gpurun_out/s2t_bench_c5.err; tail -c 900 gpurun_out/s2t_bench_c5.json
timeout 300 python bench.py --config c5s --steps 50 --warmup 5 > gpurun_out/s2t_bench_c5s.json 2> gpurun_out/s2t_bench_c5s.err; tail -c 500 gpurun_out/s2t_bench_c5s.json
POGS_B200_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none -k regex:'k_spmv_blocked' -s 230 -c 2 -o /tmp/prof_spmvb_c5 -f python bench.py --config c5 --steps 4 --warmup 3 > gpurun_out/s2t_ncu_spmvb.log 2>&1; tail -2 gpurun_out/s2t_ncu_spmvb.log
ncu -i /tmp/prof_spmvb_c5.ncu-rep --page raw --csv > gpurun_out/s2t_ncu_spmvb_c5_raw.csv 2>/dev/null
